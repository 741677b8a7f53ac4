"""Experiment: how much do long-running one-warp Fano CTAs (the shape of k_chain_fano) slow the bulk decode kernels that
share their SMs?  Decodes a batch with the chain cut short (WSPR_DEBUG_CHAIN_MAXCYCLES, wrong results, timing only) while
a background thread keeps K one-warp Fano CTAs of hopeless attempts in flight.  usage: exp_interference.py [ncap]"""
import os, sys, threading, time
os.environ["WSPR_DEBUG_CHAIN_MAXCYCLES"] = "64"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import rtlsdr_wsprd_b200 as w
import helpers as H
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
I, Q, _ = H.make_corpus(3, 64)
I = np.tile(I, (n // 64, 1)); Q = np.tile(Q, (n // 64, 1))
rng = np.random.default_rng(1)
stop = False
lat = []

def background(k):
    sym = rng.integers(0, 256, size=(32 * k, 162), dtype=np.uint8)
    while not stop:
        r = w.fano_batch(sym)
        lat.append(float(np.mean(r["clocks"])) / float(np.mean(r["cycles"])))

with w.BatchDecoder(n) as d:
    d.upload(I, Q)
    for k in (0, 37, 74, 148, 296, 592):
        stop = False
        lat.clear()
        th = None
        if k:
            th = threading.Thread(target=background, args=(k,))
            th.start()
            time.sleep(0.3)
        ms = []
        for it in range(4):
            d.upload(I, Q)
            ms.append(d.decode())
        stop = True
        if th:
            th.join()
        print("background fano CTAs %4d: decode ms %s  fano clocks/cycle %s" % (k, [round(x, 1) for x in ms],
              [round(x) for x in lat[:4]]), flush=True)
