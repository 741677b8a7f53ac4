"""Latency of ONE capture through the reference's own entry point wspr_decode() (what the daemon calls once per 2-minute slot,
rtlsdr_wsprd.c:316) on the GPU library, next to the reference's CPU code on the same capture.  usage: latency_single.py [n]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import rtlsdr_wsprd_b200 as w
from oracle import pyoracle as po
import helpers as H
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
for config, what in ((3, "10 signals, -28..-10 dB"), (2, "1 signal, -20 dB")):
    I, Q, _ = H.make_corpus(config, n, start=5000)
    w.wspr_decode(I[0].copy(), Q[0].copy())                      # context creation, module load
    ref = po.ref() or po.oracle()
    gpu, cpu, same = [], [], 0
    for c in range(n):
        i, q = I[c].copy(), Q[c].copy()
        t0 = time.perf_counter()
        r = w.wspr_decode(i, q)
        gpu.append((time.perf_counter() - t0) * 1e3)
        t0 = time.perf_counter()
        a, _, _ = po.decode(ref, I[c], Q[c])
        cpu.append((time.perf_counter() - t0) * 1e3)
        same += int(H.results_equal(a, r))
    g, cp = np.array(gpu), np.array(cpu)
    print("config %d (%s), %d captures, one wspr_decode() call each (host arrays in, host results out): GPU median %.1f ms, mean %.1f, "
          "min %.1f, max %.1f | reference C on one core: median %.1f ms, mean %.1f, max %.1f | identical results %d/%d"
          % (config, what, n, np.median(g), g.mean(), g.min(), g.max(), np.median(cp), cp.mean(), cp.max(), same, n), flush=True)
