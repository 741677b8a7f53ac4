"""Fuzz of the batch path under the host emulation: one context decodes a random batch (2-6 random captures, random options) and
then a second, differently sized batch -- leftovers of the first must not show in the second --; every capture of both must
equal the oracle's decode of that capture alone (all result fields, post-subtraction samples).
    WSPR_B200_LIB=<emulated build> python tools/cuda_emu/fuzz_emulated_batches.py 0 60"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import multiprocessing as mp
from oracle import pyoracle as po
import helpers as H
from rtlsdr_wsprd_b200 import corpus

MSGS = ["K1JT FN20 20", "VA2GKA FN35 37", "W1AW FN31 30", "G4JNT IO90 10", "PJ4/K1ABC 37", "<PJ4/K1ABC> FK52UD 37", "K1ABC/7 33",
        "<K1JT> FN20AB 20", "DL1ABC JO62 23", "JA1XYZ PM95 27", "ZL3GHI RE66 0", "EA4PQR IN80 60", "VK2DEF QF56 3", "K9AN EN50 33"]


def capture(rng, key):
    plan = [dict(message=MSGS[int(rng.integers(len(MSGS)))], f0=float(rng.uniform(-120, 120)), dt0=float(rng.uniform(-1.5, 1.5)),
                 snr=float(rng.uniform(-32, -5)), drift=float(rng.choice([0.0, 0.0, rng.uniform(-4, 4)]))) for _ in range(int(rng.integers(0, 9)))]
    return corpus.make_capture(79, key, plan, H.channel_symbols)


def one(seed):
    import rtlsdr_wsprd_b200 as w
    rng = np.random.default_rng(7000 + seed)
    opt = dict(quickmode=int(rng.random() < 0.25), npasses=int(rng.choice([1, 2, 2, 3])), subtraction=int(rng.random() < 0.85))
    sizes = [int(rng.integers(2, 7)), int(rng.integers(1, 5))]
    bad, total = 0, 0
    with w.BatchDecoder(6, corpus.NSAMP) as d:
        for b, n in enumerate(sizes):
            caps = [capture(rng, seed * 100 + 10 * b + c) for c in range(n)]
            I = np.stack([c[0] for c in caps]); Q = np.stack([c[1] for c in caps])
            d.upload(I, Q)
            d.decode(w.default_options(**opt))
            spots, cnt, Io, Qo = d.download(samples=True)
            for c in range(n):
                a, ia, qa = po.decode(po.oracle(), I[c], Q[c], po.default_options(**opt))
                ok = H.results_equal(a, spots[c, : cnt[c]]) and np.array_equal(ia, Io[c]) and np.array_equal(qa, Qo[c])
                bad += not ok
                total += len(a)
    return seed, bad, sum(sizes), total, opt


if __name__ == "__main__":
    lo, hi = int(sys.argv[1]), int(sys.argv[2])
    po.oracle()
    t = time.time(); bad = caps = spots = 0
    with mp.get_context("fork").Pool(8) as pool:
        for seed, b, n, s, opt in pool.imap_unordered(one, range(lo, hi)):
            bad += b; caps += n; spots += s
            if b: print("MISMATCH seed", seed, opt, b, flush=True)
    print("contexts", hi - lo, "batches", 2 * (hi - lo), "captures", caps, "spots", spots, "mismatching captures", bad, "time", round(time.time() - t, 1), flush=True)
