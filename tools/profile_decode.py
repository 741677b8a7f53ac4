"""Decode a config-3 batch a few times (for ncu launch lists / captures).  usage: profile_decode.py [ncap] [iters] [config]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import rtlsdr_wsprd_b200 as w
import helpers as H
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
config = int(sys.argv[3]) if len(sys.argv) > 3 else 3
base = min(n, 64)
I, Q, _ = H.make_corpus(config, base)
I = np.tile(I, ((n + base - 1) // base, 1))[:n]; Q = np.tile(Q, ((n + base - 1) // base, 1))[:n]
with w.BatchDecoder(n) as d:
    for it in range(iters):
        d.upload(I, Q)
        ms = d.decode()
        spots, nres = d.download()
        print("n", n, "decode ms", round(ms, 2), "captures/s", round(n / ms * 1e3, 1), "spots", int(nres.sum()), "launches", w.kernel_launches(), "rounds/deferred", d.schedule_stats(), flush=True)
