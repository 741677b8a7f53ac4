"""Fuzz of the CUDA path under the host emulation against the oracle: random plans and options like tools/fuzz_oracle_vs_ref.py
(0-12 signals, SNR -33..0 dB, f0 +-150 Hz, dt -2.2..+2.6 s, drift, all message types, truncated captures, -Q, 1-4 passes,
subtraction on / off), one wspr_decode() call per capture; every result field and the post-subtraction samples must be identical.
    python tools/cuda_emu/build.py /tmp/emu && WSPR_B200_LIB=/tmp/emu/libwsprd_b200_emu.so python tools/cuda_emu/fuzz_emulated_path.py 0 400
This is what found the row stride that was too short for captures whose length is below 256 modulo 512 (ctx_init)."""
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


def one(seed):
    import rtlsdr_wsprd_b200 as w
    rng = np.random.default_rng(seed)
    nsig = int(rng.integers(0, 13))
    plan = []
    for _ in range(nsig):
        plan.append(dict(message=MSGS[int(rng.integers(len(MSGS)))], f0=float(rng.uniform(-150, 150)), dt0=float(rng.uniform(-2.2, 2.6)),
                         snr=float(rng.uniform(-33, 0)), drift=float(rng.choice([0.0, 0.0, 0.0, rng.uniform(-4, 4)]))))
    i, q = corpus.make_capture(77, seed, plan, H.channel_symbols)
    if rng.random() < 0.1:
        n = int(rng.integers(30000, 45000)); i, q = np.ascontiguousarray(i[:n]), np.ascontiguousarray(q[:n])
    opt = dict(quickmode=int(rng.random() < 0.2), npasses=int(rng.choice([1, 2, 2, 2, 3, 4])), subtraction=int(rng.random() < 0.85))
    if rng.random() < 0.3:
        opt["freq"] = int(rng.choice([14095600, 7038600, 144489000]))
    a, ia, qa = po.decode(po.oracle(), i, q, po.default_options(**opt))
    ig, qg = i.copy(), q.copy()
    b = w.wspr_decode(ig, qg, len(ig), w.default_options(**opt))
    ok = H.results_equal(a, b) and np.array_equal(ia, ig) and np.array_equal(qa, qg)
    return seed, ok, len(a), len(i), opt, (None if ok else H.diff_results(a, b))

if __name__ == "__main__":
    lo, hi = int(sys.argv[1]), int(sys.argv[2])
    po.oracle()
    t = time.time(); bad = 0; spots = 0; n = 0; short = 0
    with mp.get_context("fork").Pool(8) as pool:
        for seed, ok, ns, length, opt, diff in pool.imap_unordered(one, range(lo, hi)):
            n += 1; spots += ns; short += length < 45000
            if not ok:
                bad += 1; print("MISMATCH seed", seed, length, opt, diff, flush=True)
    print("captures", n, "of which truncated", short, "spots", spots, "mismatches", bad, "time", round(time.time() - t, 1), flush=True)
