"""Fuzz: the oracle (oracle/wspr_oracle.c) against the compiled reference (oracle/_ref) on random plans and options -- 0 to 12
signals, SNR -33..0 dB, f0 +-150 Hz (past the +-110 Hz search band), dt -2.2..+2.6 s, drifting signals, type-1/2/3 and hashed
messages, truncated captures, -Q, 1-4 passes, subtraction on / off, several dial frequencies; every result field and the
post-subtraction samples must be identical.   python tools/fuzz_oracle_vs_ref.py <first seed> <last seed + 1>
(profiles/r2_oracle_fuzz.txt: 1000 captures, 3530 spots, 0 mismatches)"""
import sys, time, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import pyoracle as po
import helpers as H
from rtlsdr_wsprd_b200 import corpus
import multiprocessing as mp

MSGS = ["K1JT FN20 20", "VA2GKA FN35 37", "W1AW FN31 30", "G4JNT IO90 10", "PJ4/K1ABC 37", "<PJ4/K1ABC> FK52UD 37", "K1ABC/7 33",
        "<K1JT> FN20AB 20", "DL1ABC JO62 23", "JA1XYZ PM95 27", "ZL3GHI RE66 0", "EA4PQR IN80 60", "VK2DEF QF56 3", "K9AN EN50 33"]

def one(seed):
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
    o = po.default_options(**opt)
    a, ia, qa = po.decode(po.ref(), i, q, o)
    b, ib, qb = po.decode(po.oracle(), i, q, o)
    ok = H.results_equal(a, b) and np.array_equal(ia, ib) and np.array_equal(qa, qb)
    return seed, ok, len(a), nsig, opt, (None if ok else H.diff_results(a, b))

if __name__ == "__main__":
    lo, hi = int(sys.argv[1]), int(sys.argv[2])
    po.ref(); po.oracle()
    t = time.time()
    bad = 0; spots = 0; n = 0
    with mp.get_context("fork").Pool(8) as pool:
        for seed, ok, ns, nsig, opt, diff in pool.imap_unordered(one, range(lo, hi)):
            n += 1; spots += ns
            if not ok:
                bad += 1
                print("MISMATCH seed", seed, nsig, opt, diff, flush=True)
    print("captures", n, "spots", spots, "mismatches", bad, "time", round(time.time() - t, 1), flush=True)
