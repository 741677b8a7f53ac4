"""-H fuzz of the CUDA path under the host emulation against the compiled reference (reference -H, wsprd.c:481-494,842-852):
sequences of four wspr_decode() calls in one directory with random type-1 / 2 / 3 messages, with and without a seeded
hashtable.txt; results and hashtable.txt after every call must be identical.
    WSPR_B200_LIB=<emulated build> python tools/cuda_emu/fuzz_emulated_hashtable.py 40     (160 calls, 223 hashed spots: 0 mismatches)"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import multiprocessing as mp
from oracle import pyoracle as po
import helpers as H
from rtlsdr_wsprd_b200 import corpus

CALLS = ["K1JT", "W1AW", "G4JNT", "VA2GKA", "DL1ABC", "JA1XYZ", "PJ4/K1ABC", "K1ABC/7", "ZL3GHI", "EA4PQR", "F/OH2MNO", "VK2DEF/P"]
GRIDS6 = ["FN20AB", "FN31PR", "IO90AA", "FN35AA", "JO62QM", "PM95VQ", "FK52UD", "DM33AA", "RE66HN", "IN80DK", "JN18EU", "QF56OD"]
GRIDS4 = [g[:4] for g in GRIDS6]
def msg(rng):
    k = int(rng.integers(len(CALLS))); p = int(rng.choice(corpus.POWERS))
    t = rng.random()
    c = CALLS[k]
    if "/" in c:
        return ("%s %d" % (c, p)) if t < 0.6 else ("<%s> %s %d" % (c, GRIDS6[k], p))
    return ("%s %s %d" % (c, GRIDS4[k], p)) if t < 0.6 else ("<%s> %s %d" % (c, GRIDS6[k], p))

def run_ref(caps, seedfile):
    out, old = [], os.getcwd()
    opt = po.default_options(usehashtable=1)
    with tempfile.TemporaryDirectory(prefix="wspr_htf_") as d:
        os.chdir(d)
        try:
            if seedfile: open("hashtable.txt", "w").write(seedfile)
            for i, q in caps:
                r = po.decode(po.ref(), i.copy(), q.copy(), opt, cwd_scratch=False)[0]
                out.append((r, open("hashtable.txt").read() if os.path.exists("hashtable.txt") else None))
        finally:
            os.chdir(old)
    return out

def run_gpu(caps, seedfile):
    import rtlsdr_wsprd_b200 as w
    out, old = [], os.getcwd()
    opt = w.default_options(usehashtable=1)
    with tempfile.TemporaryDirectory(prefix="wspr_htg_") as d:
        os.chdir(d)
        try:
            if seedfile: open("hashtable.txt", "w").write(seedfile)
            for i, q in caps:
                r = w.wspr_decode(i.copy(), q.copy(), len(i), opt)
                out.append((r, open("hashtable.txt").read() if os.path.exists("hashtable.txt") else None))
        finally:
            os.chdir(old)
    return out

def one(seed):
    rng = np.random.default_rng(1000 + seed)
    caps = []
    for c in range(4):
        sig = [dict(message=msg(rng), f0=-90.0 + 30.0 * k + float(rng.uniform(-3, 3)), dt0=float(rng.uniform(-0.5, 0.5)), snr=float(rng.uniform(-20, -8))) for k in range(int(rng.integers(1, 7)))]
        caps.append(corpus.make_capture(78, seed * 10 + c, sig, H.channel_symbols))
    seedfile = H.HASHTABLE_SEED_FILE if seed % 2 else None
    a = run_ref(caps, seedfile); b = run_gpu(caps, seedfile)
    bad = 0; hashed = 0
    for (ra, fa), (rb, fb) in zip(a, b):
        ok = H.results_equal(ra, rb) and fa == fb
        hashed += sum(1 for x in ra if x["message"].decode().startswith("<"))
        bad += not ok
    return seed, bad, hashed

if __name__ == "__main__":
    po.ref(); po.oracle()
    t0 = time.time(); bad = 0; hashed = 0; n = 0
    with mp.get_context("fork").Pool(8) as pool:
        for seed, b, h in pool.imap_unordered(one, range(int(sys.argv[1]))):
            n += 1; bad += b; hashed += h
            if b: print("MISMATCH", seed, b, flush=True)
    print("sequences", n, "captures", 4 * n, "hashed-callsign spots", hashed, "mismatching calls", bad, "time", round(time.time() - t0, 1))
