"""Fuzz of the persistent-hashtable option (reference -H, wsprd.c:481-494,842-852): sequences of four captures with random
type-1 / type-2 / type-3 (hashed callsign) messages decoded one after the other in one directory, with and without a seeded
hashtable.txt; the oracle and the compiled reference must give identical results AND identical hashtable.txt after every
call.   python tools/fuzz_hashtable_oracle_vs_ref.py   (profiles/r2_oracle_fuzz.txt: 160 captures, 223 hashed spots, 0 mismatches)"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
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
def run(lib, caps, seedfile):
    out, old = [], os.getcwd()
    opt = po.default_options(usehashtable=1)
    with tempfile.TemporaryDirectory(prefix="wspr_htf_") as d:
        os.chdir(d)
        try:
            if seedfile: open("hashtable.txt", "w").write(seedfile)
            for i, q in caps:
                r = po.decode(lib, i.copy(), q.copy(), opt, cwd_scratch=False)[0]
                out.append((r, open("hashtable.txt").read() if os.path.exists("hashtable.txt") else None))
        finally:
            os.chdir(old)
    return out
bad = 0; t0 = time.time(); hashed = 0
for seed in range(40):
    rng = np.random.default_rng(1000 + seed)
    caps = []
    for c in range(4):
        sig = [dict(message=msg(rng), f0=-90.0 + 30.0 * k + float(rng.uniform(-3, 3)), dt0=float(rng.uniform(-0.5, 0.5)), snr=float(rng.uniform(-20, -8))) for k in range(int(rng.integers(1, 7)))]
        caps.append(corpus.make_capture(78, seed * 10 + c, sig, H.channel_symbols))
    seedfile = H.HASHTABLE_SEED_FILE if seed % 2 else None
    a = run(po.ref(), caps, seedfile); b = run(po.oracle(), caps, seedfile)
    for (ra, fa), (rb, fb) in zip(a, b):
        ok = H.results_equal(ra, rb) and fa == fb
        hashed += sum(1 for x in ra if x["message"].decode().startswith("<"))
        if not ok:
            bad += 1; print("MISMATCH", seed, H.diff_results(ra, rb), fa == fb)
print("sequences", 40, "captures", 160, "hashed-callsign spots", hashed, "mismatches", bad, "time", round(time.time() - t0, 1))
