"""Fuzz of the front end under the host emulation against the reference's own rtlsdr_callback (oracle/_ref): whole-stream
decimation of several streams per call (ragged lengths, both rails, runs of zeros, clipped noise) and the streaming form
(random push sizes, two slots, filter state carried across the slot switch).
    WSPR_B200_LIB=<emulated build> python tools/cuda_emu/fuzz_emulated_frontend.py     (93 streams + 40 streaming runs: 0 mismatches)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import pyoracle as po
import rtlsdr_wsprd_b200 as w

def raw_stream(rng, n_iq, kind):
    if kind == 0: return rng.integers(0, 256, size=2 * n_iq, dtype=np.uint8)
    if kind == 1: return np.full(2 * n_iq, 0, np.uint8)
    if kind == 2: return np.full(2 * n_iq, 255, np.uint8)
    if kind == 3:
        r = rng.integers(0, 256, size=2 * n_iq, dtype=np.uint8); r[rng.random(2 * n_iq) < 0.3] = 0; return r
    return np.clip(127.5 + 60 * rng.standard_normal(2 * n_iq), 0, 255).astype(np.uint8)

bad = 0; t0 = time.time(); nb = ns = 0
for seed in range(40):
    rng = np.random.default_rng(5000 + seed)
    # ---- whole streams, several per call ----
    nstreams = int(rng.integers(1, 5))
    n_iq = 6401 * int(rng.integers(1, 40)) + int(rng.integers(0, 6401)); n_iq -= n_iq % 4
    raw = np.stack([raw_stream(rng, n_iq, int(rng.integers(5))) for _ in range(nstreams)])
    Ig, Qg, nout = w.decimate_batch(raw, max_out=128, device=0)
    for s in range(nstreams):
        f = po.RefFrontend(); f.push(raw[s]); ir, qr = f.read()
        ok = nout == len(ir) and np.array_equal(Ig[s, :nout], ir) and np.array_equal(Qg[s, :nout], qr)
        nb += 1; bad += not ok
        if not ok: print("BATCH MISMATCH", seed, s, n_iq, nout, len(ir))
    # ---- streaming: random pushes, two slots, filter state carried across ----
    slot = 64
    total = 6401 * int(rng.integers(70, 140)); total -= total % 4
    stream = raw_stream(rng, total, int(rng.integers(5)))
    ref = po.RefFrontend()
    with w.FrontEnd(1, slot_samples=slot) as fe:
        pos = 0; cut = 2 * (6401 * slot + int(rng.integers(0, 3000)) * 4)    # the first slot ends after a bit more than `slot` outputs' worth
        got = []
        for end in (cut, 2 * total):
            while pos < end:
                nbytes = min(end - pos, 8 * int(rng.integers(1, 20000)))
                fe.push(stream[pos:pos + nbytes]); pos += nbytes
            n = fe.swap(); I, Q, n2 = fe.read(); got.append((I[0, :n].copy(), Q[0, :n].copy()))
    ref.push(stream[:cut]); i1, q1 = ref.read()
    # the reference callback keeps appending to one buffer: slot 2 = what it produces after the first `cut` bytes
    ref2 = po.RefFrontend(); ref2.push(stream); ia, qa = ref2.read()
    k1 = min(len(i1), slot)
    ok = np.array_equal(got[0][0], i1[:k1]) and np.array_equal(got[0][1], q1[:k1])
    rest_i, rest_q = ia[len(i1):][:slot], qa[len(i1):][:slot]
    ok = ok and np.array_equal(got[1][0][:len(rest_i)], rest_i) and np.array_equal(got[1][1][:len(rest_q)], rest_q) and len(got[1][0]) == len(rest_i)
    ns += 1; bad += not ok
    if not ok: print("STREAM MISMATCH", seed, total, cut, [len(x[0]) for x in got], len(i1), len(ia))
print("whole streams", nb, "streaming runs", ns, "mismatches", bad, "time", round(time.time() - t0, 1))
