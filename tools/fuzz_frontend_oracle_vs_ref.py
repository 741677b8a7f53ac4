"""Fuzz: the oracle's restatement of rtlsdr_callback (oracle_decimate) against the reference's own callback (oracle/_ref,
extracted by line range from rtlsdr_wsprd.c:126-244) on random raw streams: uniform bytes, both rails (the int8 -(-128)
wrap), runs of zeros, clipped Gaussian noise; fed to the callback in random chunk sizes.   python tools/fuzz_frontend_oracle_vs_ref.py
(profiles/r2_oracle_fuzz.txt: 60 streams, 0 mismatches)"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po
orc = po.oracle()
orc.oracle_decimate.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
bad = 0
t = time.time()
for seed in range(60):
    rng = np.random.default_rng(seed)
    nblk = int(rng.integers(1, 120))
    n_iq = 6401 * nblk + int(rng.integers(0, 6401))
    n_iq -= n_iq % 4                                   # the callback consumes 8 bytes (4 IQ pairs) at a time
    kind = seed % 5
    if kind == 0: raw = rng.integers(0, 256, size=2 * n_iq, dtype=np.uint8)
    elif kind == 1: raw = np.full(2 * n_iq, 0, np.uint8)          # rail: -128 after the offset, the int8 negation wrap
    elif kind == 2: raw = np.full(2 * n_iq, 255, np.uint8)
    elif kind == 3:
        raw = rng.integers(0, 256, size=2 * n_iq, dtype=np.uint8); raw[rng.random(2 * n_iq) < 0.3] = 0
    else:
        raw = np.clip(127.5 + 60 * rng.standard_normal(2 * n_iq), 0, 255).astype(np.uint8)
    chunk = int(rng.choice([65536, 8, 4096, 262144, 8 * int(rng.integers(1, 5000))]))
    f = po.RefFrontend()
    f.push(raw, chunk=chunk)
    ir, qr = f.read()
    io, qo = np.zeros(256, np.float32), np.zeros(256, np.float32)
    n = orc.oracle_decimate(raw.ctypes.data, n_iq, io.ctypes.data, qo.ctypes.data, 256)
    ok = n == len(ir) and np.array_equal(io[:n], ir) and np.array_equal(qo[:n], qr)
    bad += not ok
    if not ok: print("MISMATCH", seed, kind, n_iq, chunk, n, len(ir))
print("streams", 60, "mismatches", bad, "time", round(time.time() - t, 1))
