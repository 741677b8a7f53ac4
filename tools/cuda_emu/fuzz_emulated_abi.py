"""Fuzz of the exported sync_and_demodulate / subtract_signal2 / subtract_signal (wsprd.h:76-105) of the CUDA path (emulated)
against the compiled reference's own functions, random parameters.
    WSPR_B200_LIB=<emulated build> python tools/cuda_emu/fuzz_emulated_abi.py     (180 calls: 0 mismatches)"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import pyoracle as po
import rtlsdr_wsprd_b200 as w
import helpers as H
FP = C.POINTER(C.c_float); UP = C.POINTER(C.c_ubyte)
ref, lib = po.ref(), w.library()
I, Q, _ = H.make_corpus(3, 2, start=77)
bad = 0; n = 0; t0 = time.time()
rng = np.random.default_rng(9)
def call_sync(L, i0, q0, npts, freq, ifmin, ifmax, fstep, shift, lagmin, lagmax, lagstep, drift, symfac, mode):
    f, sh, dr, sy = C.c_float(freq), C.c_int(shift), C.c_float(drift), C.c_float(0.0)
    sym = (C.c_ubyte * 162)()
    L.sync_and_demodulate(i0.ctypes.data_as(FP), q0.ctypes.data_as(FP), C.c_long(npts) if L is ref else npts, sym, C.byref(f), ifmin, ifmax, C.c_float(fstep),
                          C.byref(sh), lagmin, lagmax, lagstep, C.byref(dr), symfac, C.byref(sy), mode)
    return f.value, sh.value, dr.value, sy.value, bytes(sym)
for L in (ref, lib):
    L.sync_and_demodulate.restype = None
for k in range(120):
    c = k % 2
    npts = int(rng.choice([45000, 45000, 45000, 43000, 44700]))
    i0, q0 = np.zeros(45000 + 512, np.float32), np.zeros(45000 + 512, np.float32)
    i0[:npts], q0[:npts] = I[c, :npts], Q[c, :npts]
    freq = float(rng.uniform(-110, 110)); shift = int(rng.integers(-300, 3000)); drift = float(rng.choice([0.0, 0.0, rng.uniform(-4, 4)]))
    mode = int(rng.integers(0, 3))
    if mode == 0:
        args = (0, 0, 0.0, shift, shift - 128, shift + 128, int(rng.choice([8, 16])), drift, 50, 0)
    elif mode == 1:
        args = (-2, 2, float(rng.choice([0.1, 0.25, 0.05])), shift, shift, shift, 1, drift, 50, 1)
    else:
        args = (0, 0, 0.0, shift, shift, shift, 1, drift, int(rng.choice([50, 64])), 2)
    a = call_sync(ref, i0.copy(), q0.copy(), npts, freq, *args)
    b = call_sync(lib, i0.copy(), q0.copy(), npts, freq, *args)
    ok = (np.float32(a[0]).tobytes(), a[1], np.float32(a[3]).tobytes()) == (np.float32(b[0]).tobytes(), b[1], np.float32(b[3]).tobytes()) and (mode != 2 or a[4] == b[4])
    n += 1; bad += not ok
    if not ok: print("SYNC MISMATCH", k, mode, npts, freq, shift, drift, a[:4], b[:4])
# subtract_signal2 / subtract_signal with random channel symbols
for k in range(30):
    c = k % 2
    chan = rng.integers(0, 4, 162).astype(np.uint8)
    f0 = float(rng.uniform(-110, 110)); shift = int(rng.integers(-200, 3500)); drift = float(rng.choice([0.0, rng.uniform(-4, 4)]))
    for name in ("subtract_signal2", "subtract_signal"):
        outs = []
        for L in (ref, lib):
            ia, qa = I[c].copy(), Q[c].copy()
            getattr(L, name)(ia.ctypes.data_as(FP), qa.ctypes.data_as(FP), C.c_long(45000), C.c_float(f0), shift, C.c_float(drift), chan.ctypes.data_as(UP))
            outs.append((ia, qa))
        ok = np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
        n += 1; bad += not ok
        if not ok: print("SUB MISMATCH", name, k, f0, shift, drift)
print("calls", n, "mismatches", bad, "time", round(time.time() - t0, 1))
