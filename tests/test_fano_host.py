"""The device Fano loop (rtlsdr_wsprd_b200/csrc/wspr_fano.cuh) compiled for the HOST (tools/fano_host_check.cpp: same
template source, one lane, scratch in ordinary memory) against fano() of the oracle on random soft-symbol vectors from
clean to hopeless, time-outs included.  Covers the logic of both instantiations -- exact / decode (time-out test every
256 trips, maxnp not tracked) -- and the re-arming of a lane (several attempts back to back through one lane, which is what
the queue-fed worker warps do) without a GPU; the GPU suite repeats the comparison for the kernels the library runs."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po
import helpers as H

UP = C.POINTER(C.c_ubyte)


@pytest.fixture(scope="module")
def host_fano(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("fano_host") / "libfano_host.so")
    src = os.path.join(H.ROOT, "tools", "fano_host_check.cpp")
    try:
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src], check=True, capture_output=True, text=True)
    except FileNotFoundError:
        pytest.skip("g++ not available")
    lib = C.CDLL(out)
    lib.fano_host.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p]
    return lib


def vectors(n, seed):
    orc = po.oracle()
    rng = np.random.default_rng(seed)
    msgs = ["K1JT FN20 20", "VA2GKA FN35 37", "G4JNT IO90 60", "<K1ABC> FN42AX 10", "PJ4/K1ABC 37"]
    out = []
    for k in range(n):
        sym = H.channel_symbols(msgs[k % len(msgs)])
        base = np.where(sym >> 1, 128.0 + 50.0, 128.0 - 50.0)
        sigma = [10, 40, 60, 70, 80, 90, 110, 300][k % 8]
        soft = np.clip(base + rng.standard_normal(162) * sigma, 0, 255).astype(np.uint8)
        orc.deinterleave(soft.ctypes.data_as(UP))
        out.append(soft)
    return out


def oracle_fano(v, maxcycles, delta=60):
    orc = po.oracle()
    mettab = ((C.c_int * 256) * 2)()
    orc.oracle_mettab(mettab)
    met, cyc, mx = C.c_uint(), C.c_uint(), C.c_uint()
    data = (C.c_ubyte * 12)()
    s = v.copy()
    rc = orc.fano(C.byref(met), C.byref(cyc), C.byref(mx), data, s.ctypes.data_as(UP), 81, mettab, delta, maxcycles)
    return rc, met.value, cyc.value, mx.value, bytes(data)[:10]


def run_host_many(lib, variant, vecs, maxcycles, delta=60, stop_after=0):
    """All vectors through ONE lane back to back (the lane re-arms between attempts)."""
    n = len(vecs)
    sym = np.ascontiguousarray(np.stack(vecs), dtype=np.uint8)
    rc, met, cyc, mx = np.zeros(n, np.int32), np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.uint32)
    data = np.zeros((n, 12), np.uint8)
    ret = lib.fano_host(variant, sym.ctypes.data, n, delta, maxcycles, stop_after, rc.ctypes.data, met.ctypes.data,
                        cyc.ctypes.data, mx.ctypes.data, data.ctypes.data)
    assert ret == 0, ret
    return [(int(rc[i]), int(met[i]), int(cyc[i]), int(mx[i]), bytes(data[i])[:10]) for i in range(n)]


def run_host(lib, variant, v, maxcycles, delta=60, stop_after=0):
    return run_host_many(lib, variant, [v], maxcycles, delta, stop_after)[0]


@pytest.mark.parametrize("maxcycles,count", [(300, 160), (10000, 16)])
def test_host_build_of_device_fano_matches_oracle(host_fano, maxcycles, count):
    vecs = vectors(count, 42 + maxcycles)
    want = [oracle_fano(v, maxcycles) for v in vecs]
    assert any(x[0] == 0 for x in want) and any(x[0] != 0 for x in want)
    for many in (False, True):                   # one attempt per call / all attempts through one re-arming lane
        run = (lambda var: run_host_many(host_fano, var, vecs, maxcycles)) if many else \
              (lambda var: [run_host(host_fano, var, v, maxcycles) for v in vecs])
        for got, x in zip(run(0), want):         # exact instantiation: every field as fano.c produces it
            assert got[:4] == x[:4], (many, got, x)
            if x[0] == 0:
                assert got[4] == x[4]
        for got, x in zip(run(1), want):         # decode instantiation: rc, cycles; metric and bytes when decoded
            assert (got[0], got[2]) == (x[0], x[2]), (many, got, x)
            if x[0] == 0:
                assert got[1] == x[1] and got[4] == x[4]


def test_host_build_small_delta_and_budget(host_fano):
    """delta <= 10 takes the multi-step threshold path; a budgeted run either gives the full answer or reports 2."""
    vecs = vectors(48, 7)
    for v in vecs:
        x = oracle_fano(v, 200, delta=7)
        assert run_host(host_fano, 0, v, 200, delta=7)[:4] == x[:4]
    fulls = run_host_many(host_fano, 1, vecs, 10000)
    for got, full in zip(run_host_many(host_fano, 1, vecs, 10000, stop_after=2048), fulls):
        if got[0] == 2:
            assert full[2] > 2048
        else:
            assert (got[0], got[2], got[4]) == (full[0], full[2], full[4])
