"""The device Fano loop (rtlsdr_wsprd_b200/csrc/wspr_fano.cuh) compiled for the HOST (tools/fano_host_check.cpp: same
template source, one lane, scratch in ordinary memory) against fano() of the oracle on random soft-symbol vectors from
clean to hopeless, time-outs included.  Covers the logic of all four instantiations -- exact / decode (time-out test every
256 trips, maxnp not tracked) x plain / pipelined (records fetched a trip ahead) -- without a GPU; the GPU suite repeats the
comparison for the instantiations the library runs."""
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
    lib.fano_host.argtypes = [C.c_int, UP, C.c_int, C.c_uint, C.c_uint, C.POINTER(C.c_uint), C.POINTER(C.c_uint),
                              C.POINTER(C.c_uint), UP]
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


def run_host(lib, variant, v, maxcycles, delta=60, stop_after=0):
    met, cyc, mx = C.c_uint(), C.c_uint(), C.c_uint()
    data = (C.c_ubyte * 12)()
    s = v.copy()
    rc = lib.fano_host(variant, s.ctypes.data_as(UP), delta, maxcycles, stop_after, C.byref(met), C.byref(cyc), C.byref(mx), data)
    return rc, met.value, cyc.value, mx.value, bytes(data)[:10]


@pytest.mark.parametrize("maxcycles,count", [(300, 160), (10000, 16)])
def test_host_build_of_device_fano_matches_oracle(host_fano, maxcycles, count):
    vecs = vectors(count, 42 + maxcycles)
    want = [oracle_fano(v, maxcycles) for v in vecs]
    assert any(x[0] == 0 for x in want) and any(x[0] != 0 for x in want)
    for variant in (0, 2):                       # exact instantiations: every field as fano.c produces it
        for v, x in zip(vecs, want):
            got = run_host(host_fano, variant, v, maxcycles)
            assert got[:4] == x[:4], (variant, got, x)
            if x[0] == 0:
                assert got[4] == x[4]
    for variant in (1, 3):                       # decode instantiations: rc, cycles; metric and bytes when decoded
        for v, x in zip(vecs, want):
            got = run_host(host_fano, variant, v, maxcycles)
            assert (got[0], got[2]) == (x[0], x[2]), (variant, got, x)
            if x[0] == 0:
                assert got[1] == x[1] and got[4] == x[4]


def test_host_build_small_delta_and_budget(host_fano):
    """delta <= 10 takes the multi-step threshold path; a budgeted run either gives the full answer or reports 2."""
    vecs = vectors(48, 7)
    for v in vecs:
        x = oracle_fano(v, 200, delta=7)
        for variant in (0, 2):
            assert run_host(host_fano, variant, v, 200, delta=7)[:4] == x[:4]
    for v in vecs:
        full = run_host(host_fano, 1, v, 10000)
        for variant in (1, 3):
            got = run_host(host_fano, variant, v, 10000, stop_after=2048)
            if got[0] == 2:
                assert full[2] > 2048
            else:
                assert (got[0], got[2], got[4]) == (full[0], full[2], full[4])
