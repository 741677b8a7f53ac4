"""CPU-side checks of libwsprd_b200.so: it loads, exports everything include/wspr_b200.h declares, its host codec
entry points (the same inline code the kernels run) agree with the oracle and pass the reference's own unit tests,
and compute entry points fail loudly -- never fall back -- when no CUDA device is present."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po
import rtlsdr_wsprd_b200 as w
import helpers as H

UP = C.POINTER(C.c_ubyte)


def cuda_device_count():
    try:
        rt = C.CDLL("libcudart.so.12")
    except OSError:
        return 0
    n = C.c_int(0)
    return n.value if rt.cudaGetDeviceCount(C.byref(n)) == 0 else 0


def declared_functions():
    text = open(os.path.join(H.ROOT, "include", "wspr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", text)
    return sorted(set(names)), re.findall(r"extern\s+unsigned\s+char\s+(\w+)\s*\[", text)


def test_library_exports_every_declared_symbol():
    lib = w.library()
    funcs, data = declared_functions()
    assert len(funcs) >= 40, funcs
    missing = [n for n in funcs + data if not hasattr(lib, n)]
    assert not missing, missing


def test_struct_layouts_match_reference_abi():
    assert C.sizeof(w.DecoderOptions) == 40 and C.sizeof(w.DecoderResults) == 80
    assert w.DecoderOptions.quickmode.offset == 24 and w.DecoderOptions.subtraction.offset == 36
    assert w.DecoderResults.message.offset == 28 and w.DecoderResults.cycles.offset == 76
    assert w.RESULT_DTYPE.itemsize == 80 and w.CAND_DTYPE.itemsize == 20


def test_no_cpu_fallback_without_device():
    if cuda_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(w.WsprCudaError):
        w.BatchDecoder(1)
    with pytest.raises(w.WsprCudaError):
        w.decode_batch(np.zeros((1, 45000), np.float32), np.zeros((1, 45000), np.float32))
    with pytest.raises(w.WsprCudaError):
        w.decimate_batch(np.zeros((1, 2 * 6401), np.uint8))
    with pytest.raises(w.WsprCudaError):
        w.wspr_decode(np.zeros(45000, np.float32), np.zeros(45000, np.float32))


def _channel_symbols(lib, msg):
    sym = (C.c_ubyte * 162)()
    ht, lt = C.create_string_buffer(32768 * 13), C.create_string_buffer(32768 * 5)
    ok = lib.get_wspr_channel_symbols(C.create_string_buffer(msg.encode(), 32), ht, lt, sym)
    return ok, bytes(sym), ht.raw, lt.raw


MESSAGES = ["K1JT FN20 20", "VA2GKA FN35 37", "W1AW FN31 0", "G4JNT IO90 60", "K9AN EN50 33", "PJ4/K1ABC 37", "K1ABC/7 30",
            "K1ABC/12 23", "<K1ABC> FN42AX 10", "<PJ4/K1ABC> FK52UD 3", "Q1Q AA00 7", "3DA0XY KG53 13", "bogus", "TOOLONGCALL FN20 10",
            "K1JT FN20 21", "A/B 10", "W1AW/P 17", "ZL/W1AW 27", "F/W1AW 50", "ABC/W1AW 43"]


@pytest.mark.parametrize("msg", MESSAGES)
def test_channel_symbols_match_oracle(msg):
    a = _channel_symbols(w.library(), msg)
    b = _channel_symbols(po.oracle(), msg)
    assert a[0] == b[0]
    if a[0]:
        assert a[1] == b[1]
    assert a[2] == b[2], "hash table side effects differ"


def _unpk(lib, data11, ht=None, lt=None):
    msg = (C.c_byte * 12)(*[(b - 256 if b > 127 else b) for b in data11], 0)
    ht = ht or C.create_string_buffer(32768 * 13)
    lt = lt or C.create_string_buffer(32768 * 5)
    bufs = [C.create_string_buffer(n) for n in (23, 13, 7, 3, 13)]
    rc = lib.unpk_(msg, ht, lt, *bufs)
    return (rc,) + tuple(b.value for b in bufs) + (ht.raw, lt.raw)


def test_unpk_matches_oracle_on_random_and_real_messages():
    ours, orc = w.library(), po.oracle()
    rng = np.random.default_rng(5)
    cases = [bytes(rng.integers(0, 256, 11, dtype=np.uint8)) for _ in range(3000)]
    # real messages of all three types: re-encode what get_wspr_channel_symbols packs (bits 2*sym>>1 are the code bits)
    for m in MESSAGES:
        ok, sym, _, _ = _channel_symbols(orc, m)
        if ok:
            bits = (np.frombuffer(sym, np.uint8) >> 1).astype(np.uint8)
            soft = (bits * 255).astype(np.uint8)
            orc.deinterleave(soft.ctypes.data_as(UP))
            mettab = ((C.c_int * 256) * 2)()
            orc.oracle_mettab(mettab)
            met, cyc, mx = C.c_uint(), C.c_uint(), C.c_uint()
            data = (C.c_ubyte * 12)()
            assert orc.fano(C.byref(met), C.byref(cyc), C.byref(mx), data, soft.ctypes.data_as(UP), 81, mettab, 60, 10000) == 0
            cases.append(bytes(data)[:11])
    # persistent tables across calls (type-3 lookups see earlier type-1/2 inserts)
    hts = [(C.create_string_buffer(32768 * 13), C.create_string_buffer(32768 * 5)) for _ in range(2)]
    for d in cases:
        a = _unpk(ours, d, *hts[0])
        b = _unpk(orc, d, *hts[1])
        assert a[:6] == b[:6], (d.hex(), a[:6], b[:6])
    assert hts[0][0].raw == hts[1][0].raw and hts[0][1].raw == hts[1][1].raw


def test_fano_host_matches_oracle_including_timeouts():
    ours, orc = w.library(), po.oracle()
    mettab = ((C.c_int * 256) * 2)()
    orc.oracle_mettab(mettab)
    rng = np.random.default_rng(9)
    sym = H.channel_symbols("K1JT FN20 20")
    base = np.where(sym >> 1, 200.0, 56.0)
    for trial, sigma in enumerate([0, 30, 60, 75, 85, 95, 400]):
        soft = np.clip(base + rng.standard_normal(162) * sigma, 0, 255).astype(np.uint8)
        orc.deinterleave(soft.ctypes.data_as(UP))
        res = []
        for lib in (ours, orc):
            met, cyc, mx = C.c_uint(), C.c_uint(), C.c_uint()
            data = (C.c_ubyte * 12)()
            s = soft.copy()
            rc = lib.fano(C.byref(met), C.byref(cyc), C.byref(mx), data, s.ctypes.data_as(UP), 81, mettab, 60, 2000)
            # on a timeout the bytes of tree nodes never reached are unspecified (malloc garbage in fano.c:107)
            nbytes = 10 if rc == 0 else (mx.value + 1) // 8
            res.append((rc, met.value, cyc.value, mx.value, bytes(data)[:nbytes]))
        assert res[0] == res[1], (sigma, res)


def test_small_codec_functions_match_oracle():
    ours, orc = w.library(), po.oracle()
    for lib in (ours, orc):
        lib.nhash.restype = C.c_uint32
        lib.nhash.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32]
        lib.pack_call.restype = C.c_ulong
        lib.pack_call.argtypes = [C.c_char_p]
        lib.get_callsign_character_code.restype = C.c_char
        lib.get_callsign_character_code.argtypes = [C.c_char]
        lib.get_locator_character_code.restype = C.c_char
        lib.get_locator_character_code.argtypes = [C.c_char]
    rng = np.random.default_rng(3)
    for _ in range(500):
        s = bytes(rng.integers(32, 127, int(rng.integers(1, 13)), dtype=np.uint8))
        assert ours.nhash(s, len(s), 146) == orc.nhash(s, len(s), 146)
    for call in [b"K1JT", b"VA2GKA", b"W1AW", b"3DA0XY", b"Q1Q", b"A", b"AB", b"TOOLONG7", b"K1", b"1A2B3C"]:
        assert ours.pack_call(call) == orc.pack_call(call), call
    for ch in range(256):
        c = bytes([ch])
        assert ours.get_callsign_character_code(c) == orc.get_callsign_character_code(c)
        assert ours.get_locator_character_code(c) == orc.get_locator_character_code(c)
    for n in list(rng.integers(0, 2 ** 28, 300)) + [262177559, 262177560]:
        a, b = C.create_string_buffer(13), C.create_string_buffer(13)
        assert ours.unpackcall(int(n), a) == orc.unpackcall(int(n), b) and a.raw == b.raw
    for n in list(rng.integers(0, 2 ** 22, 300)):
        a, b = C.create_string_buffer(b"....", 5), C.create_string_buffer(b"....", 5)
        assert ours.unpackgrid(int(n), a) == orc.unpackgrid(int(n), b) and a.raw == b.raw
    for n in list(rng.integers(0, 61000, 300)):
        a, b = C.create_string_buffer(b"K1ABC", 13), C.create_string_buffer(b"K1ABC", 13)
        assert ours.unpackpfx(int(n), a) == orc.unpackpfx(int(n), b) and a.value == b.value
    x = np.arange(162, dtype=np.uint8)
    a, b = x.copy(), x.copy()
    ours.interleave(a.ctypes.data_as(UP)); orc.interleave(b.ctypes.data_as(UP))
    assert np.array_equal(a, b)
    ours.deinterleave(a.ctypes.data_as(UP))
    assert np.array_equal(a, x)
    assert bytes((C.c_ubyte * 256).in_dll(ours, "Partab")) == bytes(bin(i).count("1") & 1 for i in range(256))


def test_glibc_float_replicas_match_host_libm():
    """wspr_math.cuh: the kernels' sinf/cosf/log10f replicas vs this host's glibc, sampled (exhaustive in tools/)."""
    lib, libm = w.library(), C.CDLL("libm.so.6")
    for f in (lib.wspr_test_sinf, lib.wspr_test_cosf, lib.wspr_test_log10f, libm.sinf, libm.cosf, libm.log10f):
        f.restype, f.argtypes = C.c_float, [C.c_float]
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.uniform(-2, 2, 20000), rng.uniform(-8e4, 8e4, 20000), rng.uniform(-130, 130, 5000),
                         [0.0, 1e-5, -1e-5, 0.785398, 120.0, 119.99, 1e9]]).astype(np.float32)
    for x in xs:
        assert lib.wspr_test_sinf(x) == libm.sinf(x), x
        assert lib.wspr_test_cosf(x) == libm.cosf(x), x
    for x in np.concatenate([rng.uniform(1e-6, 50, 20000), 10 ** rng.uniform(-6, 9, 20000)]).astype(np.float32):
        assert lib.wspr_test_log10f(x) == libm.log10f(x), x


@pytest.mark.skipif(not os.path.isdir("/root/reference/tests"), reason="reference sources not mounted")
def test_reference_unit_tests_link_against_our_library(tmp_path):
    """tests/test_wsprd.c (18 tests) compiled where it lies and linked against libwsprd_b200.so instead of wsprd/*.o."""
    exe = str(tmp_path / "test_wsprd_b200")
    libdir = os.path.dirname(w.library_path())
    subprocess.run(["gcc", "-O2", "-std=gnu17", "-w", "-I/root/reference/tests", "-o", exe, "/root/reference/tests/test_wsprd.c",
                    "-L" + libdir, "-lwsprd_b200", "-Wl,-rpath," + libdir, "-lm"], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-1000:]
    assert "18" in out.stdout


def test_packed_products_and_sums_are_not_contracted():
    """The correlation / low-pass kernels issue their exact-order multiplies and adds as FFMA2 pairs (a*b + -0.0, then
    a*1.0 + c).  ptxas contracts such a pair into one fused multiply-add whenever it can see the two constants (it does
    so even under -fmad=false), which would change roundings, so the constants are run-time kernel arguments.  Check the
    SASS of the built library: each kernel must hold TWO FFMA2 per packed multiply-accumulate of its unrolled inner loop
    (a contracted build would show half as many).  The GPU parity tests are the functional check of the same thing."""
    lib = w.library_path()
    try:
        sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    except (OSError, subprocess.CalledProcessError):
        pytest.skip("cuobjdump not available")
    count, lds64, lds128, cur = {}, {}, {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
        elif cur:
            for key, tab in (("FFMA2", count), ("LDS.64", lds64), ("LDS.128", lds128)):
                if key in line:
                    tab[cur] = tab.get(cur, 0) + 1
    # table path: packed multiply-accumulates per sample (4 tones x {i,q} x 2 products / 2 lanes = 8; low-pass: 4 outputs) x
    # samples (taps) in the unrolled loop body x 2 instructions each.  Drifting candidates advance their phasors in
    # registers: 28 FFMA2 per sample (16 for the sums + 8 products and 4 sums of the recurrence) x the samples of the
    # unrolled body (nvcc 12.9 unrolls the 4-sample bodies twice).
    expected = {"k_sync_freqsE": 8 * 8 * 2 + 8 * 28, "k_sync_freqs_sharedE": 8 * 8 * 2,
                "k_jitter_softE": 8 * 8 * 2 + 8 * 28, "k_sync_genericE": 8 * 8 * 2 + 16 * 28, "k_sub_lpfI": 4 * 8 * 2}
    for name, n in expected.items():                     # (mangled names: the suffix keeps k_sync_freqs and ..._shared apart)
        got = sum(v for k, v in count.items() if name in k)
        assert got == n, (name, got, n)
    # K4 reads its samples from shared memory, one 8-byte load per sample in either path plus two 16-byte table loads per
    # sample in the table path, which makes the check independent of how far the compiler unrolls or peels the loops:
    # 16 FFMA2 per table-path sample, 28 per drift-path sample.
    k4 = [k for k in count if "k_sync_lagsE" in k]
    assert len(k4) == 1
    table_samples = lds128[k4[0]] // 2
    drift_samples = lds64[k4[0]] - table_samples
    assert table_samples > 0 and drift_samples > 0
    assert count[k4[0]] == 16 * table_samples + 28 * drift_samples, (count[k4[0]], table_samples, drift_samples)
