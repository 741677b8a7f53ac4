"""Pins the oracle (oracle/wspr_oracle.c, our CPU restatement) against the reference: its golden spot lines, its own
unit tests, the committed golden vectors generated from the compiled reference (tools/make_golden.py) and -- where
/root/reference or a prebuilt oracle/_ref is available -- the compiled reference itself on seeded captures."""
import ctypes as C
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po
import helpers as H

FP = C.POINTER(C.c_float)
UP = C.POINTER(C.c_ubyte)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(H.GOLDEN, "golden_decode.json")) as f:
        return json.load(f)["cases"]


def check_against_gold(case, r, io, qo):
    g = case["spots"]
    assert len(r) == len(g)
    for x, y in zip(r, g):
        assert x["message"].decode() == y["message"] and x["call"].decode() == y["call"]
        assert x["loc"].decode() == y["loc"] and x["pwr"].decode() == y["pwr"]
        assert float(x["freq"]).hex() == y["freq"] and float(x["snr"]).hex() == y["snr"]
        assert float(x["dt"]).hex() == y["dt"] and float(x["sync"]).hex() == y["sync"]
        assert float(x["drift"]) == y["drift"] and int(x["jitter"]) == y["jitter"] and int(x["cycles"]) == y["cycles"]
        assert po.spot_line(x) == y["line"]
    assert sha(io) == case["i_sha"] and sha(qo) == case["q_sha"]


def test_golden_fixture_line_oracle(gold):
    # documentation/bug-fix/REPORT.md:202 -- the reference's documented output for its own fixture
    i, q = po.read_iq_file(os.path.join(H.GOLDEN, "refSignalSnr0dB.iq"))
    r, io, qo = po.decode(po.oracle(), i, q)
    assert H.spot_lines(r) == [" -0.07   0.01 144.490550  0    K1JT   FN20 20"]
    check_against_gold(gold["fixture"][0], r, io, qo)


@pytest.mark.parametrize("cfg,count", [(2, 16), (3, 8)])
def test_oracle_reproduces_reference_golden(gold, cfg, count):
    I, Q, _ = H.make_corpus(cfg, count)
    for c in range(count):
        case = gold["config%d" % cfg][c]
        assert sha(I[c]) == case["in_sha"], "corpus generator drifted from the committed golden inputs"
        r, io, qo = po.decode(po.oracle(), I[c], Q[c])
        check_against_gold(case, r, io, qo)


@pytest.mark.parametrize("name,opt", [("quick", dict(quickmode=1)), ("single", dict(npasses=1, subtraction=0))])
def test_oracle_options_variants(gold, name, opt):
    I, Q, _ = H.make_corpus(3, 2)
    for c in range(2):
        r, io, qo = po.decode(po.oracle(), I[c], Q[c], po.default_options(**opt))
        check_against_gold(gold["config3_" + name][c], r, io, qo)


def test_oracle_stage_golden():
    """sync_and_demodulate (three modes, with and without drift) and subtract_signal2 vs compiled-reference outputs."""
    st = np.load(os.path.join(H.GOLDEN, "golden_stages.npz"))
    orc = po.oracle()
    orc.sync_and_demodulate.restype = None
    orc.sync_and_demodulate.argtypes = [FP, FP, C.c_long, UP, FP, C.c_int, C.c_int, C.c_float, C.POINTER(C.c_int), C.c_int,
                                        C.c_int, C.c_int, FP, C.c_int, FP, C.c_int]
    orc.subtract_signal2.restype = None
    orc.subtract_signal2.argtypes = [FP, FP, C.c_long, C.c_float, C.c_int, C.c_float, UP]
    I, Q, _ = H.make_corpus(3, 1)
    i0, q0 = I[0].copy(), Q[0].copy()
    for k in range(3):
        for d in (0, 1):
            f1, sh, drift = st["sync_%d_%d_in" % (k, d)]
            freq, shift, dr, sync = C.c_float(f1), C.c_int(int(sh)), C.c_float(drift), C.c_float(0)
            sym = (C.c_ubyte * 162)()
            a = (i0.ctypes.data_as(FP), q0.ctypes.data_as(FP), 45000, sym, C.byref(freq))
            orc.sync_and_demodulate(*a, 0, 0, 0.0, C.byref(shift), shift.value - 128, shift.value + 128, 8, C.byref(dr), 50, C.byref(sync), 0)
            assert np.array_equal(np.array([freq.value, shift.value, sync.value]), st["sync_%d_%d_m0" % (k, d)])
            orc.sync_and_demodulate(*a, -2, 2, 0.1, C.byref(shift), shift.value, shift.value, 1, C.byref(dr), 50, C.byref(sync), 1)
            assert np.array_equal(np.array([freq.value, shift.value, sync.value]), st["sync_%d_%d_m1" % (k, d)])
            orc.sync_and_demodulate(*a, 0, 0, 0.0, C.byref(shift), shift.value, shift.value, 1, C.byref(dr), 50, C.byref(sync), 2)
            assert sync.value == st["sync_%d_%d_m2" % (k, d)][0]
            assert np.array_equal(np.frombuffer(bytes(sym), np.uint8), st["sync_%d_%d_sym" % (k, d)])
    chan = st["sub_chan"]
    for d in (0, -1):
        f1, sh, drift = st["sub_%d_in" % d]
        ia, qa = i0.copy(), q0.copy()
        orc.subtract_signal2(ia.ctypes.data_as(FP), qa.ctypes.data_as(FP), 45000, C.c_float(f1), int(sh), C.c_float(drift),
                             chan.ctypes.data_as(UP))
        assert np.array_equal(ia, st["sub_%d_i" % d]) and np.array_equal(qa, st["sub_%d_q" % d])


def frontend_golden_streams():
    fe = np.load(os.path.join(H.GOLDEN, "golden_frontend.npz"))
    seed, n_iq = [int(v) for v in fe["seed"]]
    rng = np.random.default_rng(seed)
    raw = rng.integers(0, 256, size=(3, 2 * n_iq), dtype=np.uint8)
    raw[1, : 2 * 6401 * 8] = 0
    raw[2, ::5] = 0
    raw[2, 1::3] = 255
    return raw, n_iq, fe


def test_oracle_frontend_golden():
    raw, n_iq, fe = frontend_golden_streams()
    orc = po.oracle()
    orc.oracle_decimate.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
    for s in range(3):
        io, qo = np.zeros(128, np.float32), np.zeros(128, np.float32)
        n = orc.oracle_decimate(raw[s].ctypes.data, n_iq, io.ctypes.data, qo.ctypes.data, 128)
        assert n == len(fe["i%d" % s]) == n_iq // 6401
        assert np.array_equal(io[:n], fe["i%d" % s]) and np.array_equal(qo[:n], fe["q%d" % s])


def test_fano_and_unpack_known_answers_oracle():
    """tests/test_wsprd.c:168-220 (encode -> 0/255 soft symbols -> fano) and :345-384 (unpk_ -> K1JT FN20 20)."""
    orc = po.oracle()
    sym = H.channel_symbols("K1JT FN20 20", orc)
    soft = np.where(sym >> 1, 255, 0).astype(np.uint8)
    orc.deinterleave(soft.ctypes.data_as(UP))
    mettab = ((C.c_int * 256) * 2)()
    orc.oracle_mettab(mettab)
    metric, cycles, maxnp = C.c_uint(), C.c_uint(), C.c_uint()
    data = (C.c_ubyte * 12)()
    rc = orc.fano(C.byref(metric), C.byref(cycles), C.byref(maxnp), data, soft.ctypes.data_as(UP), 81, mettab, 60, 10000)
    assert rc == 0 and cycles.value == 82
    msg = (C.c_byte * 12)(*[(b - 256 if b > 127 else b) for b in bytes(data)])
    ht, lt = C.create_string_buffer(32768 * 13), C.create_string_buffer(32768 * 5)
    clp, call, loc, pwr, cs = (C.create_string_buffer(n) for n in (23, 13, 7, 3, 13))
    assert orc.unpk_(msg, ht, lt, clp, call, loc, pwr, cs) == 0
    assert (clp.value, call.value, loc.value, pwr.value) == (b"K1JT FN20 20", b"K1JT", b"FN20", b"20")


# ---- against the compiled reference itself (this container; prebuilt oracle/_ref elsewhere) ----------------------
needs_ref = pytest.mark.skipif(not po.ref_available(), reason="no /root/reference and no prebuilt oracle/_ref")


@needs_ref
def test_reference_unit_tests_pass():
    exe = os.path.join(po.HERE, "_ref", "test_wsprd_ref")
    po.ref()
    if not os.path.exists(exe):
        pytest.skip("reference test binary not built")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout[-2000:]


@needs_ref
@pytest.mark.parametrize("cfg,start,count", [(2, 100, 6), (3, 100, 4)])
def test_oracle_equals_compiled_reference(cfg, start, count):
    ref = po.ref()
    I, Q, _ = H.make_corpus(cfg, count, start=start)
    for c in range(count):
        a, ia, qa = po.decode(ref, I[c], Q[c])
        b, ib, qb = po.decode(po.oracle(), I[c], Q[c])
        assert H.results_equal(a, b), H.diff_results(a, b)
        assert np.array_equal(ia, ib) and np.array_equal(qa, qb)


@needs_ref
def test_oracle_frontend_equals_reference_callback():
    if not os.path.exists(os.path.join(po.HERE, "_ref", "libfrontend_ref.so")):
        pytest.skip("reference front end not built")
    rng = np.random.default_rng(11)
    n_iq = 6401 * 40 + 3000
    raw = rng.integers(0, 256, size=2 * n_iq, dtype=np.uint8)
    raw[5000:9000] = 0
    f = po.RefFrontend()
    f.push(raw)
    ir, qr = f.read()
    orc = po.oracle()
    orc.oracle_decimate.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
    io, qo = np.zeros(64, np.float32), np.zeros(64, np.float32)
    n = orc.oracle_decimate(raw.ctypes.data, n_iq, io.ctypes.data, qo.ctypes.data, 64)
    assert n == len(ir) == 40
    assert np.array_equal(io[:n], ir) and np.array_equal(qo[:n], qr)


def test_persistent_hashtable_option_matches_reference():
    """options.usehashtable (reference -H): hashed callsigns of type-3 messages resolve through hashtable.txt carried
    from one call to the next; results and the file after every call are identical between oracle and reference."""
    if not po.ref_available():
        pytest.skip("compiled reference not available")
    opt = po.default_options(usehashtable=1)
    seed = H.HASHTABLE_SEED_FILE
    runs = {}
    for name, lib in (("ref", po.ref()), ("oracle", po.oracle())):
        runs[name] = H.run_hashtable_scenario(lambda i, q: po.decode(lib, i, q, opt, cwd_scratch=False)[0], seed)
    msgs = [[x["message"].decode() for x in r] for r, _ in runs["ref"]]
    assert "<K1JT> FN20AB 20" in msgs[1] and "<PJ4/K1ABC> FK52UD 37" in msgs[1] and "<...> IO90AA 23" in msgs[1], msgs
    for (ra, fa), (rb, fb) in zip(runs["ref"], runs["oracle"]):
        assert H.results_equal(ra, rb), H.diff_results(ra, rb)
        assert fa == fb
    assert "   17 ZZ9ZZZ AA00" in runs["ref"][0][1] and " 5970 W1AW FN31" in runs["ref"][0][1]


def test_persistent_hashtable_option_matches_committed_golden():
    """The same scenario against tests/golden/golden_hashtable.json (generated from the compiled reference by
    tools/make_golden_hashtable.py), so that the pin also holds where /root/reference is absent."""
    with open(os.path.join(H.GOLDEN, "golden_hashtable.json")) as f:
        gold = json.load(f)
    assert gold["seed_file"] == H.HASHTABLE_SEED_FILE
    opt = po.default_options(usehashtable=1)
    runs = H.run_hashtable_scenario(lambda i, q: po.decode(po.oracle(), i, q, opt, cwd_scratch=False)[0], gold["seed_file"])
    assert len(runs) == len(gold["calls"]) == 3
    for (r, txt), g in zip(runs, gold["calls"]):
        assert H.spots_match_golden(r, g["spots"], po.spot_line), ([x["message"] for x in r], [y["message"] for y in g["spots"]])
        assert txt == g["hashtable_txt"]


# ---- the corpora of the GPU suite (tests/test_gpu_parity.py compares the CUDA path with the ORACLE on them): the oracle
# itself against the compiled reference on the same captures and options, so that the chain CUDA == oracle == reference
# has no link that rests on the config-2 / config-3 recipes alone
def _special_captures():
    from rtlsdr_wsprd_b200 import corpus
    caps = []
    for c in range(2):                                                     # weak signals: jitter search, Fano time-outs
        plan = corpus.ten_signal_plan(900 + c, snrs=np.arange(-31.0, -24.0, 1.0))
        caps.append(("weak%d" % c, corpus.make_capture(7, 900 + c, plan, H.channel_symbols)))
    plans = [[dict(message="K1JT FN20 20", f0=30.0, dt0=0.3, snr=-15.0, drift=-3.0)],
             [dict(message="W1AW FN31 37", f0=-72.5, dt0=-0.9, snr=-18.0, drift=2.0)],
             [dict(message="G4JNT IO90 10", f0=108.0, dt0=1.4, snr=-12.0)],
             [dict(message="PJ4/K1ABC 37", f0=10.0, dt0=0.0, snr=-14.0),
              dict(message="<PJ4/K1ABC> FK52UD 37", f0=55.0, dt0=0.1, snr=-19.0)],
             [dict(message="K9AN EN50 33", f0=0.0, dt0=0.0, snr=-5.0), dict(message="K9AN EN50 33", f0=1.5, dt0=0.0, snr=-8.0)],
             []]
    for c, plan in enumerate(plans):
        caps.append(("edge%d" % c, corpus.make_capture(8, c, plan, H.channel_symbols)))
    z = np.zeros(corpus.NSAMP, np.float32)
    t = np.arange(corpus.NSAMP)
    caps.append(("zeros", (z.copy(), z.copy())))
    caps.append(("dc", (z + np.float32(0.25), z - np.float32(0.25))))
    caps.append(("carrier", ((0.5 * np.cos(2 * np.pi * 20.0 * t / 375.0)).astype(np.float32),
                             (0.5 * np.sin(2 * np.pi * 20.0 * t / 375.0)).astype(np.float32))))
    I, Q, _ = H.make_corpus(2, 2, start=40)
    for c in range(2):
        caps.append(("short%d" % c, (np.ascontiguousarray(I[c, :43000]), np.ascontiguousarray(Q[c, :43000]))))
    return caps


@needs_ref
def test_oracle_equals_compiled_reference_on_the_gpu_suite_corpora():
    ref, orc = po.ref(), po.oracle()
    spots = 0
    for name, (i, q) in _special_captures():
        a, ia, qa = po.decode(ref, i, q)
        b, ib, qb = po.decode(orc, i, q)
        assert H.results_equal(a, b), (name, H.diff_results(a, b))
        assert np.array_equal(ia, ib) and np.array_equal(qa, qb), name
        spots += len(a)
    assert spots >= 10


@needs_ref
@pytest.mark.parametrize("opt", [dict(quickmode=1), dict(subtraction=0), dict(npasses=1), dict(npasses=3), dict(npasses=4)])
def test_oracle_equals_compiled_reference_option_variants_on_weak_signals(opt):
    from rtlsdr_wsprd_b200 import corpus
    ref, orc = po.ref(), po.oracle()
    o = po.default_options(**opt)
    for c in range(2):
        plan = corpus.ten_signal_plan(950 + c, snrs=np.arange(-30.0, -19.0, 2.0))
        i, q = corpus.make_capture(9, 950 + c, plan, H.channel_symbols)
        a, ia, qa = po.decode(ref, i, q, o)
        b, ib, qb = po.decode(orc, i, q, o)
        assert H.results_equal(a, b), (opt, c, H.diff_results(a, b))
        assert np.array_equal(ia, ib) and np.array_equal(qa, qb), (opt, c)


@needs_ref
@pytest.mark.parametrize("n", [44700, 44032, 38400, 31400])
def test_short_captures_whose_last_block_reaches_past_the_end(n):
    """blocks = 4 * floor(n / 512) - 1 and block i covers samples 128 i .. 128 i + 511 (wsprd.c:516,536-541): with
    n % 512 < 256 the last blocks read up to 255 samples past n.  The reference's callers hand over full-size buffers with a
    zeroed tail (rtlsdr_wsprd.c:285-288,575-589); pyoracle.decode does the same for both CPU decoders, and the CUDA path
    reads zeros there as well (a context's sample rows are zeroed when it is created and never written past n)."""
    assert n % 512 < 256
    I, Q, _ = H.make_corpus(3, 1, start=7)
    i, q = np.ascontiguousarray(I[0, :n]), np.ascontiguousarray(Q[0, :n])
    a, ia, qa = po.decode(po.ref(), i, q)
    b, ib, qb = po.decode(po.oracle(), i, q)
    assert len(a) >= 5 and H.results_equal(a, b), H.diff_results(a, b)
    assert ia.shape == (n,) and np.array_equal(ia, ib) and np.array_equal(qa, qb)
    again, _, _ = po.decode(po.ref(), i, q)                       # (deterministic: nothing depends on what follows the buffer)
    assert H.results_equal(a, again)


@needs_ref
def test_candidate_loop_breaks_oracle_equals_reference():
    """The two `break`s of the candidate loop (wsprd.c:781-793, helpers.break_captures): nothing is decoded, in both."""
    for name, (i, q) in H.break_captures():
        a, ia, qa = po.decode(po.ref(), i, q)
        b, ib, qb = po.decode(po.oracle(), i, q)
        assert len(a) == 0 and H.results_equal(a, b), (name, H.spot_lines(a), H.spot_lines(b))
        assert np.array_equal(ia, ib) and np.array_equal(qa, qb), name
        # the same capture without the special signal's consequences: with subtraction off nothing is re-encoded, no break (a)
        c, _, _ = po.decode(po.ref(), i, q, po.default_options(subtraction=0))
        d, _, _ = po.decode(po.oracle(), i, q, po.default_options(subtraction=0))
        assert H.results_equal(c, d), name
        if name == "three-character callsign":
            assert len(c) == 3, H.spot_lines(c)
