"""GPU parity tests: the CUDA path (through the C ABI of libwsprd_b200.so) against the oracle and the committed golden
vectors.  Bar: every field of every spot identical (the contract in BASELINE.json demands callsign/locator/dBm and spot
count bit-for-bit and soft symbols within 1e-5; the implementation is exact-order float, so these tests pin bit
equality, tolerance 0), post-subtraction samples bit-identical, decimator outputs bit-identical."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import pyoracle as po
import rtlsdr_wsprd_b200 as w
from rtlsdr_wsprd_b200 import corpus
import helpers as H

pytestmark = pytest.mark.gpu
FP = C.POINTER(C.c_float)
UP = C.POINTER(C.c_ubyte)
SOFT_SYMBOL_TOLERANCE = 0.0     # allowed |difference| of soft symbols (BASELINE allows 1e-5 * symfac)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(H.GOLDEN, "golden_decode.json")) as f:
        return json.load(f)["cases"]


def assert_matches_gold(case, r, io=None, qo=None):
    g = case["spots"]
    assert len(r) == len(g), (len(r), len(g), H.spot_lines(r), [y["line"] for y in g])
    for x, y in zip(r, g):
        assert (x["message"].decode(), x["call"].decode(), x["loc"].decode(), x["pwr"].decode()) == \
            (y["message"], y["call"], y["loc"], y["pwr"])
        assert w.spot_line(x) == y["line"]
        assert float(x["freq"]).hex() == y["freq"] and float(x["snr"]).hex() == y["snr"]
        assert float(x["dt"]).hex() == y["dt"] and float(x["sync"]).hex() == y["sync"]
        assert float(x["drift"]) == y["drift"] and int(x["jitter"]) == y["jitter"] and int(x["cycles"]) == y["cycles"]
    if io is not None:
        assert sha(io) == case["i_sha"] and sha(qo) == case["q_sha"]


def gpu_decode(I, Q, options=None, samples=True):
    with w.BatchDecoder(len(I), I.shape[1]) as d:
        d.upload(I, Q)
        d.decode(options)
        return d.download(samples=samples)


def assert_batch_equals_oracle(I, Q, **opt):
    spots, n, Io, Qo = gpu_decode(I, Q, w.default_options(**opt))
    total = 0
    for c in range(len(I)):
        a, ia, qa = po.decode(po.oracle(), I[c], Q[c], po.default_options(**opt))
        b = spots[c, : n[c]]
        assert H.results_equal(a, b), (c, H.diff_results(a, b))
        assert np.array_equal(ia, Io[c]) and np.array_equal(qa, Qo[c]), c
        total += len(a)
    return total


def test_loaded_library_is_the_in_tree_cuda_build():
    assert os.path.samefile(w.library_path(), os.path.join(H.ROOT, "rtlsdr_wsprd_b200", "libwsprd_b200.so"))
    before = w.kernel_launches()
    gpu_decode(*H.make_corpus(2, 1)[:2])
    assert w.kernel_launches() > before


def test_reference_fixture_through_reference_abi(gold):
    """config 1: signals/refSignalSnr0dB.iq through wspr_decode(); idat/qdat mutated like the reference does."""
    i, q = w.read_iq_file(os.path.join(H.GOLDEN, "refSignalSnr0dB.iq"))
    r = w.wspr_decode(i, q, 45000, w.default_options())
    assert H.spot_lines(r) == [" -0.07   0.01 144.490550  0    K1JT   FN20 20"]   # documentation/bug-fix/REPORT.md:202
    assert_matches_gold(gold["fixture"][0], r, i, q)


@pytest.mark.parametrize("cfg", [2, 3])
def test_golden_corpus(gold, cfg):
    cases = gold["config%d" % cfg]
    I, Q, _ = H.make_corpus(cfg, len(cases))
    spots, n, Io, Qo = gpu_decode(I, Q)
    for c, case in enumerate(cases):
        assert_matches_gold(case, spots[c, : n[c]], Io[c], Qo[c])


@pytest.mark.parametrize("name,opt", [("quick", dict(quickmode=1)), ("single", dict(npasses=1, subtraction=0))])
def test_golden_option_variants(gold, name, opt):
    I, Q, _ = H.make_corpus(3, 4)
    spots, n, Io, Qo = gpu_decode(I, Q, w.default_options(**opt))
    for c in range(4):
        assert_matches_gold(gold["config3_" + name][c], spots[c, : n[c]], Io[c], Qo[c])


def test_config2_against_oracle():
    I, Q, _ = H.make_corpus(2, 48, start=500)
    assert assert_batch_equals_oracle(I, Q) >= 40


def test_config3_against_oracle():
    I, Q, _ = H.make_corpus(3, 16, start=500)
    assert assert_batch_equals_oracle(I, Q) >= 140


def test_weak_signals_exercise_jitter_search_and_fano_timeouts():
    """SNR -31..-25 dB: most candidates fail at jitter 0, go through the 42 further attempts and many Fano timeouts."""
    n = 4
    I = np.zeros((n, corpus.NSAMP), np.float32)
    Q = np.zeros_like(I)
    for c in range(n):
        plan = corpus.ten_signal_plan(900 + c, snrs=np.arange(-31.0, -24.0, 1.0))
        I[c], Q[c] = corpus.make_capture(7, 900 + c, plan, H.channel_symbols)
    spots, n_res, Io, Qo = gpu_decode(I, Q)
    jit = 0
    for c in range(n):
        a, ia, qa = po.decode(po.oracle(), I[c], Q[c])
        assert H.results_equal(a, spots[c, : n_res[c]]), (c, H.diff_results(a, spots[c, : n_res[c]]))
        assert np.array_equal(ia, Io[c]) and np.array_equal(qa, Qo[c])
        jit += int(np.count_nonzero(a["jitter"]))
    assert jit > 0, "corpus did not exercise the jitter path"


def test_drifting_and_edge_timed_signals():
    I = np.zeros((6, corpus.NSAMP), np.float32)
    Q = np.zeros_like(I)
    plans = [[dict(message="K1JT FN20 20", f0=30.0, dt0=0.3, snr=-15.0, drift=-3.0)],
             [dict(message="W1AW FN31 37", f0=-72.5, dt0=-0.9, snr=-18.0, drift=2.0)],
             [dict(message="G4JNT IO90 10", f0=108.0, dt0=1.4, snr=-12.0)],              # band edge + late start
             [dict(message="PJ4/K1ABC 37", f0=10.0, dt0=0.0, snr=-14.0),                 # type 2, then type 3 that
              dict(message="<PJ4/K1ABC> FK52UD 37", f0=55.0, dt0=0.1, snr=-19.0)],       # needs the hash of the first
             [dict(message="K9AN EN50 33", f0=0.0, dt0=0.0, snr=-5.0), dict(message="K9AN EN50 33", f0=1.5, dt0=0.0, snr=-8.0)],
             []]                                                                           # noise only
    for c, plan in enumerate(plans):
        I[c], Q[c] = corpus.make_capture(8, c, plan, H.channel_symbols)
    assert assert_batch_equals_oracle(I, Q) >= 5


def test_degenerate_inputs():
    z = np.zeros((3, corpus.NSAMP), np.float32)
    I, Q = z.copy(), z.copy()
    I[1], Q[1] = 0.25, -0.25                                      # DC
    t = np.arange(corpus.NSAMP)
    I[2] = (0.5 * np.cos(2 * np.pi * 20.0 * t / 375.0)).astype(np.float32)   # noise-free carrier
    Q[2] = (0.5 * np.sin(2 * np.pi * 20.0 * t / 375.0)).astype(np.float32)
    assert_batch_equals_oracle(I[1:], Q[1:])
    spots, n, Io, Qo = gpu_decode(I[:1], Q[:1])                  # all-zero capture: 0/0 everywhere, no spots
    a, _, _ = po.decode(po.oracle(), I[0], Q[0])
    assert n[0] == len(a) == 0
    with w.BatchDecoder(4) as d:                                  # empty batch
        d.upload(np.zeros((0, corpus.NSAMP), np.float32), np.zeros((0, corpus.NSAMP), np.float32))
        d.decode()
        spots, n = d.download()
        assert spots.shape[0] == 0 and n.shape[0] == 0


@pytest.mark.parametrize("n", [43000, 44700, 36964])
def test_short_capture_length(n):
    """Captures shorter than 45000 samples.  With n % 512 < 256 (44700, 36964) the last spectrogram blocks reach up to 255
    samples past the end (wsprd.c:516,536-541): zeros in the reference's full-size buffers, in pyoracle's padded ones and in
    the context's rows, whose stride leaves room for them (ctx_init) -- in a batch the over-read must not run into the next
    capture's samples."""
    I, Q, _ = H.make_corpus(3 if n < 40000 else 2, 3, start=40)
    I, Q = np.ascontiguousarray(I[:, :n]), np.ascontiguousarray(Q[:, :n])
    assert assert_batch_equals_oracle(I, Q) >= 3


def test_results_independent_of_batch_composition():
    """1024-capture batch = 32 distinct captures, each 32 times in shuffled order: every replica decodes identically
    (and equal to the oracle on the distinct captures) -- batch-scale indexing at BASELINE config-2 size."""
    I0, Q0, _ = H.make_corpus(3, 8, start=300)
    I1, Q1, _ = H.make_corpus(2, 24, start=300)
    I0, Q0 = np.concatenate([I0, I1]), np.concatenate([Q0, Q1])
    order = np.random.default_rng(0).permutation(np.repeat(np.arange(32), 32))
    spots, n, Io, Qo = gpu_decode(I0[order], Q0[order])
    ref = {}
    for k in range(32):
        a, ia, qa = po.decode(po.oracle(), I0[k], Q0[k])
        ref[k] = (a, sha(ia), sha(qa))
    for pos, k in enumerate(order):
        a, si, sq = ref[int(k)]
        assert H.results_equal(a, spots[pos, : n[pos]]), (pos, k)
        assert sha(Io[pos]) == si and sha(Qo[pos]) == sq


def test_one_shot_batch_entry_and_normalise():
    I, Q, _ = H.make_corpus(2, 5, start=60)
    res = w.decode_batch(I, Q)
    for c in range(5):
        a, _, _ = po.decode(po.oracle(), I[c], Q[c])
        assert H.results_equal(a, res[c])
    # hand-off normalisation (rtlsdr_wsprd.c:291-305) on the device == oracle_normalise
    raw_i, raw_q = (I * 3.7).astype(np.float32), (Q * 3.7).astype(np.float32)
    with w.BatchDecoder(5) as d:
        d.upload(raw_i, raw_q)
        d.normalise()
        d.decode(w.default_options(npasses=1, subtraction=0))
        _, _, In, Qn = d.download(samples=True)
    orc = po.oracle()
    for c in range(5):
        a, b = raw_i[c].copy(), raw_q[c].copy()
        orc.oracle_normalise(a.ctypes.data_as(FP), b.ctypes.data_as(FP), a.shape[0])
        assert np.array_equal(a, In[c]) and np.array_equal(b, Qn[c])


def test_stage_spectrogram_and_candidates_bit_exact():
    orc = po.oracle()
    orc.oracle_spectrogram.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    orc.oracle_candidates.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    orc.oracle_candidates.restype = C.c_int
    I, Q, _ = H.make_corpus(3, 6, start=700)
    with w.BatchDecoder(6) as d:
        d.upload(I, Q)
        ps = d.spectrogram()
        for maxdrift in (4, 0):
            cands, npk, sm = d.candidates(maxdrift, want_smspec=True)
            for c in range(6):
                pso = np.zeros((512, ps.shape[2]), np.float32)
                orc.oracle_spectrogram(I[c].ctypes.data, Q[c].ctypes.data, I.shape[1], pso.ctypes.data)
                assert np.array_equal(ps[c], pso)
                co, smo = np.zeros(200, w.CAND_DTYPE), np.zeros(411, np.float32)
                nk = orc.oracle_candidates(pso.ctypes.data, ps.shape[2], maxdrift, co.ctypes.data, smo.ctypes.data)
                assert nk == npk[c] and cands[c, :nk].tobytes() == co[:nk].tobytes()
                assert np.array_equal(sm[c], smo)


def test_sync_and_demodulate_abi_against_reference_golden():
    """soft symbols: tolerance SOFT_SYMBOL_TOLERANCE (= 0, BASELINE allows 1e-5); sync/freq/shift identical."""
    st = np.load(os.path.join(H.GOLDEN, "golden_stages.npz"))
    lib = w.library()
    I, Q, _ = H.make_corpus(3, 1)
    i0, q0 = I[0].copy(), Q[0].copy()
    for k in range(3):
        for dft in (0, 1):
            f1, sh, drift = st["sync_%d_%d_in" % (k, dft)]
            freq, shift, dr, sync = C.c_float(f1), C.c_int(int(sh)), C.c_float(drift), C.c_float(0)
            sym = (C.c_ubyte * 162)()
            a = (i0.ctypes.data_as(FP), q0.ctypes.data_as(FP), 45000, sym, C.byref(freq))
            lib.sync_and_demodulate(*a, 0, 0, 0.0, C.byref(shift), shift.value - 128, shift.value + 128, 8, C.byref(dr), 50, C.byref(sync), 0)
            assert np.array_equal(np.array([freq.value, shift.value, sync.value]), st["sync_%d_%d_m0" % (k, dft)])
            lib.sync_and_demodulate(*a, -2, 2, 0.1, C.byref(shift), shift.value, shift.value, 1, C.byref(dr), 50, C.byref(sync), 1)
            assert np.array_equal(np.array([freq.value, shift.value, sync.value]), st["sync_%d_%d_m1" % (k, dft)])
            lib.sync_and_demodulate(*a, 0, 0, 0.0, C.byref(shift), shift.value, shift.value, 1, C.byref(dr), 50, C.byref(sync), 2)
            assert sync.value == st["sync_%d_%d_m2" % (k, dft)][0]
            got = np.frombuffer(bytes(sym), np.uint8).astype(np.float64)
            assert np.max(np.abs(got - st["sync_%d_%d_sym" % (k, dft)])) <= SOFT_SYMBOL_TOLERANCE


def test_subtract_signal2_abi_against_reference_golden():
    st = np.load(os.path.join(H.GOLDEN, "golden_stages.npz"))
    lib = w.library()
    I, Q, _ = H.make_corpus(3, 1)
    chan = st["sub_chan"]
    for dft in (0, -1):
        f1, sh, drift = st["sub_%d_in" % dft]
        ia, qa = I[0].copy(), Q[0].copy()
        lib.subtract_signal2(ia.ctypes.data_as(FP), qa.ctypes.data_as(FP), 45000, C.c_float(f1), int(sh), C.c_float(drift),
                             chan.ctypes.data_as(UP))
        assert np.array_equal(ia, st["sub_%d_i" % dft]) and np.array_equal(qa, st["sub_%d_q" % dft])


def test_pipelined_contexts_do_not_interfere():
    """Three batches in flight on three contexts driven by three host threads (what bench.py does): every batch still
    decodes exactly like the oracle."""
    batches = [H.make_corpus(3, 6, start=1200 + 10 * k)[:2] for k in range(3)] + [H.make_corpus(2, 12, start=1300)[:2]]
    with w.PipelinedDecoder(3, 12) as pipe:
        futs = [pipe.decode_async(I, Q) for I, Q in batches for _ in range(2)]
        outs = [f.result() for f in futs]
    for k, (I, Q) in enumerate(batches):
        for rep in range(2):
            spots, n = outs[2 * k + rep]
            for c in range(len(I)):
                a, _, _ = po.decode(po.oracle(), I[c], Q[c])
                assert H.results_equal(a, spots[c, : n[c]]), (k, rep, c, H.diff_results(a, spots[c, : n[c]]))


@pytest.mark.parametrize("opt", [dict(quickmode=1), dict(subtraction=0), dict(npasses=1), dict(npasses=3)])
def test_option_variants_on_weak_signals(opt):
    """Options against the oracle on captures that park candidates (long Fano runs, jitter search): quick mode has no
    jitter loop, subtraction off makes pass 0 order-free, a third pass changes drift range and minsync2 (wsprd.c:524-531)."""
    n = 3
    I = np.zeros((n, corpus.NSAMP), np.float32)
    Q = np.zeros_like(I)
    for c in range(n):
        plan = corpus.ten_signal_plan(950 + c, snrs=np.arange(-30.0, -19.0, 2.0))
        I[c], Q[c] = corpus.make_capture(9, 950 + c, plan, H.channel_symbols)
    assert_batch_equals_oracle(I, Q, **opt)


def test_persistent_hashtable_option():
    """options.usehashtable (reference -H, wsprd.c:481-494,842-852) through the reference entry point: spots and the
    hashtable.txt left in the CWD after every call identical to the oracle's."""
    opt_o, opt_g = po.default_options(usehashtable=1), w.default_options(usehashtable=1)
    seed = H.HASHTABLE_SEED_FILE
    want = H.run_hashtable_scenario(lambda i, q: po.decode(po.oracle(), i, q, opt_o, cwd_scratch=False)[0], seed)
    got = H.run_hashtable_scenario(lambda i, q: w.wspr_decode(i, q, options=opt_g), seed)
    assert any(b"<K1JT>" in x["message"] for x in got[1][0]) and any(b"<...>" in x["message"] for x in got[1][0])
    for (ra, fa), (rb, fb) in zip(want, got):
        assert H.results_equal(ra, rb), H.diff_results(ra, rb)
        assert fa == fb, (fa, fb)
    # and identical to what the compiled reference produced (tests/golden/golden_hashtable.json)
    with open(os.path.join(H.GOLDEN, "golden_hashtable.json")) as f:
        gold = json.load(f)
    assert gold["seed_file"] == seed and len(gold["calls"]) == len(got)
    for (rb, fb), g in zip(got, gold["calls"]):
        assert H.spots_match_golden(rb, g["spots"], w.spot_line) and fb == g["hashtable_txt"]
    # without the option nothing is read or written and hashed calls stay unresolved
    i, q = H.hashtable_scenario()[1]
    r = w.wspr_decode(i.copy(), q.copy())
    assert not any(b"<K1JT>" in x["message"] for x in r) and any(b"<...>" in x["message"] for x in r)


def test_fano_kernel_against_oracle_random_vectors():
    """K5 alone: random soft-symbol vectors from clean to hopeless, 32 attempts per warp and one per warp, against
    fano() of the oracle (return code, metric, cycle count, deepest node, decoded bytes); timeouts at a reduced maxcycles
    keep the CPU side fast, plus a few at the reference's full 10000."""
    orc = po.oracle()
    mettab = ((C.c_int * 256) * 2)()
    orc.oracle_mettab(mettab)
    rng = np.random.default_rng(42)
    msgs = ["K1JT FN20 20", "VA2GKA FN35 37", "G4JNT IO90 60", "<K1ABC> FN42AX 10", "PJ4/K1ABC 37"]
    vecs = []
    for k in range(192):
        sym = H.channel_symbols(msgs[k % len(msgs)])
        base = np.where(sym >> 1, 128.0 + 50.0, 128.0 - 50.0)
        sigma = [10, 40, 60, 70, 80, 90, 110, 300][k % 8]
        soft = np.clip(base + rng.standard_normal(162) * sigma, 0, 255).astype(np.uint8)
        orc.deinterleave(soft.ctypes.data_as(UP))
        vecs.append(soft)
    vecs = np.stack(vecs)

    def oracle_fano(v, maxcycles):
        met, cyc, mx = C.c_uint(), C.c_uint(), C.c_uint()
        data = (C.c_ubyte * 12)()
        s = v.copy()
        rc = orc.fano(C.byref(met), C.byref(cyc), C.byref(mx), data, s.ctypes.data_as(UP), 81, mettab, 60, maxcycles)
        return rc, met.value, cyc.value, mx.value, bytes(data)[:10]

    for maxcycles, subset in ((300, vecs), (10000, vecs[:24])):
        want = [oracle_fano(v, maxcycles) for v in subset]
        assert any(x[0] == 0 for x in want) and any(x[0] != 0 for x in want)
        for solo in (0, 1):             # 32 attempts per warp / one per warp
            got = w.fano_batch(subset, maxcycles=maxcycles, solo=solo)
            for k, x in enumerate(want):
                assert (got["rc"][k], got["metric"][k], got["cycles"][k], got["maxnp"][k]) == x[:4], (maxcycles, solo, k)
                if x[0] == 0:
                    assert bytes(got["data"][k][:10]) == x[4]
        # the instantiation the decode kernels run (time-out test every 256 trips, maxnp not tracked): return code, cycle
        # count and decoded bytes as fano.c's; the path metric too whenever the decode succeeds
        got = w.fano_batch(subset, maxcycles=maxcycles, solo=4)
        for k, x in enumerate(want):
            assert (got["rc"][k], got["cycles"][k]) == (x[0], x[2]), (maxcycles, "fast", k)
            if x[0] == 0:
                assert got["metric"][k] == x[1] and bytes(got["data"][k][:10]) == x[4]
    # a budgeted run either finishes with the same answer or reports FANO_STOPPED (2) -- never a different answer
    got = w.fano_batch(vecs, maxcycles=10000, stop_after=2048)
    full = w.fano_batch(vecs, maxcycles=10000)
    for k in range(len(vecs)):
        if got["rc"][k] == 2:
            assert full["cycles"][k] > 2048
        else:
            assert (got["rc"][k], got["cycles"][k], bytes(got["data"][k])) == (full["rc"][k], full["cycles"][k], bytes(full["data"][k]))


# ---- front end ----------------------------------------------------------------------------------------------------
def oracle_decimate(raw, n_iq, max_out):
    orc = po.oracle()
    orc.oracle_decimate.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
    io, qo = np.zeros(max_out, np.float32), np.zeros(max_out, np.float32)
    n = orc.oracle_decimate(raw.ctypes.data, n_iq, io.ctypes.data, qo.ctypes.data, max_out)
    return io, qo, n


def test_frontend_against_reference_golden():
    from test_oracle_vs_ref import frontend_golden_streams
    raw, n_iq, fe = frontend_golden_streams()
    Ig, Qg, n = w.decimate_batch(raw, max_out=128)
    assert n == 70
    for s in range(3):
        assert np.array_equal(Ig[s, :n], fe["i%d" % s]) and np.array_equal(Qg[s, :n], fe["q%d" % s])
        assert not Ig[s, n:].any() and not Qg[s, n:].any()


@pytest.mark.parametrize("n_iq", [0, 5, 6400, 6401, 6402, 6401 * 37 + 6400, 6401 * 300 + 8])
def test_frontend_ragged_lengths_against_oracle(n_iq):
    rng = np.random.default_rng(n_iq)
    raw = rng.integers(0, 256, size=(4, 2 * n_iq), dtype=np.uint8)
    if n_iq:
        raw[0, :] = 0          # rail: int8 negation wrap on every negated sample
        raw[1, :] = 255
    Ig, Qg, n = w.decimate_batch(raw, n_iq=n_iq, max_out=320)
    assert n == min(n_iq // 6401, 320)
    for s in range(4):
        io, qo, no = oracle_decimate(raw[s], n_iq, 320)
        assert no == n and np.array_equal(io, Ig[s]) and np.array_equal(qo, Qg[s])


def test_frontend_full_length_stream_then_decode():
    """BASELINE config 4 at full size: one 120 s stream (288 000 000 IQ pairs, 576 MB) -> 44 992 samples, bit-identical
    to the oracle; then hand-off normalisation + decode finds the embedded signal."""
    n_iq = 288_000_000
    sym = H.channel_symbols("K1JT FN20 20")
    raw = corpus.make_raw_stream(4, 0, n_iq, f0=37.0, snr=-8.0, symbols=sym)
    Ig, Qg, n = w.decimate_batch(raw[None, :], max_out=corpus.NSAMP)
    assert n == 44992
    io, qo, no = oracle_decimate(raw, n_iq, corpus.NSAMP)
    assert no == n and np.array_equal(io, Ig[0]) and np.array_equal(qo, Qg[0])
    with w.BatchDecoder(1) as d:
        d.upload(Ig, Qg)
        d.normalise()
        d.decode()
        spots, nres = d.download()
    a, b = io.copy(), qo.copy()
    po.oracle().oracle_normalise(a.ctypes.data_as(FP), b.ctypes.data_as(FP), a.shape[0])
    r, _, _ = po.decode(po.oracle(), a, b)
    assert H.results_equal(r, spots[0, : nres[0]])
    assert nres[0] >= 1 and spots[0, 0]["message"] == b"K1JT FN20 20"


def test_streaming_frontend_matches_callback_across_pushes_and_slots():
    """wspr_frontend_push / _swap (the daemon's receive callback and slot switch, rtlsdr_wsprd.c:126-244,1181-1183):
    ragged chunk sizes, filter state carried across pushes and across slots, outputs beyond the slot length dropped."""
    n_iq = 6401 * 150 + 2002                         # (2 * n_iq is a multiple of 8, like every chunk below)
    rng = np.random.default_rng(99)
    raw = rng.integers(0, 256, size=(3, 2 * n_iq), dtype=np.uint8)
    raw[1, : 2 * 6401 * 40] = 0                      # rails: the int8 negation wrap on every negated sample
    raw[2, 2 * 6401 * 20:] = 255
    want = [oracle_decimate(raw[s], n_iq, 200) for s in range(3)]          # whole stream, zero initial state
    assert want[0][2] == 150
    chunks = [65536, 8, 12800, 16, 262144, 24, 65536 * 3, 12808, 40]
    with w.FrontEnd(3, slot_samples=64) as fe:
        pos, k, produced, slots = 0, 0, 0, []
        while pos < 2 * n_iq:
            nb = min(chunks[k % len(chunks)], 2 * n_iq - pos)
            k += 1
            produced += fe.push(raw[:, pos:pos + nb])
            pos += nb
            if len(slots) == 0 and produced >= 20:       # first slot switch, part-way through
                slots.append((produced, fe.swap(), fe.read()))
            elif len(slots) == 1 and produced >= 150:    # second slot overflowed its 64 samples
                slots.append((produced, fe.swap(), fe.read()))
        assert produced == 150 and fe.samples() == 0
    (p0, n0, (I0, Q0, r0)), (p1, n1, (I1, Q1, r1)) = slots
    assert n0 == r0 == min(p0, 64) and n1 == r1 == 64 and p1 - p0 > 64
    for s in range(3):
        io, qo, _ = want[s]
        assert np.array_equal(I0[s, :n0], io[:n0]) and np.array_equal(Q0[s, :n0], qo[:n0])
        assert not I0[s, n0:].any() and not Q0[s, n0:].any()
        assert np.array_equal(I1[s], io[p0:p0 + 64]) and np.array_equal(Q1[s], qo[p0:p0 + 64])


def test_streaming_frontend_against_reference_callback_and_decoder_hand_off():
    """One receiver fed librtlsdr-sized buffers (65536 bytes, rtlsdr_wsprd.c:42): identical to the reference's own
    rtlsdr_callback; then the device-resident hand-off into a decode context equals upload + normalise."""
    try:
        ref = po.RefFrontend()
    except (FileNotFoundError, OSError):
        pytest.skip("compiled reference front end not available")
    n_bytes = 65536 * 60
    raw = np.random.default_rng(5).integers(0, 256, size=n_bytes, dtype=np.uint8)
    ref.push(raw)
    ir, qr = ref.read()
    with w.FrontEnd(1) as fe:
        for off in range(0, n_bytes, 65536):
            fe.push(raw[off:off + 65536])
        n = fe.swap()
        I, Q, _ = fe.read()
        assert n == len(ir) == n_bytes // 2 // 6401
        assert np.array_equal(I[0, :n], ir) and np.array_equal(Q[0, :n], qr)
        with w.BatchDecoder(1) as d:
            assert fe.hand_off(d) == n
            _, _, Id, Qd = d.download(samples=True)
    a, b = I[0].copy(), Q[0].copy()
    po.oracle().oracle_normalise(a.ctypes.data_as(FP), b.ctypes.data_as(FP), a.shape[0])
    assert np.array_equal(Id[0], a) and np.array_equal(Qd[0], b)


def test_subtract_signal_abi_against_compiled_reference():
    """subtract_signal (wsprd/wsprd.h:92-98, the per-symbol variant the reference exports but never calls) through the C ABI
    against the reference's own object code (oracle/_ref travels to the GPU box prebuilt), with and without drift."""
    ref = po.ref()
    if ref is None:
        pytest.skip("oracle/_ref/libwsprd_ref.so not available")
    ref.subtract_signal.restype = None
    ref.subtract_signal.argtypes = [FP, FP, C.c_long, C.c_float, C.c_int, C.c_float, UP]
    lib = w.library()
    lib.subtract_signal.restype = None
    lib.subtract_signal.argtypes = [FP, FP, C.c_long, C.c_float, C.c_int, C.c_float, UP]
    I, Q, plans = H.make_corpus(3, 1, start=77)
    for k, (f0, shift, drift) in enumerate([(plans[0][3]["f0"], 760, 0.0), (-41.7, 3, -1.0), (12.3, 4000, 2.0), (99.0, -30, 0.0)]):
        chan = H.channel_symbols(plans[0][k]["message"])
        ia, qa, ib, qb = I[0].copy(), Q[0].copy(), I[0].copy(), Q[0].copy()
        ref.subtract_signal(ia.ctypes.data_as(FP), qa.ctypes.data_as(FP), 45000, C.c_float(f0), shift, C.c_float(drift), chan.ctypes.data_as(UP))
        lib.subtract_signal(ib.ctypes.data_as(FP), qb.ctypes.data_as(FP), 45000, C.c_float(f0), shift, C.c_float(drift), chan.ctypes.data_as(UP))
        assert not np.array_equal(ia, I[0])
        assert np.array_equal(ia, ib) and np.array_equal(qa, qb), (k, np.abs(ia - ib).max())


def test_four_passes_keep_pass_two_parameters():
    """npasses >= 4: passes 3.. keep maxdrift = 0 / minsync2 = 0.10 of pass 2 (wsprd.c:524-531 only assign for ipass < 2 and == 2)."""
    I, Q, _ = H.make_corpus(3, 4, start=2100, )
    opt_o, opt_g = po.default_options(npasses=4), w.default_options(npasses=4)
    got = w.decode_batch(I, Q, opt_g)
    for c in range(len(I)):
        a, _, _ = po.decode(po.ref() or po.oracle(), I[c], Q[c], opt_o)
        assert H.results_equal(a, got[c]), (c, H.diff_results(a, got[c]))


def test_two_devices_in_one_process():
    """Contexts on two GPUs of one process (kernel attributes and the Fano service are per device)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    I, Q, _ = H.make_corpus(3, 6, start=2200)
    want = [po.decode(po.oracle(), I[c], Q[c])[0] for c in range(len(I))]
    with w.BatchDecoder(6, device=0) as d0, w.BatchDecoder(6, device=1) as d1:
        for d in (d1, d0, d1):
            d.upload(I, Q)
            d.decode()
            spots, n = d.download()
            for c in range(len(I)):
                assert H.results_equal(want[c], spots[c, : n[c]]), (d.device, c)


def test_hashtable_batch_versus_capture_by_capture():
    """-H with several captures in ONE call: every capture sees hashtable.txt as it was on entry and the additions are merged
    in capture order.  The reference's call-by-call order differs exactly where a capture refers by hash to a callsign that
    an EARLIER capture of the same batch taught: the batch leaves <...> there, the sequential run resolves it.  Everything
    else -- and the file left behind -- is identical."""
    import tempfile
    caps = H.hashtable_scenario()                      # A teaches, B refers by hash, A again
    I, Q = np.stack([c[0] for c in caps]), np.stack([c[1] for c in caps])
    opt_o, opt_g = po.default_options(usehashtable=1), w.default_options(usehashtable=1)
    old = os.getcwd()
    with tempfile.TemporaryDirectory(prefix="wspr_htb_") as d:
        os.chdir(d)
        try:
            seq = [po.decode(po.oracle(), I[c], Q[c], opt_o, cwd_scratch=False)[0] for c in range(3)]
            seq_file = open("hashtable.txt").read()
            os.remove("hashtable.txt")
            got = w.decode_batch(I, Q, opt_g)
            batch_file = open("hashtable.txt").read()
            again = w.decode_batch(I, Q, opt_g)        # second call: the table of the first is on disk now
        finally:
            os.chdir(old)
    assert batch_file == seq_file
    assert H.results_equal(seq[0], got[0]) and H.results_equal(seq[2], got[2])
    unresolved = [x["message"] for x in got[1] if b"<...>" in x["message"]]
    resolved = [x["message"] for x in seq[1] if x["message"].startswith(b"<") and b"<...>" not in x["message"]]
    assert len(unresolved) >= len(resolved) >= 2 and len(got[1]) == len(seq[1])
    assert H.results_equal(seq[1], again[1])           # with the table on disk the batch resolves them like the reference


def test_sharded_decode_on_two_gpus(tmp_path):
    """sharding.decode_sharded with the real decoder, one process per GPU over NCCL (launched like the driver launches
    bench.py), against the oracle."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    out = str(tmp_path / "sharded.npz")
    script = os.path.join(H.ROOT, "tests", "sharded_worker.py")
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                    "--master-port", "29611", script, out, "7"], check=True, timeout=600)
    z = np.load(out)
    spots, n = z["spots"].view(w.RESULT_DTYPE).reshape(-1, w.MAX_UNIQUES), z["n"]
    I, Q, _ = H.make_corpus(3, 7, start=2300)
    assert len(n) == 7
    for c in range(7):
        a, _, _ = po.decode(po.oracle(), I[c], Q[c])
        assert H.results_equal(a, spots[c, : n[c]]), (c, H.diff_results(a, spots[c, : n[c]]))


def test_quick_and_normal_decodes_alternate_on_one_context():
    """One context decodes the same weak-signal captures in quick mode, in normal mode and in quick mode again: candidates
    are parked in all three (quick mode parks an unfinished jitter-0 attempt with ONE attempt, normal mode with 43), on the
    context's two sets of scratch records (lease_scratch in wspr_decode.cu), and every decode equals the oracle's."""
    n = 3
    I = np.zeros((n, corpus.NSAMP), np.float32)
    Q = np.zeros_like(I)
    for c in range(n):
        plan = corpus.ten_signal_plan(970 + c, snrs=np.arange(-31.0, -22.0, 1.5))
        I[c], Q[c] = corpus.make_capture(9, 970 + c, plan, H.channel_symbols)
    want = {}
    for quick in (1, 0):
        want[quick] = [po.decode(po.oracle(), I[c], Q[c], po.default_options(quickmode=quick)) for c in range(n)]
    parked = 0
    with w.BatchDecoder(n, corpus.NSAMP) as d:
        for quick in (1, 0, 1, 0):
            d.upload(I, Q)
            d.decode(w.default_options(quickmode=quick))
            spots, cnt, Io, Qo = d.download(samples=True)
            parked += d.schedule_stats()[1]
            for c in range(n):
                a, ia, qa = want[quick][c]
                assert H.results_equal(a, spots[c, : cnt[c]]), (quick, c, H.diff_results(a, spots[c, : cnt[c]]))
                assert np.array_equal(ia, Io[c]) and np.array_equal(qa, Qo[c]), (quick, c)
    assert parked > 0, "no candidate was parked: the corpus does not exercise the scratch records"
    st = w.fano_pool_stats()                                      # the pool's counters (wspr_fano_stats)
    assert st["pool_warps"] >= 1 and st["worker_warps_started"] >= 1 and st["attempts_run"] + st["attempts_dropped"] > 0
    assert 0.0 < st["lane_utilisation"] <= 1.0


def test_candidate_loop_breaks():
    """The reference's two early exits from the candidate loop (wsprd.c:781-793): a decoded message that cannot be encoded
    again for the subtraction, and the locator 'A000AA' (helpers.break_captures).  The capture's pass ends there (CapState::
    broken), nothing has been recorded yet, so pass 1 does not run either: no spots, samples untouched -- like the oracle,
    with and without the subtraction."""
    caps = H.break_captures()
    I = np.stack([c[1][0] for c in caps])
    Q = np.stack([c[1][1] for c in caps])
    spots, n, Io, Qo = gpu_decode(I, Q)
    assert list(n) == [0, 0]
    assert_batch_equals_oracle(I, Q)
    assert assert_batch_equals_oracle(I, Q, subtraction=0) >= 3


def _device_bytes(a):
    """A device copy of a uint8 array -> (keep-alive object, device pointer).  Under the host emulation (tools/cuda_emu)
    device pointers are host pointers and there is no CUDA for torch to use."""
    import torch
    if torch.cuda.is_available():
        t = torch.from_numpy(a).cuda()
        torch.cuda.synchronize()
        return t, t.data_ptr()
    assert "emu" in os.path.basename(w.library_path()), "no CUDA device and not the emulated build"
    return a, a.ctypes.data


def test_context_decimate_normalise_decode_chain():
    """Raw streams that are already on the device go through the front end straight into a context (wspr_ctx_decimate: the
    callback for whole streams, tail zeroed, rtlsdr_wsprd.c:126-244,285-288), are peak-normalised there (:291-305) and decoded
    (:316) -- the chain bench.py's config 4 runs.  Every stage against the oracle."""
    nstreams, n_iq = 3, 6401 * 70 + 2000
    stride = (2 * n_iq + 15) // 16 * 16 + 16
    rng = np.random.default_rng(4)
    raw = np.zeros((nstreams, stride), np.uint8)
    raw[:, : 2 * n_iq] = np.clip(127.5 + 30 * rng.standard_normal((nstreams, 2 * n_iq)), 0, 255).astype(np.uint8)
    raw[1, 4000:9000] = 0                                                  # a stretch on the rail: the int8 -(-128) wrap
    keep, ptr = _device_bytes(raw)
    orc = po.oracle()
    orc.oracle_decimate.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
    with w.BatchDecoder(nstreams, corpus.NSAMP) as d:
        assert d.decimate(ptr, nstreams, n_iq, stride) is not None
        _, _, I1, Q1 = d.download(samples=True)
        want = []
        for s in range(nstreams):
            io, qo = np.zeros(corpus.NSAMP, np.float32), np.zeros(corpus.NSAMP, np.float32)
            n = orc.oracle_decimate(raw[s].ctypes.data, n_iq, io.ctypes.data, qo.ctypes.data, corpus.NSAMP)
            assert n == 70
            assert np.array_equal(I1[s], io) and np.array_equal(Q1[s], qo), s          # 70 outputs, then zeros
            want.append(po.normalise_half(io, qo))
        d.normalise()
        d.decode()
        spots, cnt, I2, Q2 = d.download(samples=True)
        for s in range(nstreams):
            a, ia, qa = po.decode(orc, want[s][0], want[s][1])
            assert H.results_equal(a, spots[s, : cnt[s]]), (s, H.diff_results(a, spots[s, : cnt[s]]))
            assert np.array_equal(ia, I2[s]) and np.array_equal(qa, Q2[s]), s
    del keep
