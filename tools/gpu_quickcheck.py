"""Ad-hoc GPU bring-up check (not a test): CUDA path vs oracle on a few captures, stage by stage."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ctypes as C
from oracle import pyoracle as po
import rtlsdr_wsprd_b200 as w
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import helpers as H

orc = po.oracle()
i, q = w.read_iq_file(os.path.join(H.GOLDEN, "refSignalSnr0dB.iq"))
r = w.wspr_decode(i.copy(), q.copy())
print("fixture:", [w.spot_line(x) for x in r])

for config, n in ((2, 16), (3, 16)):
    I, Q, plans = H.make_corpus(config, n)
    with w.BatchDecoder(n) as d:
        d.upload(I, Q)
        # stages
        ps = d.spectrogram()
        blocks = ps.shape[2]
        ps_o = np.zeros((512, blocks), np.float32)
        orc.oracle_spectrogram.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        orc.oracle_spectrogram(I[0].ctypes.data, Q[0].ctypes.data, I.shape[1], ps_o.ctypes.data)
        print("cfg", config, "ps equal:", np.array_equal(ps[0], ps_o), "maxrel", float(np.max(np.abs(ps[0]-ps_o)/np.maximum(ps_o,1e-30))))
        cands, npk = d.candidates()
        co = np.zeros(200, w.CAND_DTYPE)
        orc.oracle_candidates.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        orc.oracle_candidates.restype = C.c_int
        nk = orc.oracle_candidates(ps_o.ctypes.data, blocks, 4, co.ctypes.data, None)
        print("   npk", npk[0], nk, "cands equal:", cands[0][:nk].tobytes() == co[:nk].tobytes())
        if cands[0][:nk].tobytes() != co[:nk].tobytes():
            print(cands[0][:nk]); print(co[:nk])
        d.upload(I, Q)
        ms = d.decode()
        spots, nres, Io, Qo = d.download(samples=True)
    bad = 0
    for c in range(n):
        a, ia, qa = po.decode(orc, I[c], Q[c])
        b = spots[c, :nres[c]]
        same = H.results_equal(a, b)
        siq = np.array_equal(ia, Io[c]) and np.array_equal(qa, Qo[c])
        if not (same and siq):
            bad += 1
            print("   capture", c, "MISMATCH", H.diff_results(a, b)[:6], "iq equal", siq,
                  "maxdiff", float(np.max(np.abs(ia - Io[c]))))
    print("cfg", config, "captures", n, "mismatching", bad, "decode ms", round(ms, 2), "spots", int(nres.sum()))

# front end
nblk = 200
n_iq = 6401 * nblk + 1234
raw = np.random.default_rng(1).integers(0, 256, size=(3, 2 * n_iq), dtype=np.uint8)
raw[1, :50000] = 0          # int8 negation wrap path
raw[2, ::7] = 0
Ig, Qg, nout = w.decimate_batch(raw, max_out=256)
orc.oracle_decimate.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
for s in range(3):
    io = np.zeros(256, np.float32); qo = np.zeros(256, np.float32)
    no = orc.oracle_decimate(raw[s].ctypes.data, n_iq, io.ctypes.data, qo.ctypes.data, 256)
    print("decimate stream", s, nout, no, np.array_equal(io, Ig[s]), np.array_equal(qo, Qg[s]))

# throughput
for config, n in ((3, 256),):
    I, Q, _ = H.make_corpus(config, 32)
    I = np.tile(I, (n // 32, 1)); Q = np.tile(Q, (n // 32, 1))
    with w.BatchDecoder(n) as d:
        for it in range(3):
            d.upload(I, Q)
            ms = d.decode()
            print("cfg", config, "n", n, "decode ms", round(ms, 2), "captures/s", round(n / ms * 1e3, 1), "launches", w.kernel_launches())
