#!/usr/bin/env python3
"""Generate the committed golden vectors under tests/golden/ by running the REFERENCE ITSELF (oracle/_ref, the
unmodified wsprd/*.c + rtlsdr_callback of /root/reference compiled by oracle/Makefile) in this container.

  golden_decode.json    spot lists (every field of struct decoder_results) + sha256 of the post-subtraction samples
                        for seeded captures of BASELINE configs 2 and 3 and for the reference's own fixture
  golden_stages.npz     intermediates of capture (config 3, index 0): sync_and_demodulate outputs in the three
                        modes and a subtract_signal2 result, straight from the compiled reference
  golden_frontend.npz   rtlsdr_callback outputs for a short seeded raw stream (incl. the int8 -(-128) wrap case)

Inputs are regenerated from the seed by rtlsdr_wsprd_b200/corpus.py, so only outputs are stored.
Usage: python tools/make_golden.py     (needs /root/reference; a few seconds)
"""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pyoracle as po  # noqa: E402
import helpers as H  # noqa: E402

N_CFG2, N_CFG3 = 48, 24


def spots_to_json(r):
    out = []
    for x in r:
        out.append({"freq": float(x["freq"]).hex(), "sync": float(x["sync"]).hex(), "snr": float(x["snr"]).hex(),
                    "dt": float(x["dt"]).hex(), "drift": float(x["drift"]), "jitter": int(x["jitter"]),
                    "message": x["message"].decode(), "call": x["call"].decode(), "loc": x["loc"].decode(),
                    "pwr": x["pwr"].decode(), "cycles": int(x["cycles"]), "line": po.spot_line(x)})
    return out


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ref = po.ref()
    assert ref is not None, "needs /root/reference"
    gold = {"generator": "tools/make_golden.py", "source": "compiled reference (oracle/_ref/libwsprd_ref.so)", "cases": {}}
    i, q = po.read_iq_file(os.path.join(H.GOLDEN, "refSignalSnr0dB.iq"))
    r, io, qo = po.decode(ref, i, q)
    gold["cases"]["fixture"] = [{"spots": spots_to_json(r), "i_sha": sha(io), "q_sha": sha(qo)}]
    for cfg, n in ((2, N_CFG2), (3, N_CFG3)):
        I, Q, _ = H.make_corpus(cfg, n)
        cases = []
        for c in range(n):
            r, io, qo = po.decode(ref, I[c], Q[c])
            cases.append({"spots": spots_to_json(r), "i_sha": sha(io), "q_sha": sha(qo), "in_sha": sha(I[c])})
        gold["cases"]["config%d" % cfg] = cases
        print("config", cfg, "captures", n, "spots", sum(len(c["spots"]) for c in cases))
    # options variants on config 3 capture 0..3: quick mode, single pass without subtraction
    I, Q, _ = H.make_corpus(3, 4)
    for name, opt in (("quick", po.default_options(quickmode=1)), ("single", po.default_options(npasses=1, subtraction=0))):
        cases = []
        for c in range(4):
            r, io, qo = po.decode(ref, I[c], Q[c], opt)
            cases.append({"spots": spots_to_json(r), "i_sha": sha(io), "q_sha": sha(qo)})
        gold["cases"]["config3_" + name] = cases
    with open(os.path.join(H.GOLDEN, "golden_decode.json"), "w") as f:
        json.dump(gold, f, indent=0, sort_keys=True)

    # ---- stage intermediates from the compiled reference ----
    fp = C.POINTER(C.c_float)
    i0, q0 = I[0].copy(), Q[0].copy()
    r, _, _ = po.decode(ref, i0, q0, po.default_options(npasses=1, subtraction=0))
    st = {}
    for k, x in enumerate(r[:3]):
        f1 = float(x["freq"] * 1e6 - 144489000 - 1500)
        shift0 = int(round((float(x["dt"]) + 2.0) * 375.0))
        for drift in (0.0, 1.0):
            freq, shift, dr, sync = C.c_float(round(f1)), C.c_int(shift0 - 24), C.c_float(drift), C.c_float(0)
            sym = (C.c_ubyte * 162)()
            args = lambda: (i0.ctypes.data_as(fp), q0.ctypes.data_as(fp), 45000, sym, C.byref(freq))
            ref.sync_and_demodulate(*args(), 0, 0, 0.0, C.byref(shift), shift.value - 128, shift.value + 128, 8, C.byref(dr), 50, C.byref(sync), 0)
            m0 = (freq.value, shift.value, sync.value)
            ref.sync_and_demodulate(*args(), -2, 2, 0.1, C.byref(shift), shift.value, shift.value, 1, C.byref(dr), 50, C.byref(sync), 1)
            m1 = (freq.value, shift.value, sync.value)
            ref.sync_and_demodulate(*args(), 0, 0, 0.0, C.byref(shift), shift.value, shift.value, 1, C.byref(dr), 50, C.byref(sync), 2)
            st["sync_%d_%d_in" % (k, int(drift))] = np.array([round(f1), shift0 - 24, drift], np.float64)
            st["sync_%d_%d_m0" % (k, int(drift))] = np.array(m0, np.float64)
            st["sync_%d_%d_m1" % (k, int(drift))] = np.array(m1, np.float64)
            st["sync_%d_%d_m2" % (k, int(drift))] = np.array([sync.value], np.float64)
            st["sync_%d_%d_sym" % (k, int(drift))] = np.frombuffer(bytes(sym), np.uint8).copy()
    x = r[0]
    chan = H.channel_symbols(x["message"].decode(), ref)
    f1 = np.float32(x["freq"] * 1e6 - 144489000 - 1500)
    shift0 = int(round((float(x["dt"]) + 2.0) * 375.0))
    for drift in (0.0, -1.0):
        ia, qa = i0.copy(), q0.copy()
        ref.subtract_signal2(ia.ctypes.data_as(fp), qa.ctypes.data_as(fp), 45000, C.c_float(f1), shift0, C.c_float(drift),
                             chan.ctypes.data_as(C.POINTER(C.c_ubyte)))
        st["sub_%d_in" % int(drift)] = np.array([f1, shift0, drift], np.float64)
        st["sub_%d_i" % int(drift)] = ia
        st["sub_%d_q" % int(drift)] = qa
    st["sub_chan"] = chan
    np.savez_compressed(os.path.join(H.GOLDEN, "golden_stages.npz"), **st)

    # ---- front end ----
    n_iq = 6401 * 70 + 777
    rng = np.random.default_rng(7)
    raw = rng.integers(0, 256, size=(3, 2 * n_iq), dtype=np.uint8)
    raw[1, : 2 * 6401 * 8] = 0         # saturated low: every negated sample hits -(-128)
    raw[2, ::5] = 0
    raw[2, 1::3] = 255
    fe = {}
    for s in range(3):
        f = po.RefFrontend()
        f.push(raw[s])
        io, qo = f.read()
        fe["i%d" % s], fe["q%d" % s] = io, qo
    fe["seed"] = np.array([7, n_iq])
    np.savez_compressed(os.path.join(H.GOLDEN, "golden_frontend.npz"), **fe)
    print("front end outputs per stream:", len(fe["i0"]))


if __name__ == "__main__":
    main()
