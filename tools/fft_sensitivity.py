"""How far is "parity with the reference + double-precision stand-in FFT" from "parity with the reference + FFTW3f"?
FFTW3f is not installed anywhere here, so the question is bounded instead: decode the config-3 corpus with the UNMODIFIED
reference linked against (a) the binary64 stand-in the oracle and the GPU path reproduce bit for bit and (b) the same
transform evaluated in binary32 in two different orders -- perturbations of the spectrogram of the size of FFTW3f's own
rounding error (~1e-7 relative) -- and count the captures whose spot list changes.
usage: fft_sensitivity.py [ncaptures] [config]   (CPU only; needs /root/reference: `make -C oracle ref ref_f32`)"""
import ctypes as C
import multiprocessing as mp
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

LIBS = {"f64": "libwsprd_ref.so", "f32dit": "libwsprd_ref_f32dit.so", "f32dif": "libwsprd_ref_f32dif.so"}


def work(args):
    config, lo, hi = args
    from oracle import pyoracle as po
    import helpers as H
    libs = {k: po._bind_decode(C.CDLL(os.path.join(ROOT, "oracle", "_ref", v))) for k, v in LIBS.items()}
    os.chdir(tempfile.mkdtemp(prefix="wspr_fft_"))
    I, Q, _ = H.make_corpus(config, hi - lo, start=lo)
    out = []
    for c in range(hi - lo):
        row = {}
        for k, lib in libs.items():
            r, _, _ = po.decode(lib, I[c], Q[c], cwd_scratch=False)
            row[k] = [(x["message"], x["call"], x["loc"], x["pwr"], float(x["freq"]), float(x["snr"]), float(x["dt"]), float(x["drift"]),
                       float(x["sync"]), int(x["jitter"]), int(x["cycles"])) for x in r]
        out.append(row)
    return lo, out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    config = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    jobs = [(config, a, min(a + 16, n)) for a in range(0, n, 16)]
    t0 = time.time()
    with mp.Pool(os.cpu_count()) as pool:
        parts = dict(pool.imap_unordered(work, jobs))
    rows = [r for a in sorted(parts) for r in parts[a]]
    print("config %d, %d captures, %d spots with the binary64 stand-in; %.0f s on %d cores" %
          (config, n, sum(len(r["f64"]) for r in rows), time.time() - t0, os.cpu_count()))
    for k in ("f32dit", "f32dif"):
        hard = sum(1 for r in rows if [x[:4] for x in r["f64"]] != [x[:4] for x in r[k]])
        setdiff = sum(1 for r in rows if sorted(x[:4] for x in r["f64"]) != sorted(x[:4] for x in r[k]))
        full = sum(1 for r in rows if r["f64"] != r[k])
        spots_changed = sum(len(set(x[:4] for x in r["f64"]) ^ set(x[:4] for x in r[k])) for r in rows)
        fields = {}
        for r in rows:
            if [x[:4] for x in r["f64"]] == [x[:4] for x in r[k]]:
                for a, b in zip(r["f64"], r[k]):
                    for name, u, v in zip(("freq", "snr", "dt", "drift", "sync", "jitter", "cycles"), a[4:], b[4:]):
                        if u != v:
                            fields[name] = fields.get(name, 0) + 1
        print("%-7s captures whose (message, call, loc, pwr) list differs: %d (as a set: %d; spots gained or lost: %d); "
              "captures with any field different: %d; fields that differ on otherwise identical lists: %s"
              % (k, hard, setdiff, spots_changed, full, fields))


if __name__ == "__main__":
    main()
