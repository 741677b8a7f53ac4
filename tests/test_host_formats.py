"""Host-side formats either side of the path: .iq / .c2 files, the spot text lines, report URLs,
the corpus generator.  CPU only."""
import struct
import time

import numpy as np

from oracle import pyoracle as po
import rtlsdr_wsprd_b200 as w
from rtlsdr_wsprd_b200 import corpus
import helpers as H
import os


def test_iq_file_roundtrip_and_reference_reader(tmp_path):
    I, Q, _ = H.make_corpus(2, 1)
    path = str(tmp_path / "x.iq")
    assert w.write_iq_file(path, I[0], Q[0]) == corpus.NSAMP
    raw = np.fromfile(path, "<f4")
    assert raw.shape[0] == 2 * corpus.NSAMP and np.array_equal(raw[0::2], I[0]) and np.array_equal(raw[1::2], -Q[0])
    i2, q2 = w.read_iq_file(path)
    i3, q3 = po.read_iq_file(path)                       # the oracle-side reader (rtlsdr_wsprd.c:555-592)
    assert np.array_equal(i2, i3) and np.array_equal(q2, q3)
    assert np.float32(max(np.abs(i2).max(), np.abs(q2).max())) == np.float32(0.5)
    # the reference's own fixture through both readers
    a = w.read_iq_file(os.path.join(H.GOLDEN, "refSignalSnr0dB.iq"))
    b = po.read_iq_file(os.path.join(H.GOLDEN, "refSignalSnr0dB.iq"))
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[0].shape[0] == 45000


def test_c2_file_reader(tmp_path):
    I, Q, _ = H.make_corpus(2, 1)
    path = str(tmp_path / "000000_0000.c2")
    body = np.empty(2 * corpus.NSAMP, "<f4")
    body[0::2], body[1::2] = I[0], -Q[0]
    with open(path, "wb") as f:                          # header: char[14] name, int type, double freq (rtlsdr_wsprd.c:634-637)
        f.write(b"000000_0000.c2")
        f.write(struct.pack("<i", 2))
        f.write(struct.pack("<d", 14.0956))
        f.write(body.tobytes())
    i2, q2, freq = w.read_c2_file(path)
    assert freq == 14.0956
    ni, nq = w.normalise_half(I[0], Q[0])
    assert np.array_equal(i2, ni) and np.array_equal(q2, nq)


def test_normalise_matches_oracle():
    rng = np.random.default_rng(4)
    i = (rng.standard_normal(45000) * 3).astype(np.float32)
    q = (rng.standard_normal(45000) * 3).astype(np.float32)
    a, b = w.normalise_half(i, q)
    import ctypes as C
    x, y = i.copy(), q.copy()
    po.oracle().oracle_normalise(x.ctypes.data_as(C.POINTER(C.c_float)), y.ctypes.data_as(C.POINTER(C.c_float)), 45000)
    assert np.array_equal(a, x) and np.array_equal(b, y)


def test_spot_text_contracts():
    r = np.zeros(1, w.RESULT_DTYPE)
    r[0]["snr"], r[0]["dt"], r[0]["freq"], r[0]["drift"] = -0.07, 0.01, 144.49055, 0.0
    r[0]["call"], r[0]["loc"], r[0]["pwr"], r[0]["message"] = b"K1JT", b"FN20", b"20", b"K1JT FN20 20"
    assert w.spot_line(r[0]) == " -0.07   0.01 144.490550  0    K1JT   FN20 20"       # documentation/bug-fix/REPORT.md:202
    g = time.struct_time((2026, 10, 17, 4, 20, 0, 5, 290, 0))
    assert w.print_spots_lines(r, g) == ["Spot :  2026-10-17 04:20z  -0.07   0.01 144.490550  0    K1JT   FN20 20"]
    assert w.print_spots_lines(r[:0], g) == ["No spot 2026-10-17 04:20z"]
    url = w.wsprnet_urls(r, "A1XYZ", "AB12CD", 144489000, g)[0]
    assert url == ("https://wsprnet.org/post?function=wspr&rcall=A1XYZ&rgrid=AB12CD&rqrg=144.490550&date=261017&time=0420&sig=-0&dt=0.0"
                   "&tqrg=144.490550&tcall=K1JT&tgrid=FN20&dbm=20&version=rtlsdr-056&mode=2")
    assert "function=wsprstat" in w.wsprnet_urls(r[:0], "A1XYZ", "AB12CD", 144489000, g)[0]


def test_corpus_is_seeded_and_counter_based():
    a = H.make_corpus(3, 2, start=5)
    b = H.make_corpus(3, 1, start=6)
    assert np.array_equal(a[0][1], b[0][0]) and np.array_equal(a[1][1], b[1][0])     # capture 6 regenerated on its own
    assert a[2][1] == b[2][0] and len(a[2][0]) == 10
    snrs = sorted(s["snr"] for s in a[2][0])
    assert snrs == list(np.arange(-28.0, -9.0, 2.0))


def test_batch_file_loader(tmp_path):
    """N2: a directory of .iq / .c2 recordings -> the planar [n][45000] batch, every file through the reference's reader
    semantics (Q negated, peak-normalised, short files zero-padded)."""
    I, Q, _ = H.make_corpus(2, 3)
    w.write_iq_file(str(tmp_path / "b.iq"), I[0], Q[0])
    w.write_c2_file(str(tmp_path / "a.c2"), I[1], Q[1], 7.0386)
    w.write_iq_file(str(tmp_path / "c.iq"), I[2][:30000], Q[2][:30000])
    (tmp_path / "notes.txt").write_text("ignored")
    bi, bq, freq, names = w.load_capture_files(str(tmp_path))
    assert names == ["a.c2", "b.iq", "c.iq"] and bi.shape == (3, 45000) and bi.dtype == np.float32
    assert freq.tolist() == [7.0386, 0.0, 0.0]
    for k, src in enumerate((1, 0)):
        ni, nq = w.normalise_half(I[src], Q[src])
        assert np.array_equal(bi[k], ni) and np.array_equal(bq[k], nq)
    si, sq = w.normalise_half(I[2][:30000], Q[2][:30000])
    assert np.array_equal(bi[2, :30000], si) and np.array_equal(bq[2, :30000], sq) and not bi[2, 30000:].any() and not bq[2, 30000:].any()
    # a list of paths, and the reference's own fixture through the loader and the oracle-side reader
    fix = os.path.join(H.GOLDEN, "refSignalSnr0dB.iq")
    li, lq, _, _ = w.load_capture_files([fix, str(tmp_path / "b.iq")])
    ri, rq = po.read_iq_file(fix)
    assert np.array_equal(li[0], ri) and np.array_equal(lq[0], rq) and np.array_equal(li[1], bi[1])
