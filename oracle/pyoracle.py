"""TEST INFRASTRUCTURE -- ctypes bindings for the two CPU checkers (never imported by the product).

  * ``ref()``    -> oracle/_ref/libwsprd_ref.so : the reference's own wsprd/*.c compiled unmodified
                    (oracle/Makefile target ``ref``); exports the reference ABI (wsprd/wsprd.h:76-111).
  * ``oracle()`` -> oracle/liboracle.so         : our plain-C restatement (oracle/wspr_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import shutil
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"
NSAMP = 45000            # SIGNAL_LENGHT * SIGNAL_SAMPLE_RATE, rtlsdr_wsprd.c:37-38
MAX_UNIQUES = 100        # wsprd/wsprd.h:41


class DecoderOptions(C.Structure):
    """struct decoder_options, wsprd/wsprd.h:44-52 (40 bytes, passed BY VALUE)."""
    _fields_ = [("freq", C.c_int), ("rcall", C.c_char * 13), ("rloc", C.c_char * 7),
                ("quickmode", C.c_int), ("usehashtable", C.c_int), ("npasses", C.c_int),
                ("subtraction", C.c_int)]


class DecoderResults(C.Structure):
    """struct decoder_results, wsprd/wsprd.h:62-74 (80 bytes)."""
    _fields_ = [("freq", C.c_double), ("sync", C.c_float), ("snr", C.c_float), ("dt", C.c_float),
                ("drift", C.c_float), ("jitter", C.c_int), ("message", C.c_char * 23),
                ("call", C.c_char * 13), ("loc", C.c_char * 7), ("pwr", C.c_char * 3),
                ("cycles", C.c_int)]


assert C.sizeof(DecoderOptions) == 40 and C.sizeof(DecoderResults) == 80

RESULT_DTYPE = np.dtype({
    "names": ["freq", "sync", "snr", "dt", "drift", "jitter", "message", "call", "loc", "pwr", "cycles"],
    "formats": ["<f8", "<f4", "<f4", "<f4", "<f4", "<i4", "S23", "S13", "S7", "S3", "<i4"],
    "offsets": [0, 8, 12, 16, 20, 24, 28, 51, 64, 71, 76],
    "itemsize": 80})


def default_options(freq=144489000, npasses=2, subtraction=1, quickmode=0, usehashtable=0):
    """Defaults of initDecoder_options(), rtlsdr_wsprd.c:356-362."""
    o = DecoderOptions()
    o.freq = freq
    o.rcall = b"A1XYZ"
    o.rloc = b"AB12CD"
    o.quickmode, o.usehashtable, o.npasses, o.subtraction = quickmode, usehashtable, npasses, subtraction
    return o


def _make(target):
    subprocess.run(["make", "-s", "-C", HERE, target], check=True, stdout=subprocess.DEVNULL)


_cache = {}


def ref_available():
    return os.path.exists(os.path.join(HERE, "_ref", "libwsprd_ref.so")) or os.path.isdir(REF_ROOT + "/wsprd")


def _bind_decode(lib):
    lib.wspr_decode.restype = C.c_int
    lib.wspr_decode.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, DecoderOptions,
                                C.POINTER(DecoderResults), C.POINTER(C.c_int)]
    return lib


def ref():
    """The compiled reference (None if neither /root/reference nor a prebuilt oracle/_ref exists)."""
    if "ref" not in _cache:
        path = os.path.join(HERE, "_ref", "libwsprd_ref.so")
        if os.path.isdir(REF_ROOT + "/wsprd"):
            _make("ref")
        _cache["ref"] = _bind_decode(C.CDLL(path)) if os.path.exists(path) else None
        lib = _cache["ref"]
        if lib is not None:
            fp, ip, up = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_ubyte)
            lib.sync_and_demodulate.restype = None
            lib.sync_and_demodulate.argtypes = [fp, fp, C.c_long, up, fp, C.c_int, C.c_int, C.c_float, ip,
                                                C.c_int, C.c_int, C.c_int, fp, C.c_int, fp, C.c_int]
            lib.subtract_signal2.restype = None
            lib.subtract_signal2.argtypes = [fp, fp, C.c_long, C.c_float, C.c_int, C.c_float, up]
    return _cache["ref"]


def oracle():
    """Our C restatement; rebuilt from source when stale (gcc is present on the GPU box too)."""
    if "oracle" not in _cache:
        _make("oracle")
        _cache["oracle"] = _bind_decode(C.CDLL(os.path.join(HERE, "liboracle.so")))
    return _cache["oracle"]


def decode(lib, idat, qdat, options=None, cwd_scratch=True):
    """Run lib.wspr_decode on ONE capture.  idat/qdat: float32[n]; they are mutated like the reference does
    (copies are made here, the mutated copies are returned).  Returns (results[RESULT_DTYPE], I', Q').
    The reference writes fftw_wisdom.dat / hashtable.txt into the CWD (wsprd.c:835,843): run in a scratch dir."""
    options = options or default_options()
    # The spectrogram loop reads samples up to index 512 * floor(n / 512) + 255 whatever n is (wsprd.c:516,536-541), i.e.
    # up to 255 floats PAST the end when n % 512 < 256.  The reference's callers always hand over full-size, zero-tailed
    # buffers (rtlsdr_wsprd.c:285-288,575-589), so that is what the decoder gets here too: n samples in a zero-padded buffer.
    n_in = int(np.asarray(idat).shape[0])
    i = np.zeros(n_in + 512, np.float32)
    q = np.zeros(n_in + 512, np.float32)
    i[:n_in] = idat
    q[:n_in] = qdat
    out = (DecoderResults * MAX_UNIQUES)()
    n = C.c_int(0)
    old = os.getcwd()
    scratch = tempfile.mkdtemp(prefix="wspr_oracle_") if cwd_scratch else None
    try:
        if scratch:
            os.chdir(scratch)
        lib.wspr_decode(i.ctypes.data_as(C.POINTER(C.c_float)), q.ctypes.data_as(C.POINTER(C.c_float)),
                        n_in, options, out, C.byref(n))
    finally:
        os.chdir(old)
        if scratch:
            shutil.rmtree(scratch, ignore_errors=True)
    arr = np.frombuffer(bytes(out), dtype=RESULT_DTYPE, count=MAX_UNIQUES)[: n.value].copy()
    return arr, i[:n_in].copy(), q[:n_in].copy()


def spot_line(r):
    """The reference's print contract, rtlsdr_wsprd.c:693-700 (without the 'Spot : ' prefix)."""
    return "%6.2f %6.2f %10.6f %2d %7s %6s %2s" % (
        r["snr"], r["dt"], r["freq"], int(r["drift"]), r["call"].decode(), r["loc"].decode(), r["pwr"].decode())


def read_iq_file(path, nmax=NSAMP):
    """numpy restatement of readRawIQfile(), rtlsdr_wsprd.c:555-592: interleaved f32 (I, -Q), peak-normalised
    so that max(|I|,|Q|) = 0.5; the scale is (float)(0.5 / (double)max) and the multiply is in float."""
    raw = np.fromfile(path, dtype="<f4", count=2 * nmax)
    n = raw.shape[0] // 2
    i = raw[0:2 * n:2].copy()
    q = (-raw[1:2 * n:2]).astype(np.float32)
    return normalise_half(i, q)


def normalise_half(i, q):
    """rtlsdr_wsprd.c:291-305 / :575-589."""
    m = np.float32(1e-24)
    if i.size:
        m = max(m, np.float32(np.max(np.abs(i))), np.float32(np.max(np.abs(q))))
    scale = np.float32(0.5 / float(m))
    return (i * scale).astype(np.float32), (q * scale).astype(np.float32)


class RefFrontend:
    """A fresh instance of the reference's rtlsdr_callback (function-static state => one .so copy per stream)."""

    def __init__(self):
        src = os.path.join(HERE, "_ref", "libfrontend_ref.so")
        if not os.path.exists(src):
            raise FileNotFoundError(src)
        self._dir = tempfile.mkdtemp(prefix="wspr_fe_")
        dst = os.path.join(self._dir, "fe_%d.so" % id(self))
        shutil.copy(src, dst)
        self.lib = C.CDLL(dst)
        self.lib.ref_frontend_push.argtypes = [C.POINTER(C.c_ubyte), C.c_uint32]
        self.lib.ref_frontend_count.restype = C.c_uint32
        self.lib.ref_frontend_read.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_uint32]

    def push(self, raw_u8, chunk=65536):
        """raw_u8: uint8[2*nsamples] interleaved; fed in DEFAULT_BUF_LENGTH chunks (rtlsdr_wsprd.c:42,256)."""
        buf = np.ascontiguousarray(raw_u8, dtype=np.uint8).copy()   # the callback mutates its buffer
        for off in range(0, buf.shape[0], chunk):
            part = buf[off:off + chunk]
            self.lib.ref_frontend_push(part.ctypes.data_as(C.POINTER(C.c_ubyte)), part.shape[0])

    def read(self):
        n = self.lib.ref_frontend_count()
        i = np.zeros(n, np.float32)
        q = np.zeros(n, np.float32)
        self.lib.ref_frontend_read(i.ctypes.data_as(C.POINTER(C.c_float)), q.ctypes.data_as(C.POINTER(C.c_float)), n)
        return i, q

    def __del__(self):
        shutil.rmtree(getattr(self, "_dir", ""), ignore_errors=True)
