"""Host side of the decode path: ctypes over the C ABI of libwsprd_b200.so (include/wspr_b200.h).

Names and argument meaning follow the reference (wsprd/wsprd.h:44-111, rtlsdr_wsprd.c:447-474,555-701): a caller of
the reference's ``wspr_decode(idat, qdat, samples, options, decodes, &n)`` finds the same call here, plus the batch
forms.  PyTorch is optional plumbing only (device-resident inputs via ``data_ptr()``); nothing here computes.
"""
import ctypes as C
import os
import struct
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NSAMP = 45000            # SIGNAL_LENGHT * SIGNAL_SAMPLE_RATE, rtlsdr_wsprd.c:37-38
MAX_UNIQUES = 100        # wsprd/wsprd.h:41
MAX_CANDIDATES = 200     # wsprd/wsprd.h:40
WSPR_ERR_CUDA, WSPR_ERR_ARG = -2, -3


class WsprCudaError(RuntimeError):
    """The CUDA library is missing, failed to load, or a CUDA call failed.  There is no CPU fallback."""


class DecoderOptions(C.Structure):
    """struct decoder_options, wsprd/wsprd.h:44-52 (40 bytes, passed BY VALUE)."""
    _fields_ = [("freq", C.c_int), ("rcall", C.c_char * 13), ("rloc", C.c_char * 7), ("quickmode", C.c_int),
                ("usehashtable", C.c_int), ("npasses", C.c_int), ("subtraction", C.c_int)]


class DecoderResults(C.Structure):
    """struct decoder_results, wsprd/wsprd.h:62-74 (80 bytes)."""
    _fields_ = [("freq", C.c_double), ("sync", C.c_float), ("snr", C.c_float), ("dt", C.c_float), ("drift", C.c_float),
                ("jitter", C.c_int), ("message", C.c_char * 23), ("call", C.c_char * 13), ("loc", C.c_char * 7),
                ("pwr", C.c_char * 3), ("cycles", C.c_int)]


assert C.sizeof(DecoderOptions) == 40 and C.sizeof(DecoderResults) == 80

RESULT_DTYPE = np.dtype({
    "names": ["freq", "sync", "snr", "dt", "drift", "jitter", "message", "call", "loc", "pwr", "cycles"],
    "formats": ["<f8", "<f4", "<f4", "<f4", "<f4", "<i4", "S23", "S13", "S7", "S3", "<i4"],
    "offsets": [0, 8, 12, 16, 20, 24, 28, 51, 64, 71, 76], "itemsize": 80})
CAND_DTYPE = np.dtype([("freq", "<f4"), ("snr", "<f4"), ("shift", "<i4"), ("drift", "<f4"), ("sync", "<f4")])


def default_options(freq=144489000, npasses=2, subtraction=1, quickmode=0, usehashtable=0):
    """initDecoder_options(), rtlsdr_wsprd.c:356-362."""
    o = DecoderOptions()
    o.freq = int(freq)
    o.rcall, o.rloc = b"A1XYZ", b"AB12CD"
    o.quickmode, o.usehashtable, o.npasses, o.subtraction = quickmode, usehashtable, npasses, subtraction
    return o


def library_path():
    """The in-tree build; WSPR_B200_LIB selects another build of the same library (A/B measurements)."""
    return os.environ.get("WSPR_B200_LIB") or os.path.join(HERE, "libwsprd_b200.so")


def build_library(force=False):
    """Compile csrc/*.cu into libwsprd_b200.so with nvcc for sm_100a (csrc/Makefile)."""
    args = ["make", "-s", "-C", os.path.join(HERE, "csrc"), "-j4"]
    if force:
        args.append("-B")
    subprocess.run(args, check=True)
    return library_path()


_lib = None


def library():
    """Load libwsprd_b200.so (raises WsprCudaError if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise WsprCudaError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(nvcc, sm_100a); there is no CPU fallback" % path)
    try:
        lib = C.CDLL(path)
    except OSError as e:
        raise WsprCudaError("cannot load %s: %s" % (path, e))
    fp, ip, up, vp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_ubyte), C.c_void_p
    lib.wspr_decode.restype = C.c_int
    lib.wspr_decode.argtypes = [fp, fp, C.c_int, DecoderOptions, C.POINTER(DecoderResults), ip]
    lib.wspr_decode_batch.restype = C.c_int
    lib.wspr_decode_batch.argtypes = [vp, vp, C.c_int, C.c_int, DecoderOptions, vp, vp, C.c_int]
    lib.wspr_ctx_create.restype = vp
    lib.wspr_ctx_create.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.wspr_ctx_destroy.restype = None
    lib.wspr_ctx_destroy.argtypes = [vp]
    lib.wspr_last_error.restype = C.c_char_p
    lib.wspr_frontend_last_error.restype = C.c_char_p
    lib.wspr_ctx_upload.argtypes = [vp, vp, vp, C.c_int]
    lib.wspr_ctx_upload_device.argtypes = [vp, vp, vp, C.c_int, C.c_int]
    lib.wspr_ctx_normalise.argtypes = [vp]
    if hasattr(lib, "wspr_ctx_decimate"):
        lib.wspr_ctx_decimate.argtypes = [vp, vp, C.c_int, C.c_size_t, C.c_size_t]
    lib.wspr_ctx_decode.argtypes = [vp, DecoderOptions]
    lib.wspr_ctx_download.argtypes = [vp, vp, vp, vp, vp]
    lib.wspr_ctx_last_decode_ms.restype = C.c_float
    lib.wspr_ctx_last_decode_ms.argtypes = [vp]
    lib.wspr_ctx_time_kernels.argtypes = [vp, C.c_int]
    lib.wspr_ctx_last_rounds.argtypes = [vp]
    lib.wspr_ctx_last_deferred.argtypes = [vp]
    lib.wspr_ctx_last_stats.argtypes = [vp, vp]
    lib.wspr_ctx_stream.restype = vp
    lib.wspr_ctx_stream.argtypes = [vp]
    lib.wspr_ctx_last_sync_ms.restype = C.c_float
    lib.wspr_ctx_last_sync_ms.argtypes = [vp]
    lib.wspr_ctx_last_sync_launches.argtypes = [vp]
    lib.wspr_ctx_last_sync_cells.restype = C.c_double
    lib.wspr_ctx_last_sync_cells.argtypes = [vp]
    lib.wspr_kernel_launches.restype = C.c_ulonglong
    lib.wspr_ctx_spectrogram.argtypes = [vp, vp]
    lib.wspr_ctx_candidates.argtypes = [vp, C.c_int, vp, vp, vp]
    if hasattr(lib, "wspr_fano_stats"):
        lib.wspr_fano_stats.argtypes = [C.c_int, vp, C.c_int]
    lib.wspr_fano_batch.argtypes = [vp, C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_int, vp, vp, vp, vp, vp, vp]
    lib.wspr_decimate_batch.argtypes = [vp, C.c_int, C.c_size_t, vp, vp, C.c_int, C.c_int]
    lib.wspr_decimate_device.argtypes = [vp, C.c_int, C.c_size_t, C.c_size_t, vp, vp, C.c_int, C.c_int, C.c_int]
    lib.wspr_decimate_last_ms.restype = C.c_float
    if hasattr(lib, "wspr_frontend_create"):                  # (absent from older builds selected through WSPR_B200_LIB)
        lib.wspr_frontend_create.restype = vp
        lib.wspr_frontend_create.argtypes = [C.c_int, C.c_int, C.c_int]
        lib.wspr_frontend_destroy.restype = None
        lib.wspr_frontend_destroy.argtypes = [vp]
        lib.wspr_frontend_push.argtypes = [vp, vp, C.c_size_t, C.c_uint32]
        lib.wspr_frontend_samples.argtypes = [vp]
        lib.wspr_frontend_swap.argtypes = [vp]
        lib.wspr_frontend_read.argtypes = [vp, vp, vp]
        lib.wspr_frontend_slot_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), ip]
    lib.sync_and_demodulate.restype = None
    lib.sync_and_demodulate.argtypes = [fp, fp, C.c_long, up, fp, C.c_int, C.c_int, C.c_float, ip, C.c_int, C.c_int,
                                        C.c_int, fp, C.c_int, fp, C.c_int]
    lib.subtract_signal2.restype = None
    lib.subtract_signal2.argtypes = [fp, fp, C.c_long, C.c_float, C.c_int, C.c_float, up]
    _lib = lib
    return lib


def _check(rc, what, frontend=False):
    if rc is not None and rc < 0:
        lib = library()
        msg = (lib.wspr_frontend_last_error() if frontend else lib.wspr_last_error()) or b""
        raise WsprCudaError("%s failed (%d): %s" % (what, rc, msg.decode(errors="replace")))
    return rc


def kernel_launches():
    """Kernels launched by this process through the library so far."""
    return int(library().wspr_kernel_launches())


def _f32c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def wspr_decode(idat, qdat, samples=None, options=None):
    """The reference entry point (wsprd/wsprd.h:106-111) for ONE capture.

    idat/qdat: float32[samples]; like the reference, both arrays are MUTATED in place when they are contiguous
    float32 numpy arrays (signal subtraction).  Returns the spots as a RESULT_DTYPE array (``decodes[:n_results]``)."""
    lib = library()
    options = options or default_options()
    if not (isinstance(idat, np.ndarray) and idat.dtype == np.float32 and idat.flags.c_contiguous):
        idat = _f32c(idat)
    if not (isinstance(qdat, np.ndarray) and qdat.dtype == np.float32 and qdat.flags.c_contiguous):
        qdat = _f32c(qdat)
    samples = int(idat.shape[0] if samples is None else samples)
    out = (DecoderResults * MAX_UNIQUES)()
    n = C.c_int(-1)
    fp = C.POINTER(C.c_float)
    lib.wspr_decode(idat.ctypes.data_as(fp), qdat.ctypes.data_as(fp), samples, options, out, C.byref(n))
    err = lib.wspr_last_error() or b""
    if n.value == 0 and err:
        # the reference ABI has no error channel (always returns 0); surface CUDA failures to Python callers
        raise WsprCudaError("wspr_decode: %s" % err.decode(errors="replace"))
    return np.frombuffer(bytes(out), dtype=RESULT_DTYPE, count=MAX_UNIQUES)[: max(n.value, 0)].copy()


class BatchDecoder:
    """A device context (wspr_ctx): all buffers for up to ``max_captures`` captures of ``samples`` samples on one GPU.

    upload() -> [normalise()] -> decode() -> download() ; the steps are separate so that a caller can keep captures
    resident in HBM and time the decode alone."""

    def __init__(self, max_captures, samples=NSAMP, device=-1):
        self.lib = library()
        self.max_captures, self.samples, self.device = int(max_captures), int(samples), int(device)
        self.ctx = self.lib.wspr_ctx_create(self.device, self.max_captures, self.samples)
        if not self.ctx:
            raise WsprCudaError("wspr_ctx_create: %s" % (self.lib.wspr_last_error() or b"").decode(errors="replace"))
        self.ncap = 0

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.wspr_ctx_destroy(self.ctx)
            self.ctx = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def upload(self, I, Q):
        """I, Q: float32[ncap, samples] host arrays (pinned memory makes the copy asynchronous)."""
        I, Q = _f32c(I), _f32c(Q)
        if I.ndim == 1:
            I, Q = I[None, :], Q[None, :]
        assert I.shape == Q.shape and I.shape[1] == self.samples, (I.shape, Q.shape, self.samples)
        self.ncap = I.shape[0]
        _check(self.lib.wspr_ctx_upload(self.ctx, I.ctypes.data, Q.ctypes.data, self.ncap), "wspr_ctx_upload")

    def upload_ptr(self, i_ptr, q_ptr, ncap):
        """Host pointers (e.g. torch pinned tensors' data_ptr()) to [ncap][samples] float32 planes."""
        self.ncap = int(ncap)
        _check(self.lib.wspr_ctx_upload(self.ctx, int(i_ptr), int(q_ptr), self.ncap), "wspr_ctx_upload")

    def upload_device(self, i_ptr, q_ptr, ncap, row_stride=None):
        """Device pointers to [ncap][row_stride] float32 planes already in HBM (copied device to device)."""
        self.ncap = int(ncap)
        _check(self.lib.wspr_ctx_upload_device(self.ctx, int(i_ptr), int(q_ptr), self.ncap,
                                               int(row_stride or self.samples)), "wspr_ctx_upload_device")

    def decimate(self, raw_ptr, nstreams, n_iq, stream_stride_bytes):
        """Raw u8 IQ streams on the device (16-byte aligned, stream_stride_bytes apart) through the front end straight into
        this context (rtlsdr_wsprd.c:126-244,285-288); asynchronous.  Follow with normalise() and decode()."""
        self.ncap = int(nstreams)
        return _check(self.lib.wspr_ctx_decimate(self.ctx, int(raw_ptr), int(nstreams), int(n_iq), int(stream_stride_bytes)),
                      "wspr_ctx_decimate")

    def normalise(self):
        _check(self.lib.wspr_ctx_normalise(self.ctx), "wspr_ctx_normalise")

    def decode(self, options=None):
        _check(self.lib.wspr_ctx_decode(self.ctx, options or default_options()), "wspr_ctx_decode")
        return float(self.lib.wspr_ctx_last_decode_ms(self.ctx))

    def download(self, samples=False, out=None, n_out=None):
        """Returns (spots[ncap, 100] RESULT_DTYPE, n_results[ncap]) (+ the post-subtraction I, Q when samples=True).
        ``out``/``n_out`` may be preallocated (e.g. views of pinned memory)."""
        spots = out if out is not None else np.zeros((self.ncap, MAX_UNIQUES), RESULT_DTYPE)
        n = n_out if n_out is not None else np.zeros(self.ncap, np.int32)
        I = Q = None
        ip = qp = None
        if samples:
            I = np.zeros((self.ncap, self.samples), np.float32)
            Q = np.zeros((self.ncap, self.samples), np.float32)
            ip, qp = I.ctypes.data, Q.ctypes.data
        _check(self.lib.wspr_ctx_download(self.ctx, spots.ctypes.data, n.ctypes.data, ip, qp), "wspr_ctx_download")
        return (spots, n, I, Q) if samples else (spots, n)

    def spectrogram(self):
        """ps[ncap, 512, blocks] of pass 0 in the reference's layout (wsprd.c:517,536-553)."""
        blocks = 4 * (self.samples // 512) - 1
        ps = np.zeros((self.ncap, 512, blocks), np.float32)
        _check(self.lib.wspr_ctx_spectrogram(self.ctx, ps.ctypes.data), "wspr_ctx_spectrogram")
        return ps

    def candidates(self, maxdrift=4, want_smspec=False):
        """Candidate list after the coarse sync of pass 0 (wsprd.c:555-678): (cands[ncap, 200], npk[ncap][, smspec])."""
        cands = np.zeros((self.ncap, MAX_CANDIDATES), CAND_DTYPE)
        npk = np.zeros(self.ncap, np.int32)
        sm = np.zeros((self.ncap, 411), np.float32) if want_smspec else None
        _check(self.lib.wspr_ctx_candidates(self.ctx, int(maxdrift), cands.ctypes.data, npk.ctypes.data,
                                            sm.ctypes.data if want_smspec else None), "wspr_ctx_candidates")
        return (cands, npk, sm) if want_smspec else (cands, npk)

    def stream(self):
        """cudaStream_t (as int) the context issues its work on, e.g. for torch.cuda.ExternalStream."""
        return int(self.lib.wspr_ctx_stream(self.ctx) or 0)

    def schedule_stats(self):
        """(rounds, deferred candidates) of the last decode."""
        st = (C.c_int * 8)()
        self.lib.wspr_ctx_last_stats(self.ctx, st)
        return int(self.lib.wspr_ctx_last_rounds(self.ctx)), int(self.lib.wspr_ctx_last_deferred(self.ctx)), list(st)[:3]

    def time_kernels(self, on=True):
        self.lib.wspr_ctx_time_kernels(self.ctx, int(bool(on)))

    def sync_kernel_stats(self):
        """(ms, launches, cells) of the mode-0 sync correlation kernel during the last decode (needs time_kernels)."""
        return (float(self.lib.wspr_ctx_last_sync_ms(self.ctx)), int(self.lib.wspr_ctx_last_sync_launches(self.ctx)),
                float(self.lib.wspr_ctx_last_sync_cells(self.ctx)))


class PipelinedDecoder:
    """`depth` BatchDecoder contexts, each driven by its own host thread, so consecutive batches overlap on the GPU.

    A decode ends with a tail in which only the few captures that own a hopeless candidate are still busy (one Fano
    time-out is 810 000 strictly sequential cycles); with several batches in flight the GPU works on the next batch's
    bulk while the previous one drains, and host<->device copies of one batch hide behind the kernels of another.
    submit(fn) runs fn(decoder) on a free context and returns a concurrent.futures.Future."""

    def __init__(self, depth, max_captures, samples=NSAMP, device=-1):
        import concurrent.futures as cf
        import queue
        self.depth = int(depth)
        self.decoders = [BatchDecoder(max_captures, samples, device) for _ in range(self.depth)]
        self._free = queue.Queue()
        for d in self.decoders:
            self._free.put(d)
        self._pool = cf.ThreadPoolExecutor(max_workers=self.depth)

    def submit(self, fn):
        def run():
            d = self._free.get()
            try:
                return fn(d)
            finally:
                self._free.put(d)
        return self._pool.submit(run)

    def decode_async(self, I, Q, options=None):
        """Upload host arrays, decode, download: Future of (spots[ncap, 100], n_results[ncap])."""
        def job(d):
            d.upload(I, Q)
            d.decode(options)
            return d.download()
        return self.submit(job)

    def close(self):
        self._pool.shutdown(wait=True)
        for d in self.decoders:
            d.close()
        self.decoders = []

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def decode_batch(I, Q, options=None, device=-1):
    """One-shot batch decode of host arrays I, Q: float32[ncap, samples] -> list of RESULT_DTYPE arrays."""
    lib = library()
    I, Q = _f32c(I), _f32c(Q)
    if I.ndim == 1:
        I, Q = I[None, :], Q[None, :]
    ncap, samples = I.shape
    spots = np.zeros((ncap, MAX_UNIQUES), RESULT_DTYPE)
    n = np.zeros(ncap, np.int32)
    _check(lib.wspr_decode_batch(I.ctypes.data, Q.ctypes.data, ncap, samples, options or default_options(),
                                 spots.ctypes.data, n.ctypes.data, int(device)), "wspr_decode_batch")
    return [spots[c, : n[c]].copy() for c in range(ncap)]


def fano_batch(symbols, delta=60, maxcycles=10000, stop_after=0, solo=False):
    """The Fano kernel (K5) on soft symbols uint8[n, 162] (deinterleaved).  Returns dict(rc, metric, cycles, maxnp, data[n,12])."""
    sym = np.ascontiguousarray(symbols, dtype=np.uint8).reshape(-1, 162)
    n = sym.shape[0]
    out = dict(rc=np.zeros(n, np.int32), metric=np.zeros(n, np.uint32), cycles=np.zeros(n, np.uint32),
               maxnp=np.zeros(n, np.uint32), data=np.zeros((n, 12), np.uint8), clocks=np.zeros(n, np.uint64))
    _check(library().wspr_fano_batch(sym.ctypes.data, n, int(delta), int(maxcycles), int(stop_after), int(solo),
                                     out["rc"].ctypes.data, out["metric"].ctypes.data, out["cycles"].ctypes.data,
                                     out["maxnp"].ctypes.data, out["data"].ctypes.data, out["clocks"].ctypes.data),
           "wspr_fano_batch")
    return out


def fano_pool_stats(device=-1, reset=False):
    """Counters of the device's pool of Fano worker warps (wspr_fano_stats): dict with pool, fano_sms, lane utilisation..."""
    out = (C.c_ulonglong * 8)()
    _check(library().wspr_fano_stats(int(device), out, int(bool(reset))), "wspr_fano_stats")
    v = [int(x) for x in out]
    return {"pool_warps": v[0], "fano_sms": v[1], "warp_periods": v[2], "lane_utilisation": round(v[3] / (32.0 * v[2]), 4) if v[2] else None,
            "attempts_run": v[4], "attempts_dropped": v[5], "worker_warps_started": v[6],
            "overflow_share": round(v[7] / v[2], 4) if v[2] else None}


def decimate_batch(raw, n_iq=None, max_out=NSAMP, device=-1):
    """rtlsdr_callback for whole streams (rtlsdr_wsprd.c:126-244).  raw: uint8[nstreams, 2*n_iq] interleaved (I,Q).
    Returns (I, Q) float32[nstreams, max_out] (zero padded) and the number of outputs per stream."""
    lib = library()
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    if raw.ndim == 1:
        raw = raw[None, :]
    nstreams = raw.shape[0]
    n_iq = int(raw.shape[1] // 2 if n_iq is None else n_iq)
    I = np.zeros((nstreams, max_out), np.float32)
    Q = np.zeros((nstreams, max_out), np.float32)
    n = _check(lib.wspr_decimate_batch(raw.ctypes.data, nstreams, n_iq, I.ctypes.data, Q.ctypes.data, int(max_out),
                                       int(device)), "wspr_decimate_batch", frontend=True)
    return I, Q, n


def decimate_device(raw_ptr, nstreams, n_iq, stream_stride_bytes, i_ptr, q_ptr, out_stride, max_out=NSAMP, device=-1):
    """Device-resident form: raw_ptr -> [nstreams][stream_stride_bytes] u8 (16-byte aligned), outputs [nstreams][out_stride]."""
    n = _check(library().wspr_decimate_device(int(raw_ptr), int(nstreams), int(n_iq), int(stream_stride_bytes), int(i_ptr),
                                              int(q_ptr), int(out_stride), int(max_out), int(device)),
               "wspr_decimate_device", frontend=True)
    return n, float(library().wspr_decimate_last_ms())


class FrontEnd:
    """Streaming front end for `nstreams` receivers in lockstep: the state rtlsdr_callback keeps between calls
    (rtlsdr_wsprd.c:126-244) plus the reference's double buffer (rtlsdr_wsprd.c:80-87,1181-1183).

    push(chunk) per received buffer; at every slot boundary swap() then read() (or hand_off(decoder)) for the slot that
    just ended -- what the daemon's main loop and decoder thread do."""

    def __init__(self, nstreams=1, slot_samples=NSAMP, device=-1):
        self.lib = library()
        self.nstreams, self.slot_samples = int(nstreams), int(slot_samples)
        self.fe = self.lib.wspr_frontend_create(int(device), self.nstreams, self.slot_samples)
        if not self.fe:
            raise WsprCudaError("wspr_frontend_create: %s" % (self.lib.wspr_frontend_last_error() or b"").decode(errors="replace"))

    def close(self):
        if getattr(self, "fe", None):
            self.lib.wspr_frontend_destroy(self.fe)
            self.fe = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def push(self, raw):
        """raw: uint8[nstreams, nbytes] (or uint8[nbytes] for one stream), interleaved (I,Q); nbytes % 8 == 0."""
        raw = np.ascontiguousarray(raw, dtype=np.uint8)
        if raw.ndim == 1:
            raw = raw[None, :]
        assert raw.shape[0] == self.nstreams, (raw.shape, self.nstreams)
        return _check(self.lib.wspr_frontend_push(self.fe, raw.ctypes.data, raw.shape[1], raw.shape[1]),
                      "wspr_frontend_push", frontend=True)

    def samples(self):
        return int(self.lib.wspr_frontend_samples(self.fe))

    def swap(self):
        return _check(self.lib.wspr_frontend_swap(self.fe), "wspr_frontend_swap", frontend=True)

    def read(self):
        """(I, Q, n): the slot that ended at the last swap, float32[nstreams, slot_samples], zero beyond n."""
        I = np.zeros((self.nstreams, self.slot_samples), np.float32)
        Q = np.zeros((self.nstreams, self.slot_samples), np.float32)
        n = _check(self.lib.wspr_frontend_read(self.fe, I.ctypes.data, Q.ctypes.data), "wspr_frontend_read", frontend=True)
        return I, Q, n

    def hand_off(self, decoder):
        """Give the ended slot to a BatchDecoder without leaving the device, peak-normalised like decoder() does
        (rtlsdr_wsprd.c:285-305).  Returns the samples the slot holds."""
        di, dq, stride = C.c_void_p(), C.c_void_p(), C.c_int()
        n = _check(self.lib.wspr_frontend_slot_device(self.fe, C.byref(di), C.byref(dq), C.byref(stride)),
                   "wspr_frontend_slot_device", frontend=True)
        assert decoder.samples <= stride.value
        decoder.upload_device(di.value, dq.value, self.nstreams, stride.value)
        decoder.normalise()
        return n


# ---- host-side formats either side of the path --------------------------------------------------
def spot_line(r):
    """Print contract of decodeRecordedFile (rtlsdr_wsprd.c:693-700), without the 'Spot : ' prefix."""
    return "%6.2f %6.2f %10.6f %2d %7s %6s %2s" % (r["snr"], r["dt"], r["freq"], int(r["drift"]),
                                                   r["call"].decode(), r["loc"].decode(), r["pwr"].decode())


def print_spots_lines(results, when=None):
    """printSpots (rtlsdr_wsprd.c:447-474): the daemon's per-slot report.  `when`: time.struct_time (UTC) of the slot."""
    import time as _time
    g = when or _time.gmtime()
    stamp = "%04d-%02d-%02d %02d:%02dz" % (g.tm_year, g.tm_mon, g.tm_mday, g.tm_hour, g.tm_min)
    if len(results) == 0:
        return ["No spot " + stamp]
    return ["Spot :  %s %6.2f %6.2f %10.6f %2d %7s %6s %2s" % (stamp, r["snr"], r["dt"], r["freq"], int(r["drift"]),
                                                             r["call"].decode(), r["loc"].decode(), r["pwr"].decode())
            for r in results]


def wsprnet_urls(results, rcall, rloc, dialfreq, when=None, version="rtlsdr-056"):
    """The report URLs postSpots would request (rtlsdr_wsprd.c:366-444); nothing is sent from here.  rcall/rloc are
    percent-escaped like curl_easy_escape does; `version` = the reference's wsprnet_app_version (rtlsdr_wsprd.c:122)."""
    import time as _time
    from urllib.parse import quote
    g = when or _time.gmtime()
    rc, rl = quote(rcall, safe="-._~"), quote(rloc, safe="-._~")
    if len(results) == 0:
        return ["https://wsprnet.org/post?function=wsprstat&rcall=%s&rgrid=%s&rqrg=%.6f&tpct=%.2f&tqrg=%.6f&dbm=%d&version=%s&mode=2"
                % (rc, rl, dialfreq / 1e6, 0.0, dialfreq / 1e6, 0, version)]
    return ["https://wsprnet.org/post?function=wspr&rcall=%s&rgrid=%s&rqrg=%.6f&date=%02d%02d%02d&time=%02d%02d&sig=%.0f&dt=%.1f"
            "&tqrg=%.6f&tcall=%s&tgrid=%s&dbm=%s&version=%s&mode=2"
            % (rc, rl, r["freq"], g.tm_year % 100, g.tm_mon, g.tm_mday, g.tm_hour, g.tm_min, r["snr"], r["dt"], r["freq"],
               r["call"].decode(), r["loc"].decode(), r["pwr"].decode(), version) for r in results]


def normalise_half(i, q):
    """Peak normalisation to 0.5 (rtlsdr_wsprd.c:291-305, :575-589): scale = (float)(0.5 / max), float multiply."""
    m = np.float32(1e-24)
    if i.size:
        m = max(m, np.float32(np.max(np.abs(i))), np.float32(np.max(np.abs(q))))
    scale = np.float32(0.5 / float(m))
    return (i * scale).astype(np.float32), (q * scale).astype(np.float32)


def read_iq_file(path, nmax=NSAMP):
    """readRawIQfile (rtlsdr_wsprd.c:555-592): interleaved f32 (I, -Q), peak-normalised to 0.5."""
    raw = np.fromfile(path, dtype="<f4", count=2 * nmax)
    n = raw.shape[0] // 2
    return normalise_half(raw[0:2 * n:2].copy(), (-raw[1:2 * n:2]).astype(np.float32))


def read_c2_file(path, nmax=NSAMP):
    """readC2file (rtlsdr_wsprd.c:619-667): 14-byte name, int type, double dial frequency, then the .iq body.
    Returns (I, Q, dialfreq)."""
    with open(path, "rb") as f:
        head = f.read(14 + 4 + 8)
        _name, _type, freq = head[:14], struct.unpack("<i", head[14:18])[0], struct.unpack("<d", head[18:26])[0]
        raw = np.frombuffer(f.read(8 * nmax), dtype="<f4")
    n = raw.shape[0] // 2
    i, q = normalise_half(raw[0:2 * n:2].copy(), (-raw[1:2 * n:2]).astype(np.float32))
    return i, q, freq


def write_iq_file(path, i, q):
    """writeRawIQfile (rtlsdr_wsprd.c:595-617): interleaved f32 (I, -Q)."""
    buf = np.empty(2 * len(i), "<f4")
    buf[0::2] = i
    buf[1::2] = -np.asarray(q, np.float32)
    buf.tofile(path)
    return len(i)


def load_capture_files(paths, nmax=NSAMP, pinned=False):
    """Batch loader for the on-disk formats either side of the path: a directory or a list of `.iq` / `.c2`
    recordings -> the planar [n][45000] float32 batch that BatchDecoder.upload takes.  Every file goes through the
    reference's own reader semantics (readRawIQfile / readC2file, rtlsdr_wsprd.c:555-667: Q negated, peak-normalised to
    0.5); short recordings are zero-padded like the daemon's hand-off (rtlsdr_wsprd.c:285-288).
    Returns (I, Q, dialfreq[n] (0 for .iq files), names).  pinned=True puts I and Q in page-locked memory (torch), so that
    the upload is an asynchronous DMA."""
    if isinstance(paths, (str, os.PathLike)):
        root = os.fspath(paths)
        if os.path.isdir(root):
            paths = sorted(os.path.join(root, f) for f in os.listdir(root) if f.lower().endswith((".iq", ".c2")))
        else:
            paths = [root]
    paths = [os.fspath(x) for x in paths]
    n = len(paths)
    if pinned:
        import torch
        ti = torch.zeros((n, nmax), dtype=torch.float32).pin_memory()
        tq = torch.zeros((n, nmax), dtype=torch.float32).pin_memory()
        I, Q = ti.numpy(), tq.numpy()
    else:
        I, Q = np.zeros((n, nmax), np.float32), np.zeros((n, nmax), np.float32)
    freq = np.zeros(n, np.float64)
    for k, path in enumerate(paths):
        if path.lower().endswith(".c2"):
            i, q, freq[k] = read_c2_file(path, nmax)
        else:
            i, q = read_iq_file(path, nmax)
        I[k, : len(i)], Q[k, : len(q)] = i, q
    return I, Q, freq, [os.path.basename(x) for x in paths]


def write_c2_file(path, i, q, dialfreq, name=b"000000_0000.c2", ftype=2):
    """The .c2 layout readC2file expects (rtlsdr_wsprd.c:619-640): 14-byte name, int type, double dial frequency, then
    interleaved f32 (I, -Q)."""
    buf = np.empty(2 * len(i), "<f4")
    buf[0::2] = i
    buf[1::2] = -np.asarray(q, np.float32)
    with open(path, "wb") as f:
        f.write(bytes(name)[:14].ljust(14, b"\0"))
        f.write(struct.pack("<i", int(ftype)))
        f.write(struct.pack("<d", float(dialfreq)))
        buf.tofile(f)
    return len(i)
