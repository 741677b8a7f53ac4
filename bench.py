#!/usr/bin/env python3
"""bench.py -- throughput of the WSPR decode hot path on B200 (metric and configs of BASELINE.json).

  python bench.py --gpus N --steps K --warmup W             our CUDA path (one process per GPU under torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K ...    the reference's own CPU code (oracle/_ref) on the host cores
  python bench.py --workload config2|config3|config4|config5  (default config3, the configuration the metric is quoted on)

A step = one pass of the decode path (both passes, subtraction on: reference defaults rtlsdr_wsprd.c:357-362) over one batch
per GPU.  Prints ONE JSON line on rank 0.  `value` = units/s with the inputs resident in HBM (device-timed, max over ranks);
`e2e` = the same through the C ABI from pinned host buffers (H2D of the inputs and D2H of the spot records inside the timed
region).  oracle/ is used here only as the checker: the cpu_baseline / parity leg and the --impl reference arm.
  config2  1 024 single-signal captures at -20 dB              config3  4 096 captures x 10 overlapping signals
  config4  256 raw 2.4 Msps u8 streams -> decimate + decode     config5  100 000 distinct captures, 12 500 per GPU, gathered
"""
import argparse
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from multiprocessing import shared_memory

import numpy as np

# several contexts x (1 main + 4 worker-launch) streams: give them enough hardware queues (must be set before CUDA starts)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "2-min WSPR captures decoded/sec"
UNIT = "captures/s"
NSAMP = 45000
N_IQ = 288_000_000                                   # raw samples of a 120 s stream at 2.4 Msps
WORKLOADS = {
    "config2": "config2: %d captures/GPU x 1 signal at -20 dB, 45000 samples @375 sps, 2 passes + subtraction",
    "config3": "config3: %d captures/GPU x 10 overlapping signals, SNR -28..-10 dB, 45000 samples @375 sps, 2 passes + subtraction",
    "config4": "config4: %d raw streams/GPU x 288e6 u8 IQ pairs @2.4 Msps -> decimate (rtlsdr_callback) + normalise + decode",
    "config5": "config5: %d distinct captures/GPU (config-3 recipe) in contiguous shards, results gathered on rank 0",
}
DEFAULT_UNITS = {"config2": 1024, "config3": 4096, "config4": 256, "config5": 12500}
# algorithmic bytes / flops per unit (SURVEY.md section 8d / DESIGN.md section 5)
BYTES_PER_CAPTURE = 360000 + 80 * 10 + 4
BYTES_PER_STREAM = 2 * N_IQ + 2 * 4 * 44992
SYNC_BYTES_PER_CANDIDATE = (162 * 256 + 256) * 8 + 33 * 162 * 16      # IQ window read + per-(lag,symbol) tone powers written
SYNC_FLOP_PER_CANDIDATE = 33 * 162 * 256 * 32                         # 4 tones x (4 mul + 4 add) per sample, unfused
# dram__bytes_read.sum + dram__bytes_write.sum per candidate / per stream from the ncu --set full captures under profiles/
SYNC_DRAM_BYTES_PER_CANDIDATE = (358.384e6 + 73.219e6) / 1024         # r2_ncu_full_summary.txt (1024 candidates per launch)
FRONTEND_DRAM_BYTES_PER_STREAM = (2305.552e6 + 6.170e6 + 2.895e6) / 4   # r2_ncu_full_summary.txt (4 streams per launch)
# unfused FP32 ceiling: measured multiply + add issue rate of tools/microbench/f32x2_bench.cu (profiles/r2_fp32_peak.txt);
# the nominal figure is 148 SMs x 128 lanes x clock
FP32_PEAK_FILE = os.path.join(ROOT, "profiles", "r2_fp32_peak.json")
FIELDS = ("message", "call", "loc", "pwr", "freq", "snr", "dt", "drift", "sync", "jitter", "cycles")
HARD = 4                                                              # message, call, loc, pwr: BASELINE's bit-for-bit set


# ---- corpus in shared memory (host, seeded; identical arrays go to the GPU path and to the CPU reference) -----------
class SharedPlanes:
    """float32[n, NSAMP] I and Q planes shared with the generator and CPU-decoder processes (all forked after this object
    exists).  POSIX shared memory, attached by name; when /dev/shm has no room for what all ranks of the node will ask for
    (config 5 at eight ranks is 36 GB; a write into a full tmpfs kills the writer with SIGBUS and the pool would wait for
    it forever) the planes are anonymous shared mappings instead, which the forked processes inherit."""

    def __init__(self, n):
        global _INHERITED
        self.n = n
        size = max(1, n) * NSAMP * 4
        ranks_here = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
        try:
            v = os.statvfs("/dev/shm")
            room = v.f_bavail * v.f_frsize
        except OSError:
            room = 0
        self.shm, self.maps = [], []
        if os.environ.get("BENCH_ANON_SHM") == "1" or room < ranks_here * 2 * size + (1 << 30):
            import mmap
            self.maps = [mmap.mmap(-1, size) for _ in range(2)]
            self.I, self.Q = (np.frombuffer(m, np.float32).reshape(max(1, n), NSAMP)[:n] for m in self.maps)
            self.names = None
        else:
            self.shm = [shared_memory.SharedMemory(create=True, size=size) for _ in range(2)]
            self.I, self.Q = (np.ndarray((n, NSAMP), np.float32, buffer=s.buf) for s in self.shm)
            self.names = [s.name for s in self.shm]
        _INHERITED = (self.I, self.Q)

    def close(self):
        global _INHERITED
        self.I = self.Q = _INHERITED = None
        for s in self.shm:
            try:
                s.close()
                s.unlink()
            except (OSError, BufferError):
                pass
        for m in self.maps:
            try:
                m.close()
            except (OSError, BufferError):
                pass


_w = {}                                              # per-process state of pool workers
_INHERITED = None                                    # the parent's planes, as the forked workers see them


def _attach(names, n):
    if names is None:                                # anonymous shared mappings, inherited through fork
        _w["I"], _w["Q"] = _INHERITED
        return
    shm = [shared_memory.SharedMemory(name=x) for x in names]
    _w["shm"] = shm
    _w["I"], _w["Q"] = (np.ndarray((n, NSAMP), np.float32, buffer=s.buf) for s in shm)


def _gen_init(names, n, first):
    """Generator worker: channel symbols come from the library's own get_wspr_channel_symbols (a host function of the
    C ABI, wsprsim_utils.h:3-9; no CUDA call is made in these processes)."""
    import ctypes as C
    import rtlsdr_wsprd_b200 as w
    _attach(names, n)
    lib = w.library()
    cache = {}

    def symbols(msg):
        if msg not in cache:
            sym = (C.c_ubyte * 162)()
            ht, lt = C.create_string_buffer(32768 * 13), C.create_string_buffer(32768 * 5)
            ok = lib.get_wspr_channel_symbols(C.create_string_buffer(msg.encode(), 32), ht, lt, sym)
            assert ok == 1, msg
            cache[msg] = np.frombuffer(bytes(sym), np.uint8).copy()
        return cache[msg]
    _w["symbols"], _w["first"] = symbols, first


def _gen_chunk(args):
    from rtlsdr_wsprd_b200 import corpus
    config, a, b = args                               # local indices [a, b); corpus index = first + local
    for c in range(a, b):
        idx = _w["first"] + c
        plan = corpus.single_signal_plan(idx) if config == 2 else corpus.ten_signal_plan(idx)
        _w["I"][c], _w["Q"][c] = corpus.make_capture(config, idx, plan, _w["symbols"])
    return b - a


def generate_corpus(config, first, planes, workers):
    jobs = [(config, a, min(a + 16, planes.n)) for a in range(0, planes.n, 16)]
    with mp.get_context("fork").Pool(max(1, workers), _gen_init, (planes.names, planes.n, first)) as pool:
        sum(pool.imap_unordered(_gen_chunk, jobs))


# ---- CPU leg: the reference's own decoder, one long-lived process per host core -------------------------------------
def _cpu_init(names, n, kind):
    from oracle import pyoracle as po
    _attach(names, n)
    _w["lib"] = po.ref() if kind == "reference" else po.oracle()
    _w["po"] = po
    os.chdir(tempfile.mkdtemp(prefix="wspr_cpu_"))    # the reference writes fftw_wisdom.dat into the CWD (wsprd.c:835)


def _cpu_chunk(args):
    a, b = args
    po, out = _w["po"], []
    for c in range(a, b):
        r, _, _ = po.decode(_w["lib"], _w["I"][c], _w["Q"][c], cwd_scratch=False)
        out.append(spots_as_tuples(r))
    return a, out


def spots_as_tuples(r):
    return [(x["message"], x["call"], x["loc"], x["pwr"], float(x["freq"]), float(x["snr"]), float(x["dt"]), float(x["drift"]),
             float(x["sync"]), int(x["jitter"]), int(x["cycles"])) for x in r]


class CpuPool:
    """Worker processes that hold the reference library and see the corpus through shared memory: nothing but index
    ranges goes in and spot tuples come out, so a timed decode pays neither process start-up nor pickling of samples."""

    def __init__(self, planes, workers):
        from oracle import pyoracle as po
        self.kind = "reference" if po.ref() is not None else "port"
        po.oracle()                                    # (built once, before forking)
        self.workers = max(1, workers)
        self.pool = mp.get_context("fork").Pool(self.workers, _cpu_init, (planes.names, planes.n, self.kind))
        self.pool.map(_cpu_chunk, [(0, 0)] * self.workers)   # processes up, library loaded

    def decode(self, lo, hi, chunk=4):
        """Decode captures [lo, hi): (results per capture, wall seconds)."""
        jobs = [(a, min(a + chunk, hi)) for a in range(lo, hi, chunk)]
        t0 = time.perf_counter()
        parts = dict(self.pool.imap_unordered(_cpu_chunk, jobs))
        wall = time.perf_counter() - t0
        return [r for a in sorted(parts) for r in parts[a]], wall

    def close(self):
        self.pool.close()
        self.pool.join()


def compare_spot_lists(ref, got):
    """Per-capture comparison in SURVEY section 8d's terms.  ref/got: lists (one per capture) of spot tuples (FIELDS order)."""
    out = dict(captures_checked=len(ref), identical_spot_lists=0, hard_identical=0, spots_reference=0, spots_missing=0,
               spots_extra=0, spots_field_mismatched=0)
    for a, b in zip(ref, got):
        out["spots_reference"] += len(a)
        out["identical_spot_lists"] += int(a == b)
        out["hard_identical"] += int([x[:HARD] for x in a] == [x[:HARD] for x in b])
        left = list(b)
        for x in a:
            hit = next((y for y in left if y[:HARD] == x[:HARD]), None)
            if hit is None:
                out["spots_missing"] += 1
            else:
                left.remove(hit)
                out["spots_field_mismatched"] += int(hit != x)
        out["spots_extra"] += len(left)
    return out


# ---- clocks --------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        load = [x for x in sm if x > 0.5 * mx] or sm
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def load_peaks():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    fp32, fp32_src = 148 * 128 * 1.965e9 / 1e12, "nominal 148 SM x 128 lanes x 1.965 GHz (one unfused op per lane and clock)"
    try:
        m = json.load(open(FP32_PEAK_FILE))
        fp32, fp32_src = float(m["unfused_tflops"]), "measured, %s (profiles/r2_fp32_peak.json)" % m.get("how", "f32x2_bench")
    except (OSError, ValueError, KeyError):
        pass
    return hbm, src, fp32, fp32_src


# ---- our arm: configs 2, 3, 5 (captures in, spot lists out) ---------------------------------------------------------
def run_captures(args):
    import torch
    import rtlsdr_wsprd_b200 as w
    from rtlsdr_wsprd_b200 import sharding
    rank, world, local = sharding.env_rank_world()
    config = 2 if args.workload == "config2" else 3
    total = args.units                                 # captures per GPU (weak scaling: every GPU gets its own contiguous shard)
    first = rank * total
    ncap = min(total, args.batch)                      # captures per decode call
    nbatch = (total + ncap - 1) // ncap
    host_workers = max(1, (os.cpu_count() or 8) // world)
    # every context is driven by a host thread that polls (yielding) for the two counter reads of a round; with fewer host
    # cores per rank than contexts the threads sleep in the driver instead (read once, when the library loads)
    if host_workers < args.depth:
        os.environ.setdefault("WSPR_WAIT", "block")

    planes = SharedPlanes(total)
    t0 = time.perf_counter()
    generate_corpus(config, first, planes, host_workers)
    gen_s = time.perf_counter() - t0
    # the CPU leg's processes are forked BEFORE this process touches CUDA
    nparity = 0
    if args.cpu_sample != 0:
        nparity = total if (world == 1 and args.cpu_sample < 0) else min(total, abs(args.cpu_sample) if args.cpu_sample > 0 else 256)
    cpu = CpuPool(planes, host_workers) if nparity > 0 else None

    rank, world, local = sharding.init_process_group()
    if not torch.cuda.is_available():
        raise w.WsprCudaError("bench.py needs a CUDA device: the decode path has no CPU fallback")
    torch.cuda.set_device(local)
    # pinned host planes (the e2e leg copies from these every step) and a pristine device copy (decode subtracts in place)
    hI = torch.empty((total, NSAMP), dtype=torch.float32).pin_memory()
    hQ = torch.empty((total, NSAMP), dtype=torch.float32).pin_memory()
    hI.numpy()[:] = planes.I
    hQ.numpy()[:] = planes.Q
    resident = min(total, max(ncap, args.resident))    # captures kept in HBM for the resident leg (cycled if fewer than total)
    dI, dQ = hI[:resident].cuda(non_blocking=True), hQ[:resident].cuda(non_blocking=True)
    # `depth` batches in flight (one context + host thread each): the next batch's bulk overlaps the previous one's tail
    pipe = w.PipelinedDecoder(args.depth, ncap, NSAMP, device=local)
    dec = pipe.decoders[0]
    opts = w.default_options()
    # one pinned result area per batch of the shard: every e2e step leaves the whole shard's spot records there
    hs = torch.empty((nbatch, ncap * w.MAX_UNIQUES * 80), dtype=torch.uint8).pin_memory()
    hn = torch.zeros((nbatch, ncap), dtype=torch.int32).pin_memory()
    spots = np.frombuffer(hs.numpy().data, dtype=w.RESULT_DTYPE).reshape(nbatch, ncap, w.MAX_UNIQUES)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def batch_range(b):
        lo = b * ncap
        return lo, min(lo + ncap, total)

    def job_resident(b):
        lo, hi = batch_range(b)
        lo %= resident
        hi = min(lo + (batch_range(b)[1] - batch_range(b)[0]), resident)

        def fn(d):
            d.upload_device(dI[lo:].data_ptr(), dQ[lo:].data_ptr(), hi - lo, NSAMP)
            d.decode(opts)
        return fn

    def job_e2e(b):
        lo, hi = batch_range(b)

        def fn(d):
            d.upload_ptr(hI[lo:].data_ptr(), hQ[lo:].data_ptr(), hi - lo)
            d.decode(opts)
            d.download(out=spots[b], n_out=hn[b].numpy())
        return fn

    def run_steps(job, k):                             # a step = the whole shard = nbatch decode calls
        futs = [pipe.submit(job(b)) for _ in range(k) for b in range(nbatch)]
        for f in futs:
            f.result()

    # ---- resident-input throughput (`value`) ----
    run_steps(job_resident, args.warmup)
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    launches0 = w.kernel_launches()
    w.fano_pool_stats(local, reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()                                   # device idle here (barrier above): default-stream events bracket all streams
    run_steps(job_resident, args.steps)
    torch.cuda.synchronize()
    ev1.record()
    barrier()
    pool_stats = w.fano_pool_stats(local)
    wall_ms = (time.perf_counter() - t0) * 1e3
    dev_ms = ev0.elapsed_time(ev1)
    launches = w.kernel_launches() - launches0
    total_ms = sharding.max_over_ranks(max(dev_ms, 0.0))
    wall_ms = sharding.max_over_ranks(wall_ms)

    # ---- end to end through the C ABI with host buffers ----
    run_steps(job_e2e, max(1, min(args.warmup, (args.depth + nbatch - 1) // nbatch)))
    barrier()
    t0 = time.perf_counter()
    run_steps(job_e2e, args.steps)
    barrier()
    e2e_ms = sharding.max_over_ranks((time.perf_counter() - t0) * 1e3)
    clk = clocks.stop()
    counts = hn.numpy().reshape(-1)[:total] if nbatch * ncap == total else np.concatenate([hn[b].numpy()[: batch_range(b)[1] - batch_range(b)[0]] for b in range(nbatch)])
    nspots = int(counts.sum())

    def gpu_spots(c):
        b, k = divmod(c, ncap)
        return spots_as_tuples(spots[b, k, : hn[b, k]])

    # ---- config 5: the shards' results gathered on rank 0 through the sharding layer (NCCL all_gather of the records) ----
    gathered = None
    if args.workload == "config5":
        flat_s = np.concatenate([spots[b, : batch_range(b)[1] - batch_range(b)[0]] for b in range(nbatch)])
        t0 = time.perf_counter()
        gs, gn = sharding.decode_sharded(lambda lo, hi: (None, None), lambda I, Q: (flat_s, counts.astype(np.int32)), total=world * total)
        gathered = {"gather_s": round(time.perf_counter() - t0, 3)}
        if rank == 0:
            gathered.update(captures=int(len(gn)), spots=int(gn.sum()), shard_captures=total)

    # ---- dominant kernel, timed live with CUDA events on the context's stream (one extra decode, per-wave events) ----
    dec.time_kernels(True)
    job_resident(0)(dec)
    sync_ms, sync_launches, sync_cells = dec.sync_kernel_stats()
    dec.time_kernels(False)
    candidates = sync_cells / (33 * 162) if sync_cells else 0.0
    hbm_peak, peak_src, fp32_peak, fp32_src = load_peaks()
    roofline = None
    if sync_ms > 0 and candidates > 0:
        tflops = candidates * SYNC_FLOP_PER_CANDIDATE / (sync_ms * 1e-3) / 1e12
        gbs = candidates * SYNC_BYTES_PER_CANDIDATE / (sync_ms * 1e-3) / 1e9
        roofline = {"kernel": "k_sync_lags (sync_and_demodulate mode 0)", "bound": "fp32", "achieved": round(tflops, 3),
                    "peak": round(fp32_peak, 3), "unit": "TFLOP/s", "frac": round(tflops / fp32_peak, 4),
                    "traffic": int(candidates / max(sync_launches, 1) * SYNC_DRAM_BYTES_PER_CANDIDATE),
                    "traffic_source": "ncu --set full, profiles/r2_ncu_full_summary.txt, scaled to the mean candidates per launch",
                    "peak_source": fp32_src, "launches": sync_launches, "avg_launch_ms": round(sync_ms / max(sync_launches, 1), 4),
                    "flop_per_candidate": SYNC_FLOP_PER_CANDIDATE,
                    "note": "unfused multiplies and adds (exact-order sums rule FMA contraction out), issued as packed FFMA2 pairs; "
                            "no tensor-core form exists for this path",
                    "hbm": {"achieved": round(gbs, 2), "peak": hbm_peak, "unit": "GB/s", "frac": round(gbs / hbm_peak, 5),
                            "peak_source": peak_src, "bytes_per_candidate": SYNC_BYTES_PER_CANDIDATE}}
    whole_job_gbs = world * total * args.steps * BYTES_PER_CAPTURE / (total_ms * 1e-3) / 1e9

    # ---- front end kernel (the HBM-bound one), short live measurement on rank 0 ----
    frontend = None
    if rank == 0 and not args.no_frontend:
        frontend = measure_frontend(w, torch, local, hbm_peak, peak_src)

    # ---- CPU baseline + parity: the same captures through the reference's own code on the host cores ----
    cpu_line = None
    parity = None
    if cpu is not None:
        ref_results, wall = cpu.decode(0, nparity)
        parity = compare_spot_lists(ref_results, [gpu_spots(c) for c in range(nparity)])
        if rank == 0 and world == 1:
            cpu_line = {"value": round(nparity / wall, 2), "unit": UNIT, "cores": cpu.workers, "kind": cpu.kind,
                        "sample": "captures 0..%d of the same corpus, one long-lived process per core, samples in shared memory "
                                  "(FFT = oracle/fftw_standin, no FFTW on the box)" % (nparity - 1)}
        cpu.close()
    if parity is not None and world > 1:               # every rank checked a sample of its OWN shard: the line carries the sum
        keys = sorted(parity)
        t = torch.tensor([parity[k] for k in keys], dtype=torch.int64, device="cuda")
        torch.distributed.all_reduce(t)
        parity = dict(zip(keys, [int(x) for x in t.tolist()]))
    if parity is not None:
        parity["fields"] = "all 11 result fields exact (identical_spot_lists); message/call/loc/pwr + count (hard_identical)"
        parity["checked_by"] = "oracle/_ref (unmodified reference C)" if cpu.kind == "reference" else "oracle port"
        parity["per_rank"] = nparity

    if rank == 0:
        total_caps = world * total * args.steps
        line = {"metric": METRIC, "value": round(total_caps / (total_ms * 1e-3), 1), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(total_ms / args.steps, 3),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOADS[args.workload] % total, "captures_per_gpu": total},
                "run": {"l2": "inputs (%.2f GB/step/GPU) larger than L2" % (2 * total * NSAMP * 4 / 1e9),
                        "parallelism": "independent per-GPU batches, no collective on the data path",
                        "batches_in_flight": args.depth, "captures_per_call": ncap, "calls_per_step": nbatch,
                        "fano_pool": pool_stats},
                "e2e": {"value": round(total_caps / (e2e_ms * 1e-3), 1), "unit": UNIT, "h2d_bytes_per_step": 2 * total * NSAMP * 4,
                        "d2h_bytes_per_step": total * (w.MAX_UNIQUES * 80 + 4)},
                "gpu_launches": int(launches), "spots_per_step": nspots, "wall_ms_per_step": round(wall_ms / args.steps, 3),
                "clocks": clk, "roofline": roofline, "roofline_frontend": frontend,
                "whole_job_hbm": {"achieved": round(whole_job_gbs, 3), "unit": "GB/s", "frac": round(whole_job_gbs / hbm_peak, 6),
                                  "bytes_per_capture": BYTES_PER_CAPTURE},
                "cpu_baseline": cpu_line, "parity": parity, "corpus_gen_s": round(gen_s, 1),
                "schedule": dict(zip(("rounds", "deferred", "settled_f0_jitter_never"), dec.schedule_stats()))}
        if gathered:
            line["gathered"] = gathered
        print(json.dumps(line), flush=True)
    pipe.close()
    planes.close()
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def measure_frontend(w, torch, local, hbm_peak, peak_src, nstreams=8):
    stride = 2 * N_IQ + 16
    raw = torch.randint(0, 256, (nstreams * stride,), dtype=torch.uint8, device="cuda")
    fI = torch.zeros((nstreams, NSAMP), dtype=torch.float32, device="cuda")
    fQ = torch.zeros_like(fI)
    times = []
    for _ in range(4):
        _, ms = w.decimate_device(raw.data_ptr(), nstreams, N_IQ, stride, fI.data_ptr(), fQ.data_ptr(), NSAMP, NSAMP, local)
        times.append(ms)
    ms = min(times[1:])
    gbs = nstreams * BYTES_PER_STREAM / (ms * 1e-3) / 1e9
    del raw
    return {"kernel": "k_block_moments+k_comb_fir (rtlsdr_callback)", "bound": "hbm", "achieved": round(gbs, 1), "peak": hbm_peak,
            "unit": "GB/s", "frac": round(gbs / hbm_peak, 4), "traffic": int(nstreams * FRONTEND_DRAM_BYTES_PER_STREAM),
            "traffic_source": "ncu --set full, profiles/r2_ncu_full_summary.txt, scaled to the streams per launch",
            "peak_source": peak_src, "streams_per_s": round(nstreams / (ms * 1e-3), 1),
            "workload": "%d raw streams x 288e6 u8 IQ pairs resident in HBM" % nstreams}


# ---- our arm: config 4 (raw streams in, spot lists out) -------------------------------------------------------------
def run_streams(args):
    """256 raw 2.4 Msps u8 streams per GPU: rtlsdr_callback (rtlsdr_wsprd.c:126-244) -> zero tail + normalise (:285-305)
    -> wspr_decode (:316), all on the device.  The streams are synthesised ON the device from a counter-based integer
    generator (corpus.synth_raw_stream) that numpy reproduces bit for bit, so parity streams are regenerated on the host."""
    import torch
    import rtlsdr_wsprd_b200 as w
    from rtlsdr_wsprd_b200 import corpus, sharding
    rank, world, local = sharding.init_process_group()
    if not torch.cuda.is_available():
        raise w.WsprCudaError("bench.py needs a CUDA device: the decode path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    nstreams, n_iq = args.units, args.n_iq
    first = rank * nstreams
    chunk = min(args.stream_chunk, nstreams)           # streams per decimator launch
    stride = 2 * n_iq + 16                             # bytes between streams (16-byte aligned)
    nout = min(n_iq // 6401, NSAMP)
    lib = w.library()
    import ctypes as C

    def symbols(msg):
        sym = (C.c_ubyte * 162)()
        ht, lt = C.create_string_buffer(32768 * 13), C.create_string_buffer(32768 * 5)
        assert lib.get_wspr_channel_symbols(C.create_string_buffer(msg.encode(), 32), ht, lt, sym) == 1
        return np.frombuffer(bytes(sym), np.uint8).copy()

    plans = [corpus.raw_stream_plan(first + s, symbols) for s in range(nstreams)]
    t0 = time.perf_counter()
    raw = torch.empty((nstreams, stride), dtype=torch.uint8, device=dev)
    for s in range(nstreams):
        corpus.synth_raw_stream(torch, plans[s], n_iq, out=raw[s, : 2 * n_iq], device=dev)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    opts = w.default_options()
    # the streams go through `depth` contexts of `chunk` streams each: decimate -> normalise -> decode per chunk, so that one
    # chunk's decode (latency-bound at this size) overlaps the next chunk's front end
    pipe = w.PipelinedDecoder(args.depth, chunk, NSAMP, device=local)
    hs = torch.empty((nstreams * w.MAX_UNIQUES * 80,), dtype=torch.uint8).pin_memory()
    hn = torch.zeros((nstreams,), dtype=torch.int32).pin_memory()
    spots = np.frombuffer(hs.numpy().data, dtype=w.RESULT_DTYPE).reshape(nstreams, w.MAX_UNIQUES)
    keep = {}                                          # decimator output of the parity streams (before normalisation)

    def job_resident(lo, capture=False):
        n = min(chunk, nstreams - lo)

        def fn(d):
            d.decimate(raw[lo:].data_ptr(), n, n_iq, stride)
            if capture and lo == 0:
                _, _, ki, kq = d.download(samples=True)      # (synchronises: only in the untimed parity pass)
                keep["I"], keep["Q"] = ki, kq
            d.normalise()
            d.decode(opts)
            d.download(out=spots[lo:lo + n], n_out=hn.numpy()[lo:lo + n])
        return fn

    def step_resident(capture=False, steps=1):
        futs = [pipe.submit(job_resident(lo, capture)) for _ in range(steps) for lo in range(0, nstreams, chunk)]
        for f in futs:
            f.result()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    step_resident(steps=args.warmup)
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    launches0 = w.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    step_resident(steps=args.steps)
    torch.cuda.synchronize()
    ev1.record()
    barrier()
    total_ms = sharding.max_over_ranks(ev0.elapsed_time(ev1))
    launches = w.kernel_launches() - launches0
    step_resident(capture=True)
    nspots = int(hn.numpy().sum())
    res_spots, res_n = spots.copy(), hn.numpy().copy()
    # the front-end kernels alone over the same resident streams (CUDA events around every launch)
    fI = torch.zeros((chunk, NSAMP), dtype=torch.float32, device=dev)
    fQ = torch.zeros_like(fI)
    k0_ms = []
    for rep in range(2):
        k0_ms.clear()
        for lo in range(0, nstreams, chunk):
            _, ms = w.decimate_device(raw[lo:].data_ptr(), min(chunk, nstreams - lo), n_iq, stride, fI.data_ptr(), fQ.data_ptr(), NSAMP, NSAMP, local)
            k0_ms.append(ms)
    k0_total_ms = sum(k0_ms)
    del fI, fQ

    # ---- end to end: the raw bytes come from pinned host memory (PCIe-bound by construction: 576 MB per stream) ----
    nhost = min(nstreams, args.host_streams)
    hchunk = max(1, min(chunk, nhost // 2 if nhost > 1 else 1))
    hraw = torch.empty((nhost, stride), dtype=torch.uint8).pin_memory()
    hraw.copy_(raw[:nhost])
    del raw
    torch.cuda.empty_cache()
    pipe2 = w.PipelinedDecoder(2, hchunk, NSAMP, device=local)
    stages = {id(d): torch.empty((hchunk, stride), dtype=torch.uint8, device=dev) for d in pipe2.decoders}

    def job_e2e(lo):
        n = min(hchunk, nhost - lo)

        def fn(d):
            st = stages[id(d)]
            with torch.cuda.stream(torch.cuda.ExternalStream(d.stream(), device=dev)):
                st[:n].copy_(hraw[lo:lo + n], non_blocking=True)      # H2D on the context's own stream
            d.decimate(st.data_ptr(), n, n_iq, stride)
            d.normalise()
            d.decode(opts)
            d.download(out=spots[lo:lo + n], n_out=hn.numpy()[lo:lo + n])
        return fn

    def step_e2e():
        futs = [pipe2.submit(job_e2e(lo)) for lo in range(0, nhost, hchunk)]
        for f in futs:
            f.result()

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    esteps = max(1, min(args.steps, 3))
    for _ in range(esteps):
        step_e2e()
    barrier()
    e2e_s = sharding.max_over_ranks(time.perf_counter() - t0)
    clk = clocks.stop()
    spots, hn_np = res_spots, res_n
    gI, gQ = keep.get("I"), keep.get("Q")

    # ---- parity: streams regenerated on the host, the reference's callback + hand-off + decoder ----
    parity = None
    if rank == 0 and args.cpu_sample != 0:
        from oracle import pyoracle as po
        nchk = min(nstreams, chunk, nhost, 2 if args.cpu_sample < 0 else args.cpu_sample)
        ok_raw = ok_dec = ok_fe = 0
        reflib = po.ref() or po.oracle()
        for s in range(nchk):
            host = corpus.synth_raw_stream(np, plans[s], n_iq)
            ok_raw += int(np.array_equal(host, hraw[s, : 2 * n_iq].numpy())) if s < nhost else 0
            try:
                fe = po.RefFrontend()
                fe.push(host)
                ri, rq = fe.read()
            except FileNotFoundError:                  # no oracle/_ref on this box: the oracle's restatement of the callback
                orc = po.oracle()
                orc.oracle_decimate.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
                ri, rq = np.zeros(NSAMP, np.float32), np.zeros(NSAMP, np.float32)
                n = orc.oracle_decimate(host.ctypes.data, n_iq, ri.ctypes.data, rq.ctypes.data, NSAMP)
                ri, rq = ri[:n], rq[:n]
            n = min(len(ri), NSAMP)
            ok_fe += int(n == nout and np.array_equal(ri[:n], gI[s, :n]) and np.array_equal(rq[:n], gQ[s, :n]))
            fi, fq = np.zeros(NSAMP, np.float32), np.zeros(NSAMP, np.float32)
            fi[:n], fq[:n] = ri[:n], rq[:n]
            fi, fq = po.normalise_half(fi, fq)         # rtlsdr_wsprd.c:285-305
            r, _, _ = po.decode(reflib, fi, fq)
            ok_dec += int(spots_as_tuples(r) == spots_as_tuples(spots[s, : hn_np[s]]))
        parity = {"streams_checked": nchk, "raw_bytes_identical": ok_raw, "decimator_output_identical": ok_fe,
                  "identical_spot_lists": ok_dec, "checked_by": "oracle/_ref rtlsdr_callback + wspr_decode on host-regenerated streams"}

    if rank == 0:
        hbm_peak, peak_src, _, _ = load_peaks()
        units = world * nstreams * args.steps
        k0_gbs = nstreams * (2 * n_iq + 8 * nout) / (k0_total_ms * 1e-3) / 1e9 if k0_total_ms > 0 else 0.0
        e2e_rate = world * nhost * esteps / e2e_s
        line = {"metric": "raw 2.4 Msps streams decimated + decoded/sec", "value": round(units / (total_ms * 1e-3), 2), "unit": "streams/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(total_ms / args.steps, 3),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8->i32->f32", "data": "synthetic",
                "config": {"workload": WORKLOADS["config4"] % nstreams, "streams_per_gpu": nstreams, "n_iq": n_iq},
                "run": {"l2": "inputs (%.1f GB/step/GPU) larger than L2" % (nstreams * 2 * n_iq / 1e9), "streams_per_chunk": chunk,
                        "chunks_in_flight": args.depth, "front_end_ms_per_step": round(k0_total_ms, 3)},
                "e2e": {"value": round(e2e_rate, 2), "unit": "streams/s", "h2d_bytes_per_step": nhost * 2 * n_iq,
                        "d2h_bytes_per_step": nhost * (w.MAX_UNIQUES * 80 + 4), "host_streams": nhost,
                        "pcie_gbs": round(e2e_rate * 2 * n_iq / 1e9 / world, 2),
                        "note": "PCIe-bound by construction: 576 MB of raw samples per stream cross the bus"},
                "gpu_launches": int(launches), "spots_per_step": nspots, "clocks": clk,
                "roofline": {"kernel": "k_block_moments+k_comb_fir (rtlsdr_callback)", "bound": "hbm", "achieved": round(k0_gbs, 1),
                             "peak": hbm_peak, "unit": "GB/s", "frac": round(k0_gbs / hbm_peak, 4),
                             "traffic": int(chunk * FRONTEND_DRAM_BYTES_PER_STREAM * n_iq / N_IQ), "peak_source": peak_src,
                             "traffic_source": "ncu --set full, profiles/r2_ncu_full_summary.txt, scaled to the streams per launch",
                             "launches": len(k0_ms), "avg_launch_ms": round(k0_total_ms / max(len(k0_ms), 1), 3),
                             "share_of_step": round(k0_total_ms / (total_ms / args.steps), 4)},
                "cpu_baseline": None, "parity": parity, "corpus_gen_s": round(gen_s, 1)}
        print(json.dumps(line), flush=True)
    torch.cuda.synchronize()
    stages.clear()                                     # (device tensors last used on the contexts' streams: free them while those exist)
    torch.cuda.empty_cache()
    pipe.close()
    pipe2.close()
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


# ---- reference arm -------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    config = 2 if args.workload == "config2" else 3
    per_core = 8 if config == 3 else 32
    ns = max(cores, min(args.units, per_core * cores))        # bounded sample per step: >= 8 captures per core
    planes = SharedPlanes(ns)
    generate_corpus(config, 0, planes, cores)
    cpu = CpuPool(planes, cores)                              # processes, library and samples are in place before any timing
    for _ in range(args.warmup):
        cpu.decode(0, min(ns, 2 * cores))
    wall = 0.0
    for _ in range(args.steps):
        wall += cpu.decode(0, ns)[1]
    value = ns * args.steps / wall
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(wall / args.steps * 1e3, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload] % args.units, "captures_per_gpu": args.units},
            "cpu_baseline": {"value": round(value, 2), "unit": UNIT, "cores": cpu.workers, "kind": cpu.kind,
                             "sample": "%d captures of the same corpus per step (%d per core), one long-lived process per host core, "
                                       "samples in shared memory, unmodified wsprd/*.c (gcc -O3, FFT = oracle/fftw_standin)"
                                       % (ns, ns // cores)},
            "e2e": {"value": round(value, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    cpu.close()
    planes.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=6)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3", choices=sorted(WORKLOADS))
    ap.add_argument("--units", "--captures", type=int, default=0, help="captures (streams for config4) per GPU per step")
    ap.add_argument("--batch", type=int, default=4096, help="captures per decode call")
    ap.add_argument("--resident", type=int, default=4096, help="captures kept in HBM for the resident leg")
    ap.add_argument("--cpu-sample", type=int, default=-1,
                    help="captures checked against the CPU reference: -1 = all at N=1 / 256 per rank at N>1, 0 = none")
    ap.add_argument("--no-frontend", action="store_true")
    ap.add_argument("--depth", type=int, default=0, help="batches in flight per GPU (contexts driven by host threads); default 9 (config4: 12)")
    ap.add_argument("--n-iq", type=int, default=N_IQ, help="config4: raw samples per stream")
    ap.add_argument("--stream-chunk", type=int, default=16, help="config4: streams per decimator launch")
    ap.add_argument("--host-streams", type=int, default=8, help="config4: streams of the end-to-end leg (pinned host memory)")
    args = ap.parse_args()
    args.units = args.units or DEFAULT_UNITS[args.workload]
    args.depth = args.depth or (12 if args.workload == "config4" else 9)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "config4":
        run_streams(args)
        # (the interpreter's own teardown of torch's external-stream bookkeeping after the contexts' streams are gone has
        # crashed at exit; everything is flushed and released by now)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    else:
        if args.workload == "config5":
            args.resident = max(args.resident, args.units)     # every capture of the shard is distinct in both legs
            if args.steps > 3:
                args.steps, args.warmup = 2, 3        # 12 500 captures per GPU per step: keep the run to a few minutes
        run_captures(args)


if __name__ == "__main__":
    main()
