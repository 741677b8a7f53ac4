#!/usr/bin/env python3
"""bench.py -- throughput of the WSPR decode hot path on B200 (metric and config of BASELINE.json).

  python bench.py --gpus N --steps K --warmup W            our CUDA path (one process per GPU under torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K ...   the reference's own CPU code (oracle/_ref) on the host cores

A step = one pass of the decode path (both passes, subtraction on: reference defaults rtlsdr_wsprd.c:357-362) over one
batch of BASELINE config 3 per GPU: 4 096 synthetic captures x 10 overlapping signals, SNR -28..-10 dB.  Prints ONE
JSON line on rank 0.  `value` = captures/s with the captures resident in HBM (device-timed, max over ranks); `e2e` =
the same through the C ABI from pinned host buffers (H2D of the captures and D2H of the spot records inside the timed
region).  oracle/ is used here only by the cpu_baseline / --impl reference legs and for the parity count.
"""
import argparse
import concurrent.futures as cf
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

# several contexts x (1 main + 12 side) streams: give them enough hardware queues (must be set before CUDA starts)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "2-min WSPR captures decoded/sec"
UNIT = "captures/s"
CAPTURES_PER_GPU = 4096
WORKLOAD = "config3: %d captures/GPU x 10 overlapping signals, SNR -28..-10 dB, 45000 samples @375 sps, 2 passes + subtraction"
NSAMP = 45000
# algorithmic bytes per unit (SURVEY.md section 8d / DESIGN.md)
BYTES_PER_CAPTURE = 360000 + 80 * 10 + 4
SYNC_BYTES_PER_CANDIDATE = (162 * 256 + 256) * 8 + 33 * 162 * 16      # IQ window read + per-(lag,symbol) tone powers written
SYNC_FLOP_PER_CANDIDATE = 33 * 162 * 256 * 32                         # 4 tones x (4 mul + 4 add) per sample, unfused
# dram__bytes_read.sum + dram__bytes_write.sum of k_sync_lags per candidate, from the ncu --set full capture summarised in
# profiles/r1_ncu_full_packed.txt (1024 candidates per launch: 343.33 MB read + 70.86 MB written)
SYNC_DRAM_BYTES_PER_CANDIDATE = (343.327e6 + 70.856e6) / 1024
# the same for the front end (k_block_moments + k_comb_fir) per raw stream, profiles/r1_ncu_full_frontend.txt (4 streams per
# launch: 2303.97 MB + 2.90 MB read, 5.64 MB written): the 576 MB of a stream are read exactly once
FRONTEND_DRAM_BYTES_PER_STREAM = (2303.973e6 + 2.895e6 + 5.645e6) / 4


# ---- corpus (host, seeded; identical arrays go to the GPU path and to the CPU reference) --------------------------
def _gen_chunk(args):
    config, lo, hi = args
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from rtlsdr_wsprd_b200 import corpus
    from oracle import pyoracle as po           # channel symbols for the generator (test infrastructure, not timed)
    import ctypes as C
    orc = po.oracle()

    def symbols(msg):
        sym = (C.c_ubyte * 162)()
        ht, lt = C.create_string_buffer(32768 * 13), C.create_string_buffer(32768 * 5)
        orc.get_wspr_channel_symbols(C.create_string_buffer(msg.encode(), 32), ht, lt, sym)
        return np.frombuffer(bytes(sym), np.uint8).copy()
    I, Q, _ = corpus.make_corpus(config, hi - lo, symbols, start=lo)
    return lo, I, Q


def make_corpus_parallel(config, lo, hi, out_i, out_q, workers):
    from oracle import pyoracle as po
    po.oracle()                                    # build liboracle.so once, before forking
    step = max(1, min(64, (hi - lo + workers - 1) // workers))
    jobs = [(config, a, min(a + step, hi)) for a in range(lo, hi, step)]
    with cf.ProcessPoolExecutor(max_workers=workers) as ex:
        for a, I, Q in ex.map(_gen_chunk, jobs):
            out_i[a - lo:a - lo + len(I)] = I
            out_q[a - lo:a - lo + len(Q)] = Q


# ---- CPU reference leg ---------------------------------------------------------------------------------------------
def _cpu_decode_shard(args):
    kind, I, Q = args
    from oracle import pyoracle as po
    lib = po.ref() if kind == "reference" else po.oracle()
    old = os.getcwd()
    scratch = tempfile.mkdtemp(prefix="wspr_cpu_")      # the reference writes fftw_wisdom.dat into the CWD (wsprd.c:835)
    os.chdir(scratch)
    out = []
    t0 = time.perf_counter()
    for c in range(len(I)):
        r, _, _ = po.decode(lib, I[c], Q[c], cwd_scratch=False)
        out.append([(x["message"], x["call"], x["loc"], x["pwr"], float(x["freq"]), float(x["snr"]), float(x["dt"])) for x in r])
    dt = time.perf_counter() - t0
    os.chdir(old)
    return out, dt


def cpu_reference_kind():
    from oracle import pyoracle as po
    return "reference" if po.ref() is not None else "port"


def run_cpu(I, Q, workers):
    """Decode the captures with the reference's CPU code, one process per core, contiguous shards.  Returns
    (captures/s aggregate wall, results, kind)."""
    kind = cpu_reference_kind()
    n = len(I)
    workers = max(1, min(workers, n))
    per = (n + workers - 1) // workers
    jobs = [(kind, I[a:a + per], Q[a:a + per]) for a in range(0, n, per)]
    t0 = time.perf_counter()
    with cf.ProcessPoolExecutor(max_workers=workers) as ex:
        parts = list(ex.map(_cpu_decode_shard, jobs))
    wall = time.perf_counter() - t0
    results = [r for p, _ in parts for r in p]
    return n / wall, results, kind, len(jobs)


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


# ---- our arm -------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import rtlsdr_wsprd_b200 as w
    from rtlsdr_wsprd_b200 import sharding
    rank, world, local = sharding.init_process_group()
    if not torch.cuda.is_available():
        raise w.WsprCudaError("bench.py needs a CUDA device: the decode path has no CPU fallback")
    torch.cuda.set_device(local)
    ncap = args.captures
    lo = rank * ncap                                   # weak scaling: every GPU gets its own contiguous shard of the corpus
    host_workers = max(1, (os.cpu_count() or 8) // world)
    # every context is driven by a host thread that polls (yielding) for the two counter reads of a round; with fewer host
    # cores per rank than contexts the threads sleep in the driver instead (read once, when the library loads)
    if host_workers < args.depth:
        os.environ.setdefault("WSPR_WAIT", "block")

    # pinned host planes (the e2e leg copies from these every step)
    hI = torch.empty((ncap, NSAMP), dtype=torch.float32).pin_memory()
    hQ = torch.empty((ncap, NSAMP), dtype=torch.float32).pin_memory()
    t0 = time.perf_counter()
    make_corpus_parallel(3, lo, lo + ncap, hI.numpy(), hQ.numpy(), host_workers)
    gen_s = time.perf_counter() - t0
    dI, dQ = hI.cuda(non_blocking=True), hQ.cuda(non_blocking=True)      # pristine device copy (decode subtracts in place)
    # `depth` batches in flight (one context + host thread each): the next batch's bulk overlaps the previous one's tail
    pipe = w.PipelinedDecoder(args.depth, ncap, NSAMP, device=local)
    dec = pipe.decoders[0]
    opts = w.default_options()
    outs = {}
    for d in pipe.decoders:
        hs = torch.empty((ncap * w.MAX_UNIQUES * 80,), dtype=torch.uint8).pin_memory()
        hn = torch.empty((ncap,), dtype=torch.int32).pin_memory()
        outs[id(d)] = (hs, hn, np.frombuffer(hs.numpy().data, dtype=w.RESULT_DTYPE).reshape(ncap, w.MAX_UNIQUES))
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def step_resident(d):
        d.upload_device(dI.data_ptr(), dQ.data_ptr(), ncap, NSAMP)
        d.decode(opts)

    def step_e2e(d):
        hs, hn, sp = outs[id(d)]
        d.upload_ptr(hI.data_ptr(), hQ.data_ptr(), ncap)
        d.decode(opts)
        d.download(out=sp, n_out=hn.numpy())
        return sp, hn

    def run_steps(fn, k):
        futs = [pipe.submit(fn) for _ in range(k)]
        return [f.result() for f in futs]

    # ---- resident-input throughput (`value`) ----
    run_steps(step_resident, args.warmup)
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    launches0 = w.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()                                   # device idle here (barrier above): default-stream events bracket all streams
    run_steps(step_resident, args.steps)
    torch.cuda.synchronize()
    ev1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    dev_ms = ev0.elapsed_time(ev1)
    launches = w.kernel_launches() - launches0
    total_ms = sharding.max_over_ranks(max(dev_ms, 0.0))
    wall_ms = sharding.max_over_ranks(wall_ms)

    # ---- end to end through the C ABI with host buffers ----
    run_steps(step_e2e, min(args.depth, args.steps))
    barrier()
    t0 = time.perf_counter()
    res = run_steps(step_e2e, args.steps)
    barrier()
    e2e_ms = sharding.max_over_ranks((time.perf_counter() - t0) * 1e3)
    clk = clocks.stop()
    spots_np, h_n = res[-1]
    nspots = int(h_n.numpy().sum())
    gpu_results = [[(x["message"], x["call"], x["loc"], x["pwr"], float(x["freq"]), float(x["snr"]), float(x["dt"]))
                    for x in spots_np[c, : h_n[c]]] for c in range(min(ncap, args.cpu_sample))]

    # ---- dominant kernel, timed live with CUDA events on the context's stream (one extra decode, per-wave events) ----
    dec.time_kernels(True)
    step_resident(dec)
    sync_ms, sync_launches, sync_cells = dec.sync_kernel_stats()
    dec.time_kernels(False)
    candidates = sync_cells / (33 * 162) if sync_cells else 0.0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    roofline = None
    if sync_ms > 0 and candidates > 0:
        gbs = candidates * SYNC_BYTES_PER_CANDIDATE / (sync_ms * 1e-3) / 1e9
        roofline = {"kernel": "k_sync_lags (sync_and_demodulate mode 0)", "bound": "hbm", "achieved": round(gbs, 2), "peak": hbm_peak,
                    "unit": "GB/s", "frac": round(gbs / hbm_peak, 5),
                    "traffic": int(candidates / max(sync_launches, 1) * SYNC_DRAM_BYTES_PER_CANDIDATE),
                    "traffic_source": "ncu --set full, profiles/r1_ncu_full_packed.txt, scaled to the mean candidates per launch",
                    "peak_source": peak_src,
                    "launches": sync_launches, "avg_launch_ms": round(sync_ms / max(sync_launches, 1), 4),
                    "note": "FP32-pipe bound, not HBM bound: %.2f TFLOP/s unfused fp32 as packed FFMA2 pairs (%.3g flop per candidate; ncu: 82 %% FMA-pipe active)"
                            % (candidates * SYNC_FLOP_PER_CANDIDATE / (sync_ms * 1e-3) / 1e12, SYNC_FLOP_PER_CANDIDATE)}
    whole_job_gbs = world * ncap * args.steps * BYTES_PER_CAPTURE / (total_ms * 1e-3) / 1e9

    # ---- front end kernel (the HBM-bound one), short live measurement on rank 0 ----
    frontend = None
    if rank == 0 and not args.no_frontend:
        nstreams, n_iq = 8, 288_000_000
        stride = 2 * n_iq + 16
        raw = torch.randint(0, 256, (nstreams * stride,), dtype=torch.uint8, device="cuda")
        fI = torch.zeros((nstreams, NSAMP), dtype=torch.float32, device="cuda")
        fQ = torch.zeros_like(fI)
        times = []
        for _ in range(4):
            _, ms = w.decimate_device(raw.data_ptr(), nstreams, n_iq, stride, fI.data_ptr(), fQ.data_ptr(), NSAMP, NSAMP, local)
            times.append(ms)
        ms = min(times[1:])
        gbs = nstreams * (2 * n_iq + 2 * 4 * 44992) / (ms * 1e-3) / 1e9
        frontend = {"kernel": "k_block_moments+k_comb_fir (rtlsdr_callback)", "bound": "hbm", "achieved": round(gbs, 1), "peak": hbm_peak,
                    "unit": "GB/s", "frac": round(gbs / hbm_peak, 4), "traffic": int(nstreams * FRONTEND_DRAM_BYTES_PER_STREAM),
                    "traffic_source": "ncu --set full, profiles/r1_ncu_full_frontend.txt, scaled to the streams per launch",
                    "streams_per_s": round(nstreams / (ms * 1e-3), 1),
                    "workload": "%d raw streams x 288e6 u8 IQ pairs resident in HBM" % nstreams}
        del raw

    # ---- CPU baseline on a bounded sample of the same workload (rank 0, N=1 only) + parity count ----
    cpu = None
    parity = None
    if rank == 0 and world == 1 and args.cpu_sample > 0:
        ns = min(ncap, args.cpu_sample)
        cores = os.cpu_count() or 1
        rate, cpu_results, kind, used = run_cpu(hI.numpy()[:ns], hQ.numpy()[:ns], cores)
        cpu = {"value": round(rate, 2), "unit": UNIT, "cores": used, "kind": kind,
               "sample": "first %d captures of the same corpus, one process per core (FFT = oracle/fftw_standin, no FFTW on the box)" % ns}
        same = sum(1 for a, b in zip(cpu_results, gpu_results) if a == b)
        parity = {"captures_checked": ns, "identical_spot_lists": same,
                  "fields": "message, call, loc, pwr, freq, snr, dt (exact)"}

    if rank == 0:
        total_caps = world * ncap * args.steps
        line = {"metric": METRIC, "value": round(total_caps / (total_ms * 1e-3), 1), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(total_ms / args.steps, 3),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD % ncap, "captures_per_gpu": ncap, "l2": "inputs (1.47 GB/step/GPU) larger than L2",
                           "parallelism": "independent per-GPU batches, no collective", "batches_in_flight": args.depth},
                "e2e": {"value": round(total_caps / (e2e_ms * 1e-3), 1), "unit": UNIT, "h2d_bytes_per_step": 2 * ncap * NSAMP * 4,
                        "d2h_bytes_per_step": ncap * (w.MAX_UNIQUES * 80 + 4)},
                "gpu_launches": int(launches), "spots_per_step": nspots, "wall_ms_per_step": round(wall_ms / args.steps, 3),
                "clocks": clk, "roofline": roofline, "roofline_frontend": frontend,
                "whole_job_hbm": {"achieved": round(whole_job_gbs, 3), "unit": "GB/s", "frac": round(whole_job_gbs / hbm_peak, 6),
                                  "bytes_per_capture": BYTES_PER_CAPTURE},
                "cpu_baseline": cpu, "parity": parity, "corpus_gen_s": round(gen_s, 1),
                "schedule": dict(zip(("rounds", "deferred", "settled_f0_jitter_never"), dec.schedule_stats()))}
        print(json.dumps(line), flush=True)
    pipe.close()
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
    ns = max(cores, min(args.cpu_sample, 4 * cores))          # bounded sample per step: a few captures per core
    I = np.zeros((ns, NSAMP), np.float32)
    Q = np.zeros((ns, NSAMP), np.float32)
    make_corpus_parallel(3, 0, ns, I, Q, cores)
    for _ in range(args.warmup):
        run_cpu(I[:cores], Q[:cores], cores)
    t0 = time.perf_counter()
    used = kind = None
    for _ in range(args.steps):
        _, _, kind, used = run_cpu(I, Q, cores)
    wall = time.perf_counter() - t0
    value = ns * args.steps / wall
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(wall / args.steps * 1e3, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD % args.captures, "captures_per_gpu": args.captures},
            "cpu_baseline": {"value": round(value, 2), "unit": UNIT, "cores": used, "kind": kind,
                             "sample": "%d captures of the same corpus per step, one process per host core, unmodified wsprd/*.c "
                                       "(gcc -O3, FFT = oracle/fftw_standin)" % ns},
            "e2e": {"value": round(value, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=6)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--captures", type=int, default=CAPTURES_PER_GPU, help="captures per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=96, help="captures decoded by the CPU baseline / parity leg")
    ap.add_argument("--no-frontend", action="store_true")
    ap.add_argument("--depth", type=int, default=9, help="batches in flight per GPU (contexts driven by host threads)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
