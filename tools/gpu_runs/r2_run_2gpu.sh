#!/bin/bash
# two GPUs of one box: the GPU suite (two-device and sharded tests included), config 3 and config 5 at N = 2
mkdir -p gpurun_out
nvidia-smi -L; nproc
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_2gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu_2gpu.log
./tools/microbench/f32x2_bench > gpurun_out/r2_fp32_peak.txt 2>&1; tail -3 gpurun_out/r2_fp32_peak.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571"
timeout 600 $TR bench.py --gpus 2 --steps 8 --warmup 4 --no-frontend > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; cat gpurun_out/r2_bench_n2.json | cut -c1-3000; tail -3 gpurun_out/r2_bench_n2.err
timeout 900 $TR bench.py --gpus 2 --workload config5 --no-frontend > gpurun_out/r2_bench_config5_n2.json 2> gpurun_out/r2_bench_config5_n2.err; cat gpurun_out/r2_bench_config5_n2.json | cut -c1-3000; tail -3 gpurun_out/r2_bench_config5_n2.err
