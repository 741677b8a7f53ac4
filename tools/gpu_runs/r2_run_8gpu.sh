#!/bin/bash
# eight GPUs of one box: config 3 (the metric's configuration, 4096 captures per GPU) and config 5 (100 000 distinct captures in
# eight contiguous shards, results gathered on rank 0), each with per-rank parity against the compiled reference
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; free -g | head -2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581"
timeout 600 $TR bench.py --gpus 8 --steps 8 --warmup 4 --no-frontend --cpu-sample 64 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; cat gpurun_out/r2_bench_n8.json | cut -c1-3500; tail -3 gpurun_out/r2_bench_n8.err
timeout 900 $TR bench.py --gpus 8 --workload config5 --no-frontend --cpu-sample 64 > gpurun_out/r2_bench_config5_n8.json 2> gpurun_out/r2_bench_config5_n8.err; cat gpurun_out/r2_bench_config5_n8.json | cut -c1-3500; tail -3 gpurun_out/r2_bench_config5_n8.err
