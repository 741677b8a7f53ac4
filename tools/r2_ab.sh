#!/bin/bash
# A/B of the Fano worker pool placement on one box: GPU suite first, then bench.py under several WSPR_FANO_SMS / WSPR_FANO_POOL.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2_pytest_gpu.log
timeout 300 python tools/fano_microbench.py > gpurun_out/r2_fano_microbench.txt 2>&1; cat gpurun_out/r2_fano_microbench.txt
B="python bench.py --steps 8 --warmup 4 --cpu-sample 0 --no-frontend"
: > gpurun_out/r2_ab.txt
for cfg in "0 148" "0 296" "0 74" "16 112" "12 84" "8 56"; do
  set -- $cfg
  WSPR_FANO_SMS=$1 WSPR_FANO_POOL=$2 timeout 300 $B 2>>gpurun_out/r2_ab.err | python tools/bench_brief.py "sms=$1 pool=$2" | tee -a gpurun_out/r2_ab.txt
done
