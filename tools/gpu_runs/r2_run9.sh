#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
WSPR_B200_LIB=$PWD/rtlsdr_wsprd_b200/libwsprd_b200_exp.so timeout 300 python tools/exp_warp_times.py > gpurun_out/r2_warp_times.txt 2>&1; cat gpurun_out/r2_warp_times.txt
timeout 300 python tools/latency_single.py 24 > gpurun_out/r2_latency_single.txt 2>&1; cat gpurun_out/r2_latency_single.txt
B="python bench.py --steps 8 --warmup 4 --cpu-sample 0 --no-frontend"
: > gpurun_out/r2_ab9.txt
run() {  # label, extra bench args, env...
  label=$1; extra=$2; shift; shift
  env "$@" timeout 300 $B $extra 2>>gpurun_out/r2_ab9.err | tee -a gpurun_out/r2_ab9.jsonl | python tools/bench_brief.py "$label" | tee -a gpurun_out/r2_ab9.txt
}
run "default" "" WSPR_X=1
run "linger1000" "" WSPR_PARK_LINGER_US=1000
tail -3 gpurun_out/r2_ab9.err
