#!/bin/bash
# CTA-shape coupling experiment (tools/microbench/coupling.cu) and the A/B of K4 / K6 with warps claiming their tasks dynamically
mkdir -p gpurun_out
./tools/microbench/coupling > gpurun_out/r2_coupling_microbench.txt 2>&1; cat gpurun_out/r2_coupling_microbench.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
B="python bench.py --steps 8 --warmup 4 --cpu-sample 0 --no-frontend"
: > gpurun_out/r2_ab14.txt
run() {  # label, extra bench args, env...
  label=$1; extra=$2; shift; shift
  env "$@" timeout 300 $B $extra 2>>gpurun_out/r2_ab14.err | tee -a gpurun_out/r2_ab14.jsonl | python tools/bench_brief.py "$label" | tee -a gpurun_out/r2_ab14.txt
}
L=$PWD/rtlsdr_wsprd_b200
run "default" "" WSPR_X=1
run "exp(static)" "" WSPR_B200_LIB=$L/libwsprd_b200_exp.so
run "k4w8" "" WSPR_B200_LIB=$L/libwsprd_b200_k4w8.so
run "lpf32" "" WSPR_B200_LIB=$L/libwsprd_b200_lpf32.so
run "dyn(k4w8+lpf32)" "" WSPR_B200_LIB=$L/libwsprd_b200_dyn.so
run "dyn12(k4w12+lpf32)" "" WSPR_B200_LIB=$L/libwsprd_b200_dyn12.so
run "dyn-nochain" "" WSPR_B200_LIB=$L/libwsprd_b200_dyn.so WSPR_DEBUG_CHAIN_MAXCYCLES=64
B="python bench.py --steps 4 --warmup 3 --cpu-sample 512 --no-frontend"
run "dyn-parity512" "" WSPR_B200_LIB=$L/libwsprd_b200_dyn.so
tail -3 gpurun_out/r2_ab14.err
