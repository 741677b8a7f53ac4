#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
B="python bench.py --steps 8 --warmup 4 --cpu-sample 0 --no-frontend"
: > gpurun_out/r2_ab13.txt
run() {  # label, extra bench args, env...
  label=$1; extra=$2; shift; shift
  env "$@" timeout 300 $B $extra 2>>gpurun_out/r2_ab13.err | tee -a gpurun_out/r2_ab13.jsonl | python tools/bench_brief.py "$label" | tee -a gpurun_out/r2_ab13.txt
}
run "default" "" WSPR_X=1
run "default-d7" "--depth 7" WSPR_X=1
run "default-d12" "--depth 12" WSPR_X=1
run "nochain" "" WSPR_B200_LIB=$PWD/rtlsdr_wsprd_b200/libwsprd_b200_exp.so WSPR_DEBUG_CHAIN_MAXCYCLES=64
tail -3 gpurun_out/r2_ab13.err
timeout 300 python tools/latency_single.py 24 > gpurun_out/r2_latency_single.txt 2>&1; cat gpurun_out/r2_latency_single.txt
