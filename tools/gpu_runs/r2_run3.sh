#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2_pytest_gpu.log
WSPR_TRACE=1 WSPR_FANO_PER_SM=2 WSPR_FANO_POOL=296 timeout 120 python tools/profile_decode.py 4096 2 > gpurun_out/r2_trace.out 2> gpurun_out/r2_trace.txt; cat gpurun_out/r2_trace.out; tail -40 gpurun_out/r2_trace.txt
B="python bench.py --steps 8 --warmup 4 --cpu-sample 0 --no-frontend"
: > gpurun_out/r2_ab3.txt
run() {  # label, extra bench args, env...
  label=$1; extra=$2; shift; shift
  env "$@" timeout 300 $B $extra 2>>gpurun_out/r2_ab3.err | python tools/bench_brief.py "$label" | tee -a gpurun_out/r2_ab3.txt
}
run "pool296-persm2-d9" "" WSPR_FANO_POOL=296 WSPR_FANO_PER_SM=2
run "pool296-persm2-d12" "--depth 12" WSPR_FANO_POOL=296 WSPR_FANO_PER_SM=2
run "pool296-persm2-d14" "--depth 14" WSPR_FANO_POOL=296 WSPR_FANO_PER_SM=2
run "pool444-persm3-d12" "--depth 12" WSPR_FANO_POOL=444 WSPR_FANO_PER_SM=3
run "pool148-persm1-d12" "--depth 12" WSPR_FANO_POOL=148 WSPR_FANO_PER_SM=1
run "pool222-persm2-d12" "--depth 12" WSPR_FANO_POOL=222 WSPR_FANO_PER_SM=2
tail -3 gpurun_out/r2_ab3.err
WSPR_FANO_PER_SM=2 WSPR_FANO_POOL=296 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python tools/profile_decode.py 4096 1 > gpurun_out/r2_ncu_launch.log 2>&1
python tools/summarise_launches.py gpurun_out/r2_launches.csv > gpurun_out/r2_launches.txt 2>&1; cat gpurun_out/r2_launches.txt
