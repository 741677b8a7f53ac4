#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
B="python bench.py --steps 8 --warmup 4 --cpu-sample 0"
: > gpurun_out/r2_ab6.txt
run() {  # label, extra bench args, env...
  label=$1; extra=$2; shift; shift
  env "$@" timeout 300 $B $extra 2>>gpurun_out/r2_ab6.err | tee -a gpurun_out/r2_ab6.jsonl | python tools/bench_brief.py "$label" | tee -a gpurun_out/r2_ab6.txt
}
EXP=$PWD/rtlsdr_wsprd_b200/libwsprd_b200_exp.so
run "default" "" WSPR_X=1
run "maxcycles2500" "--no-frontend" WSPR_B200_LIB=$EXP WSPR_DEBUG_CHAIN_MAXCYCLES=2500
run "maxcycles5000" "--no-frontend" WSPR_B200_LIB=$EXP WSPR_DEBUG_CHAIN_MAXCYCLES=5000
run "cta2-persm2" "--no-frontend" WSPR_FANO_CTA_WARPS=2
run "cta4-persm4-pool592-carve196" "--no-frontend" WSPR_FANO_CTA_WARPS=4 WSPR_FANO_PER_SM=4 WSPR_FANO_POOL=592 WSPR_CARVEOUT_KB=196
run "cta4-persm4-pool296" "--no-frontend" WSPR_FANO_CTA_WARPS=4 WSPR_FANO_PER_SM=4 WSPR_FANO_POOL=296 WSPR_CARVEOUT_KB=196
tail -3 gpurun_out/r2_ab6.err
grep -o '"roofline_frontend": {[^}]*}' gpurun_out/r2_ab6.jsonl | head -2
timeout 900 python bench.py --workload config4 --steps 3 --warmup 3 > gpurun_out/r2_bench_config4.json 2> gpurun_out/r2_bench_config4.err; cat gpurun_out/r2_bench_config4.json; tail -5 gpurun_out/r2_bench_config4.err
