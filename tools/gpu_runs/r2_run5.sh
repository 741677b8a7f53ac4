#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 8 --warmup 4 --cpu-sample 0 --no-frontend"
: > gpurun_out/r2_ab5.txt
run() {  # label, extra bench args, env...
  label=$1; extra=$2; shift; shift
  env "$@" timeout 300 $B $extra 2>>gpurun_out/r2_ab5.err | python tools/bench_brief.py "$label" | tee -a gpurun_out/r2_ab5.txt
}
EXP=$PWD/rtlsdr_wsprd_b200/libwsprd_b200_exp.so
run "default(carve164-pool296-persm2)" ""  WSPR_X=1
run "nochain-carve164" "" WSPR_B200_LIB=$EXP WSPR_DEBUG_CHAIN_MAXCYCLES=64
run "nochain-carve0" "" WSPR_B200_LIB=$EXP WSPR_DEBUG_CHAIN_MAXCYCLES=64 WSPR_CARVEOUT_KB=0
run "nochain-carve164-d3" "--depth 3" WSPR_B200_LIB=$EXP WSPR_DEBUG_CHAIN_MAXCYCLES=64
run "default-d6" "--depth 6" WSPR_X=1
tail -3 gpurun_out/r2_ab5.err
timeout 600 python bench.py > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err; cat gpurun_out/r2_bench_full.json; tail -3 gpurun_out/r2_bench_full.err
timeout 300 python bench.py --workload config2 --steps 8 --warmup 4 --no-frontend > gpurun_out/r2_bench_config2.json 2> gpurun_out/r2_bench_config2.err; cat gpurun_out/r2_bench_config2.json; tail -3 gpurun_out/r2_bench_config2.err
timeout 600 python bench.py --workload config4 --units 16 --steps 3 --warmup 3 --host-streams 4 > gpurun_out/r2_bench_config4_16.json 2> gpurun_out/r2_bench_config4_16.err; cat gpurun_out/r2_bench_config4_16.json; tail -5 gpurun_out/r2_bench_config4_16.err
