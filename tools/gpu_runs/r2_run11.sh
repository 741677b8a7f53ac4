#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
B="python bench.py --steps 8 --warmup 4 --cpu-sample 0 --no-frontend"
: > gpurun_out/r2_ab11.txt
run() {  # label, extra bench args, env...
  label=$1; extra=$2; shift; shift
  env "$@" timeout 300 $B $extra 2>>gpurun_out/r2_ab11.err | tee -a gpurun_out/r2_ab11.jsonl | python tools/bench_brief.py "$label" | tee -a gpurun_out/r2_ab11.txt
}
run "default" "" WSPR_X=1
run "nochain" "" WSPR_B200_LIB=$PWD/rtlsdr_wsprd_b200/libwsprd_b200_exp.so WSPR_DEBUG_CHAIN_MAXCYCLES=64
tail -3 gpurun_out/r2_ab11.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_c.csv python tools/profile_decode.py 4096 1 > gpurun_out/r2_ncu_launch.log 2>&1
python tools/summarise_launches.py gpurun_out/r2_launches_c.csv > gpurun_out/r2_launches_c.txt 2>&1; cat gpurun_out/r2_launches_c.txt
timeout 900 python bench.py --workload config4 --steps 3 --warmup 3 > gpurun_out/r2_bench_config4.json 2> gpurun_out/r2_bench_config4.err; cat gpurun_out/r2_bench_config4.json; tail -5 gpurun_out/r2_bench_config4.err
