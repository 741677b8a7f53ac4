#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
B="python bench.py --steps 8 --warmup 4 --cpu-sample 0 --no-frontend"
( timeout 300 $B | python tools/bench_brief.py head
  WSPR_B200_LIB=$PWD/gpurun_ab/lib_c6cabb6.so timeout 300 $B | python tools/bench_brief.py c6cabb6
  timeout 300 $B --depth 9 | python tools/bench_brief.py head-d9
  WSPR_FANO_BUDGET=1024 timeout 300 $B | python tools/bench_brief.py head-budget1024
  WSPR_FANO_BUDGET=16384 timeout 300 $B | python tools/bench_brief.py head-budget16384
) > gpurun_out/exp9.txt 2>gpurun_out/exp9.err
cat gpurun_out/exp9.txt
