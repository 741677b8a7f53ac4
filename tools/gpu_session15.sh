#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
B="python bench.py --steps 8 --warmup 4 --cpu-sample 0 --no-frontend"
( timeout 300 $B | python tools/bench_brief.py head-tables
  WSPR_B200_LIB=$PWD/gpurun_ab/lib_k4s9.so timeout 300 $B | python tools/bench_brief.py before-tables-k4s9
  timeout 300 $B | python tools/bench_brief.py head-tables-again
) > gpurun_out/exp15.txt 2>gpurun_out/exp15.err
cat gpurun_out/exp15.txt
