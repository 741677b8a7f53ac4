#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
B="python bench.py --steps 8 --warmup 4 --cpu-sample 0 --no-frontend"
( timeout 300 $B | python tools/bench_brief.py head
  WSPR_B200_LIB=$PWD/gpurun_ab/lib_c6cabb6.so timeout 300 $B | python tools/bench_brief.py c6cabb6
  timeout 300 $B | python tools/bench_brief.py head-again
  WSPR_B200_LIB=$PWD/gpurun_ab/lib_c6cabb6.so timeout 300 $B | python tools/bench_brief.py c6cabb6-again
) > gpurun_out/exp8.txt 2>gpurun_out/exp8.err
cat gpurun_out/exp8.txt
