#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "weak or config3 or golden_corpus or drifting or subtract" > gpurun_out/pytest_s14.log 2>&1; tail -3 gpurun_out/pytest_s14.log
B="python bench.py --steps 8 --warmup 4 --cpu-sample 0 --no-frontend"
( timeout 300 $B | python tools/bench_brief.py head-s18
  WSPR_B200_LIB=$PWD/gpurun_ab/lib_k4s9.so timeout 300 $B | python tools/bench_brief.py k4s9
  WSPR_B200_LIB=$PWD/gpurun_ab/lib_k4s6.so timeout 300 $B | python tools/bench_brief.py k4s6
  WSPR_B200_LIB=$PWD/gpurun_ab/lib_k4s9_lpf128.so timeout 300 $B | python tools/bench_brief.py k4s9-lpf128
  WSPR_B200_LIB=$PWD/gpurun_ab/lib_k4s9.so WSPR_DEBUG_CHAIN_MAXCYCLES=64 timeout 300 $B | python tools/bench_brief.py k4s9-nochain
  WSPR_DEBUG_CHAIN_MAXCYCLES=64 timeout 300 $B | python tools/bench_brief.py head-s18-nochain
) > gpurun_out/exp14.txt 2>gpurun_out/exp14.err
cat gpurun_out/exp14.txt
