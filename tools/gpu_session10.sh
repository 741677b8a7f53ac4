#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 8 --warmup 4 --cpu-sample 0 --no-frontend"
( timeout 300 $B | python tools/bench_brief.py default
  WSPR_CARVEOUT=86 timeout 300 $B | python tools/bench_brief.py all-86pct
  WSPR_CARVEOUT=72 timeout 300 $B | python tools/bench_brief.py all-72pct
  WSPR_CARVEOUT=58 timeout 300 $B | python tools/bench_brief.py all-58pct
  WSPR_CARVEOUT=max timeout 300 $B | python tools/bench_brief.py all-max
  WSPR_CARVEOUT=86 WSPR_DEBUG_CHAIN_MAXCYCLES=64 timeout 300 $B | python tools/bench_brief.py all-86pct-nochain
  WSPR_DEBUG_CHAIN_MAXCYCLES=64 timeout 300 $B | python tools/bench_brief.py default-nochain
) > gpurun_out/exp10.txt 2>gpurun_out/exp10.err
cat gpurun_out/exp10.txt
