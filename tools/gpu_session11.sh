#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/exp_interference.py 4096 > gpurun_out/interference5_modes.txt 2>&1; cat gpurun_out/interference5_modes.txt
B="python bench.py --steps 8 --warmup 4 --cpu-sample 0 --no-frontend"
( timeout 300 $B --depth 9 | python tools/bench_brief.py d9
  timeout 300 $B --depth 12 | python tools/bench_brief.py d12
  WSPR_FANO_BUDGET=16384 timeout 300 $B --depth 9 | python tools/bench_brief.py d9-budget16k
) > gpurun_out/exp11.txt 2>gpurun_out/exp11.err
cat gpurun_out/exp11.txt
