#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/exp_interference.py 4096 > gpurun_out/interference3_carveout.txt 2>&1; cat gpurun_out/interference3_carveout.txt
B="python bench.py --steps 8 --warmup 4 --cpu-sample 0 --no-frontend"
( timeout 300 $B | python tools/bench_brief.py carveout-max
  WSPR_CARVEOUT=default timeout 300 $B | python tools/bench_brief.py carveout-default
  WSPR_LPF_WARP=1 timeout 300 $B | python tools/bench_brief.py carveout-max-lpfwarp
  timeout 300 $B --depth 3 | python tools/bench_brief.py carveout-max-d3
) > gpurun_out/exp5.txt 2>gpurun_out/exp5.err
cat gpurun_out/exp5.txt
