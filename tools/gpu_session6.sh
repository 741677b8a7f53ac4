#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "weak or fano or config3 or golden_corpus or drifting" > gpurun_out/pytest_rot.log 2>&1; tail -3 gpurun_out/pytest_rot.log
B="python bench.py --steps 8 --warmup 4 --cpu-sample 0 --no-frontend"
( timeout 300 $B | python tools/bench_brief.py rot4-default
  WSPR_CHAIN_WARPS=1 timeout 300 $B | python tools/bench_brief.py rot1-default
  WSPR_CARVEOUT=chain timeout 300 $B | python tools/bench_brief.py rot4-chain
  WSPR_CHAIN_WARPS=2 timeout 300 $B | python tools/bench_brief.py rot2-default
  timeout 300 $B --depth 9 | python tools/bench_brief.py rot4-default-d9
) > gpurun_out/exp6.txt 2>gpurun_out/exp6.err
cat gpurun_out/exp6.txt
