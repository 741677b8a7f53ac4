#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "weak or config3 or golden_corpus or drifting or option" > gpurun_out/pytest_s12.log 2>&1; tail -3 gpurun_out/pytest_s12.log
B="python bench.py --steps 8 --warmup 4 --cpu-sample 0 --no-frontend --depth 9"
( timeout 300 $B | python tools/bench_brief.py pieces2
  WSPR_CHAIN_PIECES=1 timeout 300 $B | python tools/bench_brief.py pieces1
  WSPR_CARVEOUT=chain timeout 300 $B | python tools/bench_brief.py pieces2-chaincarve
  timeout 300 $B --depth 6 | python tools/bench_brief.py pieces2-d6
) > gpurun_out/exp12.txt 2>gpurun_out/exp12.err
cat gpurun_out/exp12.txt
