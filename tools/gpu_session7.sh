#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "streaming or weak or hashtable" > gpurun_out/pytest_s7.log 2>&1; tail -5 gpurun_out/pytest_s7.log
B="python bench.py --steps 8 --warmup 4 --cpu-sample 0 --no-frontend"
( timeout 300 $B | python tools/bench_brief.py prio
  WSPR_STREAM_PRIORITY=0 timeout 300 $B | python tools/bench_brief.py noprio
  WSPR_CARVEOUT=chain timeout 300 $B | python tools/bench_brief.py prio-chaincarve
  timeout 300 $B --depth 3 | python tools/bench_brief.py prio-d3
  timeout 300 $B --depth 9 | python tools/bench_brief.py prio-d9
) > gpurun_out/exp7.txt 2>gpurun_out/exp7.err
cat gpurun_out/exp7.txt
timeout 300 python tools/exp_interference.py 4096 > gpurun_out/interference4_prio.txt 2>&1; cat gpurun_out/interference4_prio.txt
