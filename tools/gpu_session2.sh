#!/bin/bash
# ad-hoc measurement session (results under gpurun_out/)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
B="python bench.py --steps 8 --warmup 4 --cpu-sample 0 --no-frontend"
( timeout 300 $B | python tools/bench_brief.py packed
  WSPR_DEBUG_CHAIN_MAXCYCLES=64 timeout 300 $B | python tools/bench_brief.py nochain
  timeout 300 $B --depth 3 | python tools/bench_brief.py d3
  timeout 300 $B --depth 9 | python tools/bench_brief.py d9
) > gpurun_out/exp2.txt 2>gpurun_out/exp2.err
cat gpurun_out/exp2.txt
WSPR_TRACE=1 timeout 200 python tools/profile_decode.py 4096 2 > gpurun_out/trace_4096.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_packed.csv python tools/profile_decode.py 1024 1 > gpurun_out/ncu_launch.log 2>&1
python tools/summarise_launches.py gpurun_out/launches_packed.csv > gpurun_out/launches_packed.txt 2>&1; cat gpurun_out/launches_packed.txt
