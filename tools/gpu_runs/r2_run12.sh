#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
B="python bench.py --steps 8 --warmup 4 --cpu-sample 0 --no-frontend"
: > gpurun_out/r2_ab12.txt
run() {  # label, extra bench args, env...
  label=$1; extra=$2; shift; shift
  env "$@" timeout 300 $B $extra 2>>gpurun_out/r2_ab12.err | tee -a gpurun_out/r2_ab12.jsonl | python tools/bench_brief.py "$label" | tee -a gpurun_out/r2_ab12.txt
}
run "default" "" WSPR_X=1
run "hybrid-sms8-ovf240" "" WSPR_FANO_SMS=8 WSPR_FANO_OVERFLOW=240
run "hybrid-sms12-ovf212" "" WSPR_FANO_SMS=12 WSPR_FANO_OVERFLOW=212
run "hybrid-sms12-ovf212-backlog2" "" WSPR_FANO_SMS=12 WSPR_FANO_OVERFLOW=212 WSPR_FANO_OVERFLOW_BACKLOG=2
run "hybrid-sms16-ovf184" "" WSPR_FANO_SMS=16 WSPR_FANO_OVERFLOW=184
run "hybrid-sms16-ovf184-backlog24" "" WSPR_FANO_SMS=16 WSPR_FANO_OVERFLOW=184 WSPR_FANO_OVERFLOW_BACKLOG=24
run "hybrid-sms20-ovf156" "" WSPR_FANO_SMS=20 WSPR_FANO_OVERFLOW=156
tail -3 gpurun_out/r2_ab12.err
