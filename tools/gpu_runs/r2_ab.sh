#!/bin/bash
# A/B of the Fano worker pool on one box: bench.py under several WSPR_FANO_POOL / WSPR_FANO_PER_SM / WSPR_PARK_LINGER_US.
mkdir -p gpurun_out
B="python bench.py --steps 8 --warmup 4 --cpu-sample 0 --no-frontend"
: > gpurun_out/r2_ab2.txt
run() {  # label, extra bench args, env...
  label=$1; extra=$2; shift; shift
  env "$@" timeout 300 $B $extra 2>>gpurun_out/r2_ab2.err | python tools/bench_brief.py "$label" | tee -a gpurun_out/r2_ab2.txt
}
run "pool296" "" WSPR_FANO_POOL=296
run "pool444" "" WSPR_FANO_POOL=444
run "pool592" "" WSPR_FANO_POOL=592
run "pool888" "" WSPR_FANO_POOL=888
run "pool296-persm2" "" WSPR_FANO_POOL=296 WSPR_FANO_PER_SM=2
run "pool592-persm4" "" WSPR_FANO_POOL=592 WSPR_FANO_PER_SM=4
run "pool296-linger0" "" WSPR_FANO_POOL=296 WSPR_PARK_LINGER_US=0
run "pool296-linger2000" "" WSPR_FANO_POOL=296 WSPR_PARK_LINGER_US=2000
run "pool296-depth6" "--depth 6" WSPR_FANO_POOL=296
