#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 8 --warmup 4 --cpu-sample 0 --no-frontend"
: > gpurun_out/r2_ab4.txt
run() {  # label, extra bench args, env...
  label=$1; extra=$2; shift; shift
  env "$@" timeout 300 $B $extra 2>>gpurun_out/r2_ab4.err | python tools/bench_brief.py "$label" | tee -a gpurun_out/r2_ab4.txt
}
run "carve0-pool296-persm2" "" WSPR_FANO_POOL=296 WSPR_FANO_PER_SM=2
run "carve132-pool148-persm1" "" WSPR_CARVEOUT_KB=132 WSPR_FANO_POOL=148 WSPR_FANO_PER_SM=1
run "carve164-pool148-persm1" "" WSPR_CARVEOUT_KB=164 WSPR_FANO_POOL=148 WSPR_FANO_PER_SM=1
run "carve164-pool296-persm2" "" WSPR_CARVEOUT_KB=164 WSPR_FANO_POOL=296 WSPR_FANO_PER_SM=2
run "carve196-pool296-persm2" "" WSPR_CARVEOUT_KB=196 WSPR_FANO_POOL=296 WSPR_FANO_PER_SM=2
run "carve196-pool444-persm3" "" WSPR_CARVEOUT_KB=196 WSPR_FANO_POOL=444 WSPR_FANO_PER_SM=3
run "carve228-pool296-persm2" "" WSPR_CARVEOUT_KB=228 WSPR_FANO_POOL=296 WSPR_FANO_PER_SM=2
run "carve164-pool296-persm2-d12" "--depth 12" WSPR_CARVEOUT_KB=164 WSPR_FANO_POOL=296 WSPR_FANO_PER_SM=2
tail -3 gpurun_out/r2_ab4.err
