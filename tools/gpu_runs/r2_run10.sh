#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 8 --warmup 4 --cpu-sample 0 --no-frontend"
: > gpurun_out/r2_ab10.txt
run() {  # label, extra bench args, env...
  label=$1; extra=$2; shift; shift
  env "$@" timeout 300 $B $extra 2>>gpurun_out/r2_ab10.err | tee -a gpurun_out/r2_ab10.jsonl | python tools/bench_brief.py "$label" | tee -a gpurun_out/r2_ab10.txt
}
run "default" "" WSPR_X=1
run "cta4-persm4-pool296-carve164" "" WSPR_FANO_CTA_WARPS=4 WSPR_FANO_PER_SM=4 WSPR_FANO_POOL=296
run "cta4-persm4-pool592-carve164" "" WSPR_FANO_CTA_WARPS=4 WSPR_FANO_PER_SM=4 WSPR_FANO_POOL=592
run "cta4-persm4-pool296-carve228" "" WSPR_FANO_CTA_WARPS=4 WSPR_FANO_PER_SM=4 WSPR_FANO_POOL=296 WSPR_CARVEOUT_KB=228
run "cta2-persm2-pool296" "" WSPR_FANO_CTA_WARPS=2 WSPR_FANO_PER_SM=2 WSPR_FANO_POOL=296
run "cta4-persm4-pool148-carve164" "" WSPR_FANO_CTA_WARPS=4 WSPR_FANO_PER_SM=4 WSPR_FANO_POOL=148
tail -3 gpurun_out/r2_ab10.err
