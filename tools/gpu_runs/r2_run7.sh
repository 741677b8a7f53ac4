#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 8 --warmup 4 --cpu-sample 0 --no-frontend"
: > gpurun_out/r2_ab7.txt
run() {  # label, extra bench args, env...
  label=$1; extra=$2; shift; shift
  env "$@" timeout 300 $B $extra 2>>gpurun_out/r2_ab7.err | tee -a gpurun_out/r2_ab7.jsonl | python tools/bench_brief.py "$label" | tee -a gpurun_out/r2_ab7.txt
}
run "default" "" WSPR_X=1
run "share-sms16-pool112-carve164" "" WSPR_FANO_SMS=16 WSPR_FANO_SHARE=1 WSPR_FANO_POOL=112
run "share-sms16-pool112-carve228" "" WSPR_FANO_SMS=16 WSPR_FANO_SHARE=1 WSPR_FANO_POOL=112 WSPR_CARVEOUT_KB=228
run "share-sms24-pool168-carve164" "" WSPR_FANO_SMS=24 WSPR_FANO_SHARE=1 WSPR_FANO_POOL=168
run "share-sms32-pool224-carve164" "" WSPR_FANO_SMS=32 WSPR_FANO_SHARE=1 WSPR_FANO_POOL=224
run "share-sms48-pool288-persm6" "" WSPR_FANO_SMS=48 WSPR_FANO_SHARE=1 WSPR_FANO_POOL=288 WSPR_FANO_PER_SM=6
run "share-sms74-pool296-persm4" "" WSPR_FANO_SMS=74 WSPR_FANO_SHARE=1 WSPR_FANO_POOL=296 WSPR_FANO_PER_SM=4 WSPR_CARVEOUT_KB=228
tail -3 gpurun_out/r2_ab7.err
bash tools/r2_ncu.sh
