#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
WSPR_K0_BULK=1 timeout 600 python -m pytest tests -m gpu -x -q -k "frontend or streaming" > gpurun_out/r2_pytest_k0bulk.log 2>&1; tail -2 gpurun_out/r2_pytest_k0bulk.log
{ echo "# tools/profile_frontend.py 8 (8 full-length streams of uniform random bytes resident in HBM; k_block_moments + k_comb_fir, CUDA events)"; echo "# --- plain 16-byte loads (ld.global.nc.L1::no_allocate)"; timeout 120 python tools/profile_frontend.py 8; echo "# --- WSPR_K0_BULK=1: cp.async.bulk + mbarrier staging through shared memory"; WSPR_K0_BULK=1 timeout 120 python tools/profile_frontend.py 8; echo "# --- plain again"; timeout 120 python tools/profile_frontend.py 8; } > gpurun_out/r2_k0_bulk_ab.txt 2>&1; cat gpurun_out/r2_k0_bulk_ab.txt
WSPR_B200_LIB=$PWD/rtlsdr_wsprd_b200/libwsprd_b200_exp.so timeout 300 python tools/exp_warp_times.py > gpurun_out/r2_warp_times.txt 2>&1; cat gpurun_out/r2_warp_times.txt
WSPR_B200_LIB=$PWD/rtlsdr_wsprd_b200/libwsprd_b200_exp.so WSPR_FANO_PER_SM=0 timeout 300 python tools/exp_warp_times.py > gpurun_out/r2_warp_times_nocap.txt 2>&1; cat gpurun_out/r2_warp_times_nocap.txt
B="python bench.py --steps 8 --warmup 4 --cpu-sample 0 --no-frontend"
: > gpurun_out/r2_ab8.txt
run() {  # label, extra bench args, env...
  label=$1; extra=$2; shift; shift
  env "$@" timeout 300 $B $extra 2>>gpurun_out/r2_ab8.err | tee -a gpurun_out/r2_ab8.jsonl | python tools/bench_brief.py "$label" | tee -a gpurun_out/r2_ab8.txt
}
run "default" "" WSPR_X=1
run "nochain" "" WSPR_B200_LIB=$PWD/rtlsdr_wsprd_b200/libwsprd_b200_exp.so WSPR_DEBUG_CHAIN_MAXCYCLES=64
tail -3 gpurun_out/r2_ab8.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_b.csv python tools/profile_decode.py 4096 1 > gpurun_out/r2_ncu_launch.log 2>&1
python tools/summarise_launches.py gpurun_out/r2_launches_b.csv > gpurun_out/r2_launches_b.txt 2>&1; cat gpurun_out/r2_launches_b.txt
