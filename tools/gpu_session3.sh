#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hashtable or fixture" > gpurun_out/pytest_ht.log 2>&1; tail -5 gpurun_out/pytest_ht.log
timeout 300 python tools/exp_interference.py 4096 > gpurun_out/interference.txt 2>&1; cat gpurun_out/interference.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sync_lags|k_sub_lpf" -s 6 -c 2 -o gpurun_out/r1_full_packed -f python tools/profile_decode.py 1024 1 > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/r1_full_packed.ncu-rep --page raw --csv > gpurun_out/r1_full_packed_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -5
