#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/exp_interference.py 4096 > gpurun_out/interference2.txt 2>&1; cat gpurun_out/interference2.txt
WSPR_LPF_WARP=1 timeout 300 python tools/exp_interference.py 4096 > gpurun_out/interference2_lpfwarp.txt 2>&1; cat gpurun_out/interference2_lpfwarp.txt
