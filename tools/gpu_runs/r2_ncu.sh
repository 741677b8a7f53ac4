#!/bin/bash
# one ncu --set full capture of each kernel of interest during a 1024-capture decode; raw CSV pages land in gpurun_out/
mkdir -p gpurun_out
for k in k_sub_ref k_coarse k_spectrogram k_sync_freqs k_sync_lags k_sub_lpf k_fano_workers; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^$k" -s 1 -c 1 -o gpurun_out/r2_full_$k -f python tools/profile_decode.py 1024 1 > gpurun_out/r2_ncu_$k.log 2>&1
  ncu -i gpurun_out/r2_full_$k.ncu-rep --page raw --csv > gpurun_out/r2_full_${k}_raw.csv 2>/dev/null
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_block_moments|k_comb_fir" -c 2 -o gpurun_out/r2_full_frontend -f python tools/profile_frontend.py > gpurun_out/r2_ncu_fe.log 2>&1
ncu -i gpurun_out/r2_full_frontend.ncu-rep --page raw --csv > gpurun_out/r2_full_frontend_raw.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep | tail -12
