#!/bin/bash
# Round-end validation on a GPU box (run through gpurun): GPU test suite, smoke, bench (ours + reference arm), ncu launch
# list of one 4096-capture decode and a full capture of the front-end kernels.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 500 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv python tools/profile_decode.py 4096 1 > gpurun_out/ncu_launch.log 2>&1
python tools/summarise_launches.py gpurun_out/launches_final.csv > gpurun_out/launches_final.txt 2>&1; cat gpurun_out/launches_final.txt
timeout 300 ncu --set full --clock-control none -k regex:"k_block_moments|k_comb_fir" -c 2 -o gpurun_out/r1_full_frontend -f python tools/profile_frontend.py > gpurun_out/ncu_fe.log 2>&1
ncu -i gpurun_out/r1_full_frontend.ncu-rep --page raw --csv > gpurun_out/r1_full_frontend_raw.csv 2>/dev/null
