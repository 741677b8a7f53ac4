#!/bin/bash
# Round-end validation on a GPU box (run through gpurun): GPU test suite, smoke, bench (ours + reference arm, all workloads),
# ncu launch list of one 4096-capture decode.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_final.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu_final.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_ref.err; cat gpurun_out/r2_bench_reference_arm.json
timeout 600 python bench.py --steps 20 --warmup 6 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; cat gpurun_out/r2_bench_n1.json; tail -3 gpurun_out/r2_bench_n1.err
timeout 300 python bench.py --workload config2 --steps 20 --warmup 6 --no-frontend > gpurun_out/r2_bench_config2.json 2> gpurun_out/r2_bench_config2.err; cat gpurun_out/r2_bench_config2.json | cut -c1-600
timeout 900 python bench.py --workload config4 --steps 3 --warmup 3 > gpurun_out/r2_bench_config4.json 2> gpurun_out/r2_bench_config4.err; cat gpurun_out/r2_bench_config4.json; tail -3 gpurun_out/r2_bench_config4.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_final.csv python tools/profile_decode.py 4096 1 > gpurun_out/r2_ncu_launch.log 2>&1
python tools/summarise_launches.py gpurun_out/r2_launches_final.csv > gpurun_out/r2_launches_final.txt 2>&1; cat gpurun_out/r2_launches_final.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^k_sync_lags" -s 1 -c 1 -o gpurun_out/r2_full_k_sync_lags -f python tools/profile_decode.py 1024 1 > gpurun_out/r2_ncu_k4.log 2>&1
ncu -i gpurun_out/r2_full_k_sync_lags.ncu-rep --page raw --csv > gpurun_out/r2_full_k_sync_lags_raw.csv 2>/dev/null
