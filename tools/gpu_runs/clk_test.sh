nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,pstate --format=csv,noheader -lms 100 > gpurun_out/clk.csv &
NPID=$!
python - <<'PY'
import sys, time
sys.path.insert(0,'.')
import numpy as np
import rtlsdr_wsprd_b200 as w
rng=np.random.default_rng(0)
base=rng.integers(0,256,size=(4096,162),dtype=np.uint8)
w.fano_batch(base[:1], maxcycles=10, solo=1)
for n in (1,1,1,148,148):
    t0=time.perf_counter(); w.fano_batch(base[:n], maxcycles=10000, solo=1); print(n, round((time.perf_counter()-t0)*1e3,1),"ms", flush=True)
PY
kill $NPID
sort gpurun_out/clk.csv | uniq -c | sort -rn | head
