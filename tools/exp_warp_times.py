"""Experiment build only (make -C rtlsdr_wsprd_b200/csrc exp; WSPR_B200_LIB=.../libwsprd_b200_exp.so): average run time of
the warps of K4 (k_sync_lags) and K6 (k_sub_lpf) by the number of Fano worker warps resident on their SM, measured inside
a pipelined decode of the config-3 corpus (device clocks)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import rtlsdr_wsprd_b200 as w
import helpers as H
n, depth, steps = 4096, 9, 18
I, Q, _ = H.make_corpus(3, 256)
I = np.tile(I, (n // 256, 1)); Q = np.tile(Q, (n // 256, 1))
lib = w.library()
hist = (C.c_ulonglong * 16)()
with w.PipelinedDecoder(depth, n) as pipe:
    def job(d):
        d.upload(I, Q); d.decode(); return d.download()
    [f.result() for f in [pipe.submit(job) for _ in range(depth)]]
    lib.wspr_debug_hist(hist, 1)
    [f.result() for f in [pipe.submit(job) for _ in range(steps)]]
    lib.wspr_debug_hist(hist, 0)
    print(w.fano_pool_stats())
h = np.array(list(hist), dtype=np.float64).reshape(2, 4, 2)
for k, name in enumerate(("k_sync_lags", "k_sub_lpf")):
    base = h[k, 0, 0] / max(h[k, 0, 1], 1)
    for cls, what in enumerate(("no worker on the SM", "workers on the SM, none on this scheduler", "one worker on this scheduler", "two or more on this scheduler")):
        cnt = h[k, cls, 1]
        avg = h[k, cls, 0] / max(cnt, 1)
        print("%-12s %-42s: %10d warps, mean %9.0f clocks (x%.3f)" % (name, what, cnt, avg, avg / base if base else 0))
