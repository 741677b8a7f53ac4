"""Fano kernel alone (wspr_fano_batch on hopeless vectors: every attempt runs the full 810 000 cycles): SM clocks per Fano
cycle of one warp on its own, and the aggregate rate when many warps share the SMs (n / 32 one-warp CTAs)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import rtlsdr_wsprd_b200 as w
rng = np.random.default_rng(0)
base = rng.integers(0, 256, size=(32768, 162), dtype=np.uint8)
w.fano_batch(base[:32], maxcycles=100, solo=4)
for solo in (0, 4):   # 0: 32 attempts per warp, every field exact; 4: the instantiation the decode kernels use
    for n in (32, 4096, 148 * 32 * 4, 148 * 32 * 7):
        v = base[:n]
        t0 = time.perf_counter()
        r = w.fano_batch(v, maxcycles=10000, solo=solo)
        dt = time.perf_counter() - t0
        clk = r["clocks"].astype(np.float64)
        print("mode=%d n=%5d (%4d warps): %7.1f ms wall (timeouts %d); SM clocks per Fano cycle: median %.1f max %.1f -> %.1f ms at 1965 MHz; %.1f Mcycles/s aggregate"
              % (solo, n, n // 32, dt * 1e3, int((r["rc"] == -1).sum()), np.median(clk) / 810000, clk.max() / 810000,
                 clk.max() / 1.965e6, n * 0.81 / dt), flush=True)
