"""Experiment: what does a resident one-warp CTA cost the bulk decode kernels that share its SM?  Decodes a batch with the
chain cut short (WSPR_DEBUG_CHAIN_MAXCYCLES, wrong results, timing only) while K one-warp CTAs are in flight on another
stream: Fano on hopeless attempts with the tree in shared memory (the shape of k_chain_fano) or in global memory, CTAs that
only hold shared memory and sleep, CTAs that sleep, CTAs that run a dependent integer loop.  usage: exp_interference.py [ncap]"""
import os, sys, time, ctypes as C
os.environ["WSPR_DEBUG_CHAIN_MAXCYCLES"] = "64"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import rtlsdr_wsprd_b200 as w
import helpers as H
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
I, Q, _ = H.make_corpus(3, 64)
I = np.tile(I, (n // 64, 1)); Q = np.tile(Q, (n // 64, 1))
lib = w.library()
lib.wspr_debug_fano_load.argtypes = [C.c_int, C.c_uint, C.c_int]
MODES = {0: "fano, tree in shared memory", 2: "fano, tree in global memory", 8: "sleep holding 84 KB shared", 9: "sleep",
         10: "integer loop"}
with w.BatchDecoder(n) as d:
    d.upload(I, Q)
    d.decode()
    for mode in (0, 2, 8, 9, 10):
        for k in (0, 74):
            ms = []
            for it in range(3):
                d.upload(I, Q)
                if k:
                    lib.wspr_debug_fano_load(k, 40000, mode)       # ~4 x 120 ms: outlasts the decode
                ms.append(d.decode())
                t1 = time.perf_counter()
                lib.wspr_debug_fano_load(0, 0, 0)
                t2 = time.perf_counter()
            print("%-30s background CTAs %4d: decode ms %s   (load drained %.0f ms after the decode)"
                  % (MODES[mode], k, [round(x, 1) for x in ms], (t2 - t1) * 1e3), flush=True)
