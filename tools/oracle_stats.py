"""Decode-path statistics from the instrumented oracle: how many candidates reach the jitter search, Fano timeouts, ...
usage: oracle_stats.py [config] [count]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pyoracle as po
import helpers as H
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32
orc = po.oracle()
st = (C.c_long * 16).in_dll(orc, "oracle_stat")
I, Q, _ = H.make_corpus(cfg, n, start=2000)
spots = 0
for c in range(n):
    r, _, _ = po.decode(orc, I[c], Q[c])
    spots += len(r)
names = ["cand pass0", "cand pass1+", "worth pass0", "worth pass1+", "fano calls", "fano timeouts", "decoded @jitter0", "decoded @jitter>0",
         "success cycles>4096", "success cycles>32768", "never decoded pass0", "never decoded pass1+", "sum winning idt"]
print("captures", n, "spots", spots)
for k, nm in enumerate(names):
    print("%-24s %8d  per capture %.3f" % (nm, st[k], st[k] / n))
