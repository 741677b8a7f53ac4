"""Decimate a few full-length raw streams resident in HBM (for ncu captures of K0).  usage: profile_frontend.py [nstreams]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import rtlsdr_wsprd_b200 as w
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n_iq = 288_000_000
stride = 2 * n_iq + 16
raw = torch.randint(0, 256, (ns * stride,), dtype=torch.uint8, device="cuda")
I = torch.zeros((ns, 45000), dtype=torch.float32, device="cuda")
Q = torch.zeros_like(I)
for _ in range(4):
    n, ms = w.decimate_device(raw.data_ptr(), ns, n_iq, stride, I.data_ptr(), Q.data_ptr(), 45000, 45000, 0)
    print("streams", ns, "outputs", n, "ms", round(ms, 3), "GB/s", round(ns * 2 * n_iq / ms / 1e6, 1), flush=True)
