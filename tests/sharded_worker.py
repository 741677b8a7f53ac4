"""Worker of test_sharded_decode_on_two_gpus: one rank per GPU (torchrun), contiguous shards of a small config-3 corpus decoded
by the CUDA path, records gathered on rank 0 through sharding.decode_sharded.  usage: sharded_worker.py out.npz ncaptures"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rtlsdr_wsprd_b200 as w
from rtlsdr_wsprd_b200 import sharding
import helpers as H

out, total = sys.argv[1], int(sys.argv[2])
rank, world, local = sharding.init_process_group()


def source(lo, hi):
    I, Q, _ = H.make_corpus(3, hi - lo, start=2300 + lo)
    return I, Q


def decode(I, Q):
    with w.BatchDecoder(len(I), device=local) as d:
        d.upload(I, Q)
        d.decode()
        return d.download()


spots, n = sharding.decode_sharded(source, decode, total=total)
if rank == 0:
    np.savez(out, spots=np.frombuffer(spots.tobytes(), np.uint8), n=n)
import torch.distributed as dist
dist.barrier()
dist.destroy_process_group()
