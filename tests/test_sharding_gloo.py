"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: contiguous shards, no data-path collective, results gathered
on rank 0 in corpus order.  The per-shard decoder is injected (the oracle stands in for the GPU, which CI does not have)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from rtlsdr_wsprd_b200 import sharding


def test_shard_range_covers_everything_once():
    for total in (0, 1, 7, 8, 100000, 4096):
        for world in (1, 2, 3, 4, 8):
            got = []
            for r in range(world):
                lo, hi = sharding.shard_range(total, r, world)
                assert 0 <= lo <= hi <= total
                got.extend(range(lo, hi))
            assert got == list(range(total))
    assert sharding.shard_range(100000, 7, 8) == (87500, 100000)


def _worker(rank, world, port, tmp):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import torch.distributed as dist
    from oracle import pyoracle as po
    import helpers as H
    from rtlsdr_wsprd_b200.wsprd import RESULT_DTYPE, MAX_UNIQUES
    sharding.init_process_group("gloo")
    total = 5                                              # odd: shards of 3 and 2
    seen = []

    def gen(lo, hi):
        seen.append((lo, hi))
        I, Q, _ = H.make_corpus(2, hi - lo, start=lo)
        return I, Q

    def decode_fn(I, Q):
        spots = np.zeros((len(I), MAX_UNIQUES), RESULT_DTYPE)
        n = np.zeros(len(I), np.int32)
        for c in range(len(I)):
            r, _, _ = po.decode(po.oracle(), I[c], Q[c])
            spots[c, : len(r)], n[c] = r, len(r)
        return spots, n

    spots, n = sharding.decode_sharded(gen, decode_fn, total=total)
    assert seen[0] == sharding.shard_range(total, rank, world)
    t = sharding.max_over_ranks(float(rank + 1))
    s = sharding.sum_over_ranks(float(rank + 1))
    assert t == world and s == world * (world + 1) / 2
    if rank == 0:
        np.save(os.path.join(tmp, "n.npy"), n)
        open(os.path.join(tmp, "spots.bin"), "wb").write(np.ascontiguousarray(spots).tobytes())
    else:
        assert spots is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_decode_matches_single_process(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    from oracle import pyoracle as po
    import helpers as H
    from rtlsdr_wsprd_b200.wsprd import RESULT_DTYPE
    n = np.load(os.path.join(str(tmp_path), "n.npy"))
    spots = np.frombuffer(open(os.path.join(str(tmp_path), "spots.bin"), "rb").read(), dtype=RESULT_DTYPE).reshape(5, -1)
    I, Q, _ = H.make_corpus(2, 5)
    for c in range(5):
        r, _, _ = po.decode(po.oracle(), I[c], Q[c])
        assert n[c] == len(r) and H.results_equal(r, spots[c, : n[c]])
