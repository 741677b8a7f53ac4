"""Multi-GPU plumbing: captures (and raw streams) are independent units, so N GPUs = N independent per-device batches
with NO collective on the data path.  One process per GPU (torchrun); torch.distributed is used
only to agree on timings and to collect the small result records on rank 0 after the decode."""
import os

import numpy as np


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(total, rank, world):
    """Contiguous shard [lo, hi) of `total` units for `rank`: ceil(total/world) per rank, the tail ranks may be short/empty."""
    per = (total + world - 1) // world
    lo = min(rank * per, total)
    return lo, min(lo + per, total)


def init_process_group(backend=None):
    """Initialise torch.distributed from the torchrun environment (nccl when CUDA is available, else gloo)."""
    import torch
    import torch.distributed as dist
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def max_over_ranks(value, device=None):
    """MAX of a python float over all ranks (identity when not distributed)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def decode_sharded(source, decode_fn, total=None):
    """Decode a corpus sharded over the ranks of the default process group.

    source: (I, Q), this rank's view of the WHOLE corpus (float32[total, samples]; only the local shard is touched), or a
    callable (lo, hi) -> (I, Q) producing just the shard.  decode_fn(I_shard, Q_shard) -> (spots[n, 100] RESULT_DTYPE, n_results[n]).
    Returns (spots, n_results) for the whole corpus on rank 0 (None elsewhere); no collective runs before every rank
    has finished its own decode."""
    import torch
    import torch.distributed as dist
    from .wsprd import RESULT_DTYPE, MAX_UNIQUES
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    rank, world = (dist.get_rank(), dist.get_world_size()) if distributed else (0, 1)
    total = int(total if total is not None else len(source[0]))
    lo, hi = shard_range(total, rank, world)
    Is, Qs = source(lo, hi) if callable(source) else (source[0][lo:hi], source[1][lo:hi])
    if hi > lo:
        spots, n = decode_fn(Is, Qs)
    else:
        spots, n = np.zeros((0, MAX_UNIQUES), RESULT_DTYPE), np.zeros(0, np.int32)
    if not distributed:
        return spots, n
    per = (total + world - 1) // world
    pad_spots = np.zeros((per, MAX_UNIQUES), RESULT_DTYPE)
    pad_n = np.zeros(per, np.int32)
    pad_spots[: hi - lo], pad_n[: hi - lo] = spots, n
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    ts = torch.from_numpy(np.frombuffer(pad_spots.tobytes(), dtype=np.uint8).copy()).to(dev)
    tn = torch.from_numpy(pad_n).to(dev)
    gs = [torch.empty_like(ts) for _ in range(world)]
    gn = [torch.empty_like(tn) for _ in range(world)]
    dist.all_gather(gs, ts)
    dist.all_gather(gn, tn)
    if rank != 0:
        return None, None
    out_s = np.zeros((world * per, MAX_UNIQUES), RESULT_DTYPE)      # (np.concatenate would repack the padded record dtype)
    for r, g in enumerate(gs):
        out_s[r * per:(r + 1) * per] = np.frombuffer(g.cpu().numpy().tobytes(), dtype=RESULT_DTYPE).reshape(per, MAX_UNIQUES)
    out_s = out_s[:total]
    out_n = np.concatenate([g.cpu().numpy() for g in gn])[:total]
    return out_s, out_n
