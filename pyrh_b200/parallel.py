"""Column sharding across the GPUs of one box.

Columns are independent (the reference farms whole atmospheres to processes:
tests/test_mpi.py:1-15), so the data path needs no collective: every rank synthesises its own
contiguous block of columns with its own context.  ``torch.distributed`` is only plumbing here
(barrier, timing reduction, optional gather of the spectra to rank 0): backend "nccl" on GPUs,
"gloo" in the CPU tests.
"""
from __future__ import annotations

import numpy as np


def column_shard(ncol: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [first, first + count) of rank `rank`; blocks differ by at most one
    column and tile [0, ncol) exactly."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(int(ncol), int(world))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def gather_spectra(local: np.ndarray, ncol: int, dst: int = 0):
    """Gather per-rank ``[count, 4, nlambda]`` blocks into ``[ncol, 4, nlambda]`` on rank `dst`
    (None elsewhere).  No-op without an initialised process group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local
    rank, world = dist.get_rank(), dist.get_world_size()
    use_cuda = dist.get_backend() == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    counts = [column_shard(ncol, r, world)[1] for r in range(world)]
    mine = torch.from_numpy(np.ascontiguousarray(local)).to(dev)
    # gather wants equal sizes on every rank: pad to the largest block
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(mine.shape[1:]), dtype=mine.dtype, device=dev)
    pad[: mine.shape[0]] = mine
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    return np.concatenate([b[:c].cpu().numpy() for b, c in zip(bufs, counts)], axis=0)


def max_over_ranks(x: float) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return x
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
