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


# ---- wavelength sharding of ONE atmosphere (second mode): the exchange step -------

class _RawBuffer:
    """A raw pointer dressed as an array for torch: CUDA array interface for device memory, the numpy
    array interface for host memory (gloo tests)."""

    def __init__(self, ptr: int, count: int, cuda: bool):
        iface = dict(shape=(count,), typestr="<f8", data=(ptr, False), version=3, strides=None)
        if cuda:
            self.__cuda_array_interface__ = iface
        else:
            self.__array_interface__ = iface


def make_allreduce(group=None):
    """A ``rhb200_allreduce_fn`` (include/rhb200.h) that reduces the library's buffer in place with
    ``torch.distributed.all_reduce``: NCCL over NVLink on the GPUs (the pointer is device memory and is
    wrapped without a copy), gloo on host memory in the CPU tests.  Returns (callback, stats); keep the
    callback alive while it is registered with ``rhb200_nlte_set_shard``."""
    import torch
    import torch.distributed as dist
    from . import _lib
    cuda = dist.get_backend(group) == "nccl"
    stats = {"calls": 0, "bytes": 0}

    def reduce(_user, ptr, count, op):
        try:
            if cuda:
                t = torch.as_tensor(_RawBuffer(int(ptr), int(count), True), device=torch.device("cuda", torch.cuda.current_device()))
            else:
                t = torch.from_numpy(np.asarray(_RawBuffer(int(ptr), int(count), False)))
            dist.all_reduce(t, op=dist.ReduceOp.MAX if op == _lib.REDUCE_MAX else dist.ReduceOp.SUM, group=group)
            if cuda:
                torch.cuda.current_stream().synchronize()
            stats["calls"] += 1
            stats["bytes"] += 8 * int(count)
            return 0
        except Exception as e:          # never unwind through the C frame
            print(f"[pyrh_b200] allreduce callback failed: {e!r}", flush=True)
            return 1

    return _lib.ALLREDUCE_FN(reduce), stats


def shard_nlte_native(ctx, group=None):
    """The same with NCCL called by the library itself (rhb200_nlte_set_shard_nccl_id): torch.distributed only ships the
    128-byte ncclUniqueId of rank 0; the reductions run on the context's stream without callbacks or host syncs."""
    import ctypes as C
    import torch.distributed as dist
    from . import _lib
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    buf = C.create_string_buffer(128)
    if rank == 0:
        _lib.check(ctx.lib.rhb200_nccl_unique_id(buf))
    box = [buf.raw]
    dist.broadcast_object_list(box, src=0, group=group)
    uid = C.create_string_buffer(box[0], 128)
    _lib.check(ctx.lib.rhb200_nlte_set_shard_nccl_id(ctx.h, rank, world, uid))


def shard_nlte(ctx, group=None):
    """Register this rank's wavelength shard with the context: later ``nlte.iterate`` / ``nlte.formal``
    calls formally solve only this rank's wavelengths and all-reduce rates once per iteration."""
    import torch.distributed as dist
    from . import _lib
    fn, stats = make_allreduce(group)
    ctx._shard_fn = fn                                      # keep the ctypes thunk alive
    _lib.check(ctx.lib.rhb200_nlte_set_shard(ctx.h, dist.get_rank(group), dist.get_world_size(group), fn, None))
    return stats
