"""Worker of tests/test_parallel_cpu.py: one rank of a world_size-N gloo group."""
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))


def main():
    rank, world, port, ncol, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4]), sys.argv[5]
    import torch.distributed as dist
    from pyrh_b200.parallel import column_shard, gather_spectra, max_over_ranks
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = port
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = column_shard(ncol, rank, world)
    cols = np.arange(first, first + count, dtype=np.float64)
    local = cols[:, None, None] * 10.0 + np.arange(4)[None, :, None] + 0.001 * np.arange(5)[None, None, :]
    full = gather_spectra(local, ncol, dst=0)
    t = max_over_ranks(float(rank + 1))
    # the allreduce callback the wavelength-sharded NLTE solve hands to librhb200 (host buffer under gloo)
    import ctypes
    from pyrh_b200 import _lib
    from pyrh_b200.parallel import make_allreduce
    fn, stats = make_allreduce()
    buf = np.arange(6, dtype=np.float64) * (rank + 1)
    mx = np.array([1.0 + rank, 5.0 - rank])
    rc = fn(None, buf.ctypes.data_as(ctypes.c_void_p), buf.size, _lib.REDUCE_SUM)
    rc |= fn(None, mx.ctypes.data_as(ctypes.c_void_p), mx.size, _lib.REDUCE_MAX)
    dist.barrier()
    if full is not None:
        np.save(out + f".rank{rank}.npy", full)
    Path(out + f".rank{rank}.json").write_text(json.dumps({"t": t, "has_full": full is not None, "rc": rc, "sum": buf.tolist(),
                                                            "max": mx.tolist(), "calls": stats["calls"]}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
