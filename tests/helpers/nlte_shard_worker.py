"""One rank of the wavelength-sharded NLTE solve (torchrun, backend nccl, one GPU per rank).

Every rank holds the full CaII problem of tests/golden/nlte_caii.npz, formally solves its own chunk of
wavelengths and all-reduces Gamma / rates over NCCL once per MALI iteration (include/rhb200.h,
rhb200_nlte_set_shard).  Rank 0 compares with the reference's recorded result and writes a JSON report.
"""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import torch
    import torch.distributed as dist
    from pyrh_b200 import nlte, parallel
    from pyrh_b200.api import Context
    out = Path(sys.argv[1])
    ncol = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    g = dict(np.load(ROOT / "tests" / "golden" / "nlte_caii.npz"))
    prob = nlte.NlteProblem.from_golden(g, ncol=ncol)
    prob.J0 = np.zeros_like(prob.J0)
    ctx = Context(local)
    native = os.environ.get("RHB200_SHARD_NATIVE_NCCL", "0") != "0"     # NCCL called by the library (dlopen), no callback
    if native:
        parallel.shard_nlte_native(ctx)
        stats = {"calls": 0, "bytes": 0}
    else:
        stats = parallel.shard_nlte(ctx)
    lo, hi = nlte.shard_range(prob, rank, world)
    nscat = int(g["hdr"][11])
    res = nlte.iterate(ctx, prob, nscatter=nscat)             # warm-up + result
    dist.barrier(); torch.cuda.synchronize()
    calls0 = stats["calls"]
    t0 = time.perf_counter()
    res = nlte.iterate(ctx, prob, nscatter=nscat)
    torch.cuda.synchronize(); dist.barrier()
    dt = parallel.max_over_ranks(time.perf_counter() - t0)
    # all ranks must hold identical populations (the reduced rates are identical on every rank)
    n = torch.from_numpy(res["n"].copy()).cuda()
    nmax, nmin = n.clone(), n.clone()
    dist.all_reduce(nmax, op=dist.ReduceOp.MAX); dist.all_reduce(nmin, op=dist.ReduceOp.MIN)
    same = bool(torch.equal(nmax, nmin))
    if rank == 0:
        rel = float(np.max(np.abs(res["n"][0] / g["n_final"] - 1)))
        relJ = float(np.nanmax(np.abs(res["J"][0] / g["J_final"] - 1)))
        rep = dict(world=world, ncol=ncol, exchange="native ncclAllReduce group on the compute stream" if native else "torch.distributed callback", niter=int(res["niter"][0]), niter_ref=int(g["niter"]),
                   pops_max_rel_vs_reference=rel, J_max_rel_vs_reference=relJ, ranks_identical=same,
                   shard_rank0=[lo, hi], Nspect=int(g["hdr"][0]), seconds=dt,
                   allreduce_calls_per_solve=stats["calls"] - calls0,
                   allreduce_bytes_per_solve=stats["bytes"] // max(1, stats["calls"]) * (stats["calls"] - calls0))
        out.parent.mkdir(parents=True, exist_ok=True)
        out.write_text(json.dumps(rep, indent=1))
        print(json.dumps(rep))
        assert same and rep["niter"] == rep["niter_ref"] and rel < 1e-6 and relJ < 1e-6
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
