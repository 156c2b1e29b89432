"""world_size-2 gloo test of the N>1 plumbing (column shard + gather + timing reduction)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_column_shard_tiles_exactly():
    from pyrh_b200.parallel import column_shard
    for ncol in (0, 1, 7, 16384, 16385):
        for world in (1, 2, 3, 8):
            blocks = [column_shard(ncol, r, world) for r in range(world)]
            assert blocks[0][0] == 0
            for (f0, c0), (f1, _) in zip(blocks, blocks[1:]):
                assert f0 + c0 == f1
            assert blocks[-1][0] + blocks[-1][1] == ncol
            cs = [c for _, c in blocks]
            assert max(cs) - min(cs) <= 1


def test_two_rank_gloo_shard_and_gather(tmp_path):
    import json
    import subprocess
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ncol, world = 7, 2
    out = str(tmp_path / "res")
    procs = [subprocess.Popen([sys.executable, str(ROOT / "tests" / "helpers" / "gloo_worker.py"),
                               str(r), str(world), str(port), str(ncol), out]) for r in range(world)]
    for p in procs:
        assert p.wait(timeout=180) == 0
    meta = [json.loads(Path(out + f".rank{r}.json").read_text()) for r in range(world)]
    assert meta[0]["has_full"] and not meta[1]["has_full"]
    assert meta[0]["t"] == 2.0 and meta[1]["t"] == 2.0          # max over ranks
    full0 = np.load(out + ".rank0.npy")
    cols = np.arange(ncol, dtype=np.float64)
    want = cols[:, None, None] * 10.0 + np.arange(4)[None, :, None] + 0.001 * np.arange(5)[None, None, :]
    assert np.array_equal(full0, want)
    # allreduce callback of the wavelength shard: in-place SUM / MAX on raw host pointers
    for m in meta:
        assert m["rc"] == 0 and m["calls"] == 2
        assert m["sum"] == [3.0 * i for i in range(6)] and m["max"] == [2.0, 5.0]


def test_nlte_wavelength_shard_ranges_tile_and_balance():
    """rhb200_nlte_shard_range (host-only): chunks tile [0, Nspect) in order and carry a balanced number
    of rays (angle-dependent wavelengths count two directions per mu, formal.c:157-171)."""
    from conftest import GOLD
    from pyrh_b200 import nlte
    g = dict(np.load(GOLD / "nlte_caii.npz"))
    prob = nlte.NlteProblem.from_golden(g, ncol=1)
    Ns = prob.hdr["Nspect"]
    assert nlte.shard_range(prob, 0, 1) == (0, Ns)
    bb = np.zeros(Ns, bool)
    for ns in range(Ns):
        for e in range(prob.as_first[ns], prob.as_first[ns + 1]):
            bb[ns] |= prob.trans[prob.as_trans[e], nlte.TR_TYPE] == 0
    rays = len(prob.muz) * np.where(bb | (np.asarray(prob.bg_hasline) != 0), 2, 1)
    for world in (2, 3, 8):
        rng = [nlte.shard_range(prob, r, world) for r in range(world)]
        assert rng[0][0] == 0 and rng[-1][1] == Ns
        for (a, b), (c, d) in zip(rng, rng[1:]):
            assert b == c and a <= b
        per = [rays[a:b].sum() for a, b in rng]
        assert max(per) - min(per) <= 2 * rays.max() + rays.sum() // (10 * world)

