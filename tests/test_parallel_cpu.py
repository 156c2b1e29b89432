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
