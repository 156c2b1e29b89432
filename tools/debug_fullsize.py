"""Debug: full-size batch vs the same columns computed alone / in small batches."""
import os, sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from oracle import refdriver as rd
from pyrh_b200 import host, synthetic
os.environ["PYRH_PATH"] = str(rd.REFDIR / "pyrh_path")
ncol, ndep = 16384, 70
cwd = rd.make_workdir("benchmark")
s = host.Session(cwd, rd.hinode_wave(301))
atm = synthetic.perturbed_batch(np.load(ROOT / "tests/golden/falc_base.npy"), ncol, ndep=ndep)
st = s.compute(atm)
st2 = s.compute(atm)
print("repeatable:", np.array_equal(st, st2))
pick = np.sort(np.random.default_rng(7).choice(ncol, 256, replace=False))
alone = s.compute(atm[pick])
d = np.abs(st[pick][:, 0] / alone[:, 0] - 1).max(axis=1)
bad = np.nonzero(d > 0)[0]
print("picked columns differing between full batch and sub-batch:", len(bad), pick[bad][:20], d[bad][:20])
# chunk-level: which columns of the full batch differ from a recompute in 2048-blocks
nbad = 0
for c0 in range(0, ncol, 2048):
    blk = s.compute(atm[c0:c0 + 2048])
    dd = np.nonzero(np.any(blk != st[c0:c0 + 2048], axis=(1, 2)))[0]
    nbad += len(dd)
    if len(dd): print("block", c0, "differs at", len(dd), "columns, first", dd[:10] + c0)
print("total differing vs 2048-blocks:", nbad)
ref = []
for c in pick[:8]:
    o = rd.rhf1d(atm[c], rd.hinode_wave(301), cwd)
    ref.append(np.array([o[k] for k in "IQUV"]))
ref = np.array(ref)
print("vs reference, first 8 picks: full", [bool(np.array_equal(st[c], r)) for c, r in zip(pick[:8], ref)],
      "alone", [bool(np.array_equal(a, r)) for a, r in zip(alone[:8], ref)])
s.close()
