"""Debug fixture: per-iteration populations of the unmodified reference on config-5 sample column COL -> /root/repo/gpurun_in_nlte_iter.npz"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import bench
from oracle import refdriver as rd
from oracle.gen_golden import recs_by_tag
from pyrh_b200 import synthetic
bench._pyrh_data_path()
col = int(sys.argv[1])
c = bench.NLTE_CASES["config5_sample"]
atm = synthetic.perturbed_batch(np.load(ROOT / "tests/golden/falc_base.npy"), 1, ndep=bench.NDEP, first=10000 + col)[0]
wave = np.linspace(*c["wave"])
o = rd.rhf1d(atm, wave, bench._nlte_workdir("config5_sample"), probe=rd.PROBE_NLTE, get_populations=True)
R = recs_by_tag(o["records"])
its = sorted({m[1] for m, _ in R["up_n"]})
N = atm.shape[1]
n_iter = np.array([np.concatenate([d[:-1].reshape(-1, N) for m, d in sorted(R["up_n"], key=lambda x: x[0][0]) if m[1] == it]) for it in its])
dp = np.array([[d[-1] for m, d in sorted(R["up_n"], key=lambda x: x[0][0]) if m[1] == it][0] for it in its])
print("iterations", len(its), "dpops", dp[:5], dp[-5:], "min n over iterations", n_iter.reshape(len(its), -1).min(1)[:80])
np.savez_compressed(ROOT / "tools" / f"_nlte_iter_col{col}.npz", n_iter=n_iter, dpops=dp, atm=atm, wave=wave, n_final=np.concatenate([o["pops"][k]["n"] for k in o["pops"]]))
