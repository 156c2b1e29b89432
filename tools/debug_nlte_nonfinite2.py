"""Debug: columns 5, 8, 11 of the config-5 sample: exact vs fast rates vs the unmodified reference."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import bench
from oracle import refdriver as rd
from pyrh_b200 import nlte_host, synthetic
bench._pyrh_data_path()
case = "config5_sample"
c = bench.NLTE_CASES[case]
base = np.load(ROOT / "tests/golden/falc_base.npy")
cols = [5, 8, 11, 16, 1, 2]
atm = np.stack([synthetic.perturbed_batch(base, 1, ndep=bench.NDEP, first=10000 + b)[0] for b in cols])
wave = np.linspace(*c["wave"])
out = {}
for mode in ("1", "0"):
    os.environ["RHB200_NLTE_EXACT"] = mode
    s = nlte_host.NlteSession(bench._nlte_workdir(case), wave, 0)
    res = s.compute(atm)
    out[mode] = res
    print("exact" if mode == "1" else "fast ", "niter", res["niter"], "finite", np.isfinite(res["n"]).reshape(len(cols), -1).all(1))
    s.close()
for q, b in enumerate(cols):
    try:
        o = rd.rhf1d(atm[q], wave, bench._nlte_workdir(case), get_populations=True)
    except SystemExit:
        print("reference exits on", b); continue
    n_ref = np.concatenate([o["pops"][k]["n"] for k in o["pops"]])
    for mode in ("1", "0"):
        n = out[mode]["n"][q].reshape(n_ref.shape)
        with np.errstate(all="ignore"):
            print("col", b, "exact" if mode == "1" else "fast ", "max rel diff n vs reference", np.nanmax(np.abs(n / n_ref - 1)), "I equal", np.array_equal(out[mode]["I"][q][:len(o["I"])], o["I"]) if out[mode]["I"][q].shape[0] >= len(o["I"]) else None)
