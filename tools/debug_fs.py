import os, sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
os.environ["RHB200_NLTE_EXACT"] = "1"
from oracle import refdriver as rd
from pyrh_b200 import nlte_host
case = "caii_r3_fs"
g = np.load(ROOT / "tests/golden/nlte_front.npz")
c = json.loads(str(g["cases"]))[case]
os.environ["PYRH_PATH"] = str(rd.REFDIR / "pyrh_path")
atm, wave, mu = g["atmosphere"], g[f"{case}_wave"], float(g[f"{case}_mu"])
for nit in (1, 2, 3):
    kw = dict(c["kw"], N_MAX_ITER=nit)
    cwd = rd.make_workdir("tests", keywords=kw, atoms_active=tuple(c["active"]), atoms_extra=(("CaII.atom", "ACTIVE"),))
    s = nlte_host.NlteSession(cwd, wave)
    res = s.compute(atm[:1], mu=mu)
    s.close()
    o = rd.rhf1d(atm[0], wave, cwd, mu=mu, get_populations=True)
    nref = np.concatenate([o["pops"][k]["n"] for k in c["keys"]])
    n = res["n"][0]
    with np.errstate(all="ignore"):
        e = np.abs(n / nref - 1)
    print("N_MAX_ITER", nit, "niter", res["niter"], "finite ours", np.isfinite(n).all(), "ref", np.isfinite(nref).all(),
          "max rel n", np.nanmax(e), "per level", np.nanmax(e, axis=1), "worst depth", np.unravel_index(np.nanargmax(e), e.shape),
          "I maxrel", np.nanmax(np.abs(res["I"][0] / o["I"] - 1)))
