import sys, os, json, numpy as np
sys.path.insert(0, "/root/repo")
os.environ["RHB200_NLTE_EXACT"] = "1"
from oracle import refdriver as rd
g = np.load("/root/repo/tests/golden/nlte_front.npz")
case = "h_caii_r3"
c = json.loads(str(g["cases"]))[case]
cwd = rd.make_workdir("tests", keywords=c["kw"], atoms_active=tuple(c["active"]), atoms_extra=(("CaII.atom", "ACTIVE"),))
atm, wave, mu = g["atmosphere"], g[f"{case}_wave"], float(g[f"{case}_mu"])
o = rd.rhf1d(atm[1], wave, cwd, mu=mu, get_populations=True, variant="bridged")
n = np.concatenate([o["pops"][k]["n"] for k in c["keys"]]); ns = np.concatenate([o["pops"][k]["nstar"] for k in c["keys"]])
rel = lambda a, b: np.max(np.abs(a / b - 1), axis=-1)
print("I maxrel", rel(o["I"], g[f"{case}_I"][1]))
print("n per level", rel(n, g[f"{case}_n"][1]))
print("nstar per level", rel(ns, g[f"{case}_nstar"][1]))
b = rd.rhf1d_batch(atm[1:2], wave, cwd, mu=mu, get_populations=True, nlev=n.shape[0])
print("niter", b["niter"], g[f"{case}_niter"][1])
