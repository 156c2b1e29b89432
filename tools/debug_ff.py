import os, sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
os.environ["RHB200_NLTE_EXACT"] = "1"
from oracle import refdriver as rd
from pyrh_b200 import nlte_host
case = sys.argv[1] if len(sys.argv) > 1 else "h_caii_r5_ff"
g = np.load(ROOT / "tests/golden/nlte_front.npz")
c = json.loads(str(g["cases"]))[case]
os.environ["PYRH_PATH"] = str(rd.REFDIR / "pyrh_path")
cwd = rd.make_workdir("tests", keywords=c["kw"], atoms_active=tuple(c["active"]), atoms_extra=(("CaII.atom", "ACTIVE"),))
atm, wave, mu = g["atmosphere"], g[f"{case}_wave"], float(g[f"{case}_mu"])
s = nlte_host.NlteSession(cwd, wave)
res = s.compute(atm[:2], mu=mu)
lam = s.wavelengths
I, Iref = res["I"][0], g[f"{case}_I"][0]
quv, qref = np.stack([res["Q"][0], res["U"][0], res["V"][0]]), g[f"{case}_QUV"][0]
eI = np.abs(I / Iref - 1)
bad = np.nonzero(eI > 0)[0]
print("n wavelengths", len(lam), "differing in I:", len(bad), "QUV nonzero ref:", int((np.abs(qref).max(0) > 0).sum()), "ours:", int((np.abs(quv).max(0) > 0).sum()))
tr = s.plan["trans"]
pol = s.line_pol
full = s.lam
user = full != s.lambda_ref
idx = np.nonzero(user)[0]
def flags(ns):
    out = []
    for t in tr:
        if t[1] == 0 and t[4] <= ns < t[4] + t[5]:
            out.append(("L%d%s" % (int(t[13]), "p" if pol[int(t[13])] else "")))
    return ",".join(out)
wf = np.zeros(len(full), np.int32)
from pyrh_b200 import _lib
_lib.check(s.ctx.lib.rhb200_get_wavelength_flags(s.ctx.h, wf.ctypes.data_as(_lib.ip)))
for b in bad[:40]:
    ns = idx[b]
    print(f"{lam[b]:.5f} ns {ns} eI {eI[b]:.2e} refQUV {qref[:, b]} ours {quv[:, b]} bgflags {wf[ns]} lines {flags(ns)}")
nz = np.nonzero((np.abs(qref).max(0) > 0) | (np.abs(quv).max(0) > 0))[0]
print("first wavelengths with QUV:", [(round(float(lam[b]), 4), wf[idx[b]], flags(idx[b]), bool(np.array_equal(quv[:, b], qref[:, b]))) for b in nz[:30]])
s.close()
