"""Debug / evidence run (not a test): the NLTE front end on the FAL-C fixtures -- every per-column input the device
hands to Iterate() against the reference's recorded state, then populations, iteration count and spectrum."""
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ["RHB200_NLTE_FRONT_DEBUG"] = "1"
from oracle import refdriver as rd          # noqa: E402  (stages the working directory; evidence script)
from pyrh_b200 import nlte_host as nh       # noqa: E402

os.environ["PYRH_PATH"] = str(rd.REFDIR / "pyrh_path")
KW = {"NRAYS": 3, "N_MAX_SCATTER": 2, "N_MAX_ITER": 50, "NG_ORDER": 2, "NG_DELAY": 10, "NG_PERIOD": 3,
      "ITER_LIMIT": "1.0E-4", "PRD_N_MAX_ITER": 0, "STOKES_MODE": "NO_STOKES", "HYDROGEN_LTE": "TRUE"}


def rel(a, b):
    d = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    d[(a == b)] = 0.0
    return float(np.max(d))


def run(name, keywords, atoms_active):
    g = np.load(ROOT / "tests" / "golden" / f"{name}.npz")
    cwd = rd.make_workdir("tests", keywords=keywords, atoms_active=atoms_active, atoms_extra=(("CaII.atom", "ACTIVE"),))
    if "atmosphere" in g:
        atm, wave, mu = g["atmosphere"], g["wave"], float(g["mu"])
    else:
        atm, wave, mu = rd.falc("tests"), np.linspace(630.25, 630.5, 21), 1.0
        atm[5] = 500.0
    s = nh.NlteSession(cwd, wave)
    res = s.compute(atm, mu=mu)
    N, Ns = atm.shape[1], len(s.lam)
    nlev, ngam, Na, nline = int(np.sum(g["atom_nlevel"])), int(np.sum(g["atom_nlevel"] ** 2)), len(g["atom_nlevel"]), len(g["adamp"])
    out = dict(fixture=name)
    for which, key, shape, ref in ((0, "C", (ngam, N), g["C"]), (1, "nstar", (nlev, N), g["nstar"]),
                                   (2, "ntotal", (Na, N), g["ntotal"]), (3, "adamp", (nline, N), g["adamp"]),
                                   (4, "vbroad", (Na, N), g["vbroad"]), (5, "chi_c", (Ns, N), g["bg"][0]),
                                   (6, "eta_c", (Ns, N), g["bg"][1]), (7, "sca_c", (Ns, N), g["bg"][2]),
                                   (8, "height", (N,), g["height"]), (9, "J_final", (Ns, N), g["J_final"]),
                                   (10, "fs_chi_c", (Ns, N), g["fs_bg"][0]), (11, "fs_eta_c", (Ns, N), g["fs_bg"][1]),
                                   (12, "fs_sca_c", (Ns, N), g["fs_bg"][2]), (13, "fs_phi", g["fs_phi"].shape, g["fs_phi"]),
                                   (14, "fs_wphi", g["fs_wphi"].shape, g["fs_wphi"]), (15, "fs_adamp", g["fs_adamp"].shape, g["fs_adamp"])):
        d = s.debug(which, shape)
        out[key] = dict(exact=bool(np.array_equal(d, ref)), maxrel=rel(d, ref))
    out["bg_hasline_exact"] = bool(np.array_equal(s.plan["bg_hasline"], g["bgflags"][:, 0]))
    out["niter"] = [int(res["niter"]), int(g["niter"])]
    out["n_final"] = dict(exact=bool(np.array_equal(res["n"], g["pops_final"])), maxrel=rel(res["n"], g["pops_final"]))
    out["spec_I"] = dict(exact=bool(np.array_equal(res["I"], g["spec_I"])), maxrel=rel(res["I"], g["spec_I"]))
    d = np.abs(res["I"] / g["spec_I"] - 1)
    out["spec_I_worst"] = [(float(s.wavelengths[i]), float(d[i])) for i in np.argsort(d)[-5:]]
    s.close()
    return out


if __name__ == "__main__":
    rep = [run("nlte_caii_pert", KW, ()), run("nlte_caii", KW, ()), run("nlte_h_caii", dict(KW, HYDROGEN_LTE="FALSE"), ("H_6.atom",))]
    os.makedirs(ROOT / "gpurun_out", exist_ok=True)
    (ROOT / "gpurun_out" / "nlte_front_check.json").write_text(json.dumps(rep, indent=1))
    print(json.dumps(rep, indent=1))
