"""TEST INFRASTRUCTURE ONLY (oracle).  Reads the numeric tables that the reference keeps as function-static
arrays in rh/hydrogen.c and rh/ohchbf.c (published cross-section tables: Geltman 1962, Stilley & Callaway 1970,
Bell 1980, Bates 1952, Victor & Dalgarno 1969, Kurucz et al. 1987) so that fixtures can hand them to the
library through its C ABI exactly like the RH host would.  Nothing is copied into the
repository's sources: the values only live in tests/golden/*.npz next to the outputs they produced."""
import re
from pathlib import Path

import numpy as np

REF = Path("/root/reference/rh")


def _functions(text):
    """split a reference source file at its '/* ---- begin ---- NAME.c' markers"""
    parts = re.split(r"/\* -+ begin -+ (\w+)\.c -+ \*/", text)
    return {parts[i]: parts[i + 1] for i in range(1, len(parts) - 1, 2)}


def _arrays(body):
    body = re.sub(r"/\*.*?\*/", " ", body, flags=re.S)
    out = {}
    for m in re.finditer(r"static\s+double\s+(\w+)\s*((?:\[[^\]]*\])+)\s*=\s*\{(.*?)\}\s*;", body, flags=re.S):
        vals = [float(x) for x in re.findall(r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?", m.group(3))]
        out[m.group(1)] = np.array(vals)
    return out


def continuum_tables():
    hyd = _functions((REF / "hydrogen.c").read_text())
    och = _functions((REF / "ohchbf.c").read_text())
    t = {}
    a = _arrays(hyd["Hminus_bf"]);   t["hmbf_lambda"], t["hmbf_alpha"] = a["lambdaBF"], a["alphaBF"]
    a = _arrays(hyd["Hminus_ff"]);   t["hmff_lambda"], t["hmff_theta"], t["hmff_kappa"] = a["lambdaFF"], a["thetaFF"], a["kappaFF"]
    a = _arrays(hyd["H2minus_ff"]);  t["h2mff_lambda"], t["h2mff_theta"], t["h2mff_kappa"] = a["lambdaFF"], a["thetaFF"], a["kappaFF"]
    a = _arrays(hyd["H2plus_ff"]);   t["h2pff_lambda"], t["h2pff_temp"], t["h2pff_kappa"] = a["lambdaFF"], a["tempFF"], a["kappaFF"]
    a = _arrays(hyd["Rayleigh_H2"]); t["rh2_a"], t["rh2_lambda"], t["rh2_sigma"] = a["a"], a["lambdaRH2"], a["sigma"]
    a = _arrays(och["OH_bf_opac"]);  t["oh_T"], t["oh_E"], t["oh_cross"] = a["TOH"], a["EOH"], a["OH_cross"]
    a = _arrays(och["CH_bf_opac"]);  t["ch_T"], t["ch_E"], t["ch_cross"] = a["TCH"], a["ECH"], a["CH_cross"]
    assert len(t["hmbf_lambda"]) == 34 and len(t["hmff_kappa"]) == len(t["hmff_lambda"]) * len(t["hmff_theta"])
    assert len(t["h2mff_kappa"]) == len(t["h2mff_lambda"]) * len(t["h2mff_theta"])
    assert len(t["h2pff_kappa"]) == len(t["h2pff_lambda"]) * len(t["h2pff_temp"])
    assert len(t["oh_cross"]) == len(t["oh_T"]) * len(t["oh_E"]) and len(t["ch_cross"]) == len(t["ch_T"]) * len(t["ch_E"])
    return t


if __name__ == "__main__":
    for k, v in continuum_tables().items():
        print(k, v.shape, v[:3], v[-2:])
