"""TEST INFRASTRUCTURE ONLY (oracle).  A many-line Kurucz list through the reference: benchmark/lines_4016
(18 lines of Fe, Co, Ti, V, Ni, Mn, Nd, Ce ... in two ionisation stages around 401.7 nm, five of them with term
labels RLKdeterminate cannot read, i.e. NOT polarizable -> VoigtArmstrong branch, kurucz.c:824) on benchmark
column 1, mu = 1 and mu = 0.7.  Three kinds of wavelengths occur (formal.c:84-103): polarised line (Stokes DELO),
unpolarised line only (scalar Bezier ray in a moving column), no line (Feautrier).
Output: tests/golden/lines4016.npz.   Usage: python -m oracle.gen_golden_lines4016
"""
from pathlib import Path

import numpy as np

from oracle import refdriver as rd
from oracle.gen_golden import GOLD, recs_by_tag, one


def main():
    g = np.load(GOLD / "synth70_c1.npz")
    atm = g["atmosphere"]
    cwd = rd.make_workdir("benchmark")
    (Path(cwd) / "kurucz.input").write_text("lines_4016\n")
    wave = np.linspace(401.45, 401.90, 181)
    rd.rhf1d(atm, wave, cwd)                                   # warm-up (see gen_golden_rf_fd.py)
    out = dict(atmosphere=atm, wave=wave)
    for name, mu in (("mu1", 1.0), ("mu07", 0.7)):
        o = rd.rhf1d(atm, wave, cwd, mu=mu, probe=rd.PROBE_SNAP)
        R = recs_by_tag(o["records"])
        out[name + "_stokes"] = np.array([o[s] for s in "IQUV"])
        out[name + "_backgrflags"] = one(R, "backgrflags").reshape(-1, 2).astype(np.int32)
        out["lam"] = o["lam"]
    static = atm.copy()
    static[3] = 0.0                                             # VMACRO_TRESH = 0: still "moving" (|0| >= 0)
    cwd2 = rd.make_workdir("benchmark", keywords={"VMACRO_TRESH": "0.1"})
    (Path(cwd2) / "kurucz.input").write_text("lines_4016\n")
    o = rd.rhf1d(static, wave, cwd2, probe=rd.PROBE_SNAP)       # static column: unpolarised-line wavelengths -> Feautrier
    out["static_stokes"] = np.array([o[s] for s in "IQUV"])
    out["static_atmosphere"] = static
    f = out["mu1_backgrflags"]
    print("[golden] lines4016: classes (hasline, ispolarized):",
          {(h, p): int(np.sum((f[:, 0] == h) & (f[:, 1] == p))) for h in (0, 1) for p in (0, 1)},
          "depth", 1 - out["mu1_stokes"][0].min() / out["mu1_stokes"][0].max())
    np.savez_compressed(GOLD / "lines4016.npz", **out)


if __name__ == "__main__":
    main()
