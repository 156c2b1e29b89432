"""TEST INFRASTRUCTURE ONLY (oracle).  Keyword RLK_SCATTER = TRUE (kurucz.c:641-652, 682-694): the Kurucz lines'
opacity is split into a thermal and a scattering part.  Benchmark column 0, Hinode window (both Fe I lines, neutral
stage: x = 0.68) and lines_4016 (neutral and singly ionised lines, unpolarizable ones included).
Output: tests/golden/rlkscatter.npz.   Usage: python -m oracle.gen_golden_rlkscatter
"""
from pathlib import Path

import numpy as np

from oracle import refdriver as rd
from oracle.gen_golden import GOLD


def main():
    g = np.load(GOLD / "synth70_c0.npz")
    atm, wave = g["atmosphere"], g["wave"]
    cwd = rd.make_workdir("benchmark", keywords={"RLK_SCATTER": "TRUE"})
    rd.rhf1d(atm, wave, cwd)
    o = rd.rhf1d(atm, wave, cwd)
    out = dict(atmosphere=atm, hinode_wave=wave, hinode_stokes=np.array([o[s] for s in "IQUV"]))
    print("[golden] rlkscatter/hinode: max rel change of I vs RLK_SCATTER = FALSE:",
          np.abs(out["hinode_stokes"][0] / g["stokes_scalar"][0] - 1).max())
    cwd2 = rd.make_workdir("benchmark", keywords={"RLK_SCATTER": "TRUE", "N_MAX_SCATTER": "3"})
    (Path(cwd2) / "kurucz.input").write_text("lines_4016\n")
    w2 = np.linspace(401.45, 401.90, 91)
    o = rd.rhf1d(atm, w2, cwd2, mu=0.8)
    out.update(l4016_wave=w2, l4016_stokes=np.array([o[s] for s in "IQUV"]))
    np.savez_compressed(GOLD / "rlkscatter.npz", **out)


if __name__ == "__main__":
    main()
