"""TEST INFRASTRUCTURE ONLY (oracle).  Opacity fudge factors (pyrh.compute1d's fudge_wave / fudge_value): benchmark
column 0, Hinode window, factors for H-, scattering and metal bound-free opacity varying across the window.
Output: tests/golden/fudge.npz.   Usage: python -m oracle.gen_golden_fudge
"""
import numpy as np

from oracle import refdriver as rd
from oracle.gen_golden import GOLD


def main():
    g = np.load(GOLD / "synth70_c0.npz")
    atm, wave = g["atmosphere"], g["wave"]
    cwd = rd.make_workdir("benchmark")
    rd.rhf1d(atm, wave, cwd)
    fw = np.array([400.0, 500.0, 630.2, 630.35, 700.0])
    fv = np.array([[1.30, 1.10, 1.25, 0.90, 1.00], [1.0, 2.0, 1.5, 1.2, 1.0], [0.70, 1.40, 1.80, 0.60, 1.00]])
    o = rd.rhf1d(atm, wave, cwd, fudge_wave=fw, fudge_value=fv)
    st = np.array([o[s] for s in "IQUV"])
    print("[golden] fudge: max change of I:", np.abs(st[0] / g["stokes_scalar"][0] - 1).max())
    np.savez_compressed(GOLD / "fudge.npz", atmosphere=atm, wave=wave, fudge_wave=fw, fudge_value=fv, stokes=st, lam=o["lam"])


if __name__ == "__main__":
    main()
