"""TEST INFRASTRUCTURE ONLY (oracle).  N_MAX_SCATTER > 0 in LTE (pyrh_compute1dray.c:332-337): after Iterate() the
reference Lambda-iterates the continuum scattering term of the angle-independent (line-free) wavelengths,
S = (eta + sca J) / chi (formal.c:289-309), until max |1 - Jdag/J| <= ITER_LIMIT or N_MAX_SCATTER passes.
Benchmark column 2 on a grid reaching well outside the line windows, for ITER_LIMIT = 1e-2 (keyword file) and 1e-4.
Output: tests/golden/scatter.npz.   Usage: python -m oracle.gen_golden_scatter
"""
import numpy as np

from oracle import refdriver as rd
from oracle.gen_golden import GOLD


def main():
    atm = np.load(GOLD / "synth70_c2.npz")["atmosphere"]
    wave = np.linspace(629.0, 631.5, 51)
    out = dict(atmosphere=atm, wave=wave)
    base = None
    for name, kw in (("n0", {"N_MAX_SCATTER": "0"}), ("n5", {"N_MAX_SCATTER": "5"}),
                     ("n5_tight", {"N_MAX_SCATTER": "5", "ITER_LIMIT": "1.0E-4"}), ("n1", {"N_MAX_SCATTER": "1", "ITER_LIMIT": "1.0E-6"})):
        cwd = rd.make_workdir("benchmark", keywords=kw)
        rd.rhf1d(atm, wave, cwd)
        o = rd.rhf1d(atm, wave, cwd)
        st = np.array([o[s] for s in "IQUV"])
        out[name + "_stokes"] = st
        out["lam"] = o["lam"]
        if base is None:
            base = st
        print(f"[golden] scatter/{name}: max rel change of I vs N_MAX_SCATTER = 0: {np.abs(st[0] / base[0] - 1).max():.3e}, "
              f"wavelengths changed: {int(np.sum(st[0] != base[0]))}")
    np.savez_compressed(GOLD / "scatter.npz", **out)


if __name__ == "__main__":
    main()
