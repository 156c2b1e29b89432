"""TEST INFRASTRUCTURE ONLY (oracle).  A working directory whose atoms.input differs from the shipped one: CaII.atom
added as a twelfth PASSIVE atom; Ca II K window (393.2 - 393.7 nm, passive_bb line with van der Waals + quadratic
Stark damping) and the Hinode window, benchmark column 2.  Output: tests/golden/atoms12.npz.
Usage: python -m oracle.gen_golden_atoms12
"""
import numpy as np

from oracle import refdriver as rd
from oracle.gen_golden import GOLD


def main():
    atm = np.load(GOLD / "synth70_c2.npz")["atmosphere"]
    cwd = rd.make_workdir("benchmark", atoms_extra=(("CaII.atom", "PASSIVE"),))
    print(open(cwd + "/atoms.input").read())
    out = dict(atmosphere=atm)
    rd.rhf1d(atm, rd.hinode_wave(), cwd)
    for name, wave in (("CaK", np.linspace(393.2, 393.7, 101)), ("hinode", rd.hinode_wave())):
        o = rd.rhf1d(atm, wave, cwd)
        out[name + "_wave"], out[name + "_lam"] = wave, o["lam"]
        out[name + "_stokes"] = np.array([o[s] for s in "IQUV"])
        print(f"[golden] atoms12/{name}: depth {1 - o['I'].min() / o['I'].max():.3f}")
    base = np.load(GOLD / "synth70_c2.npz")["stokes_scalar"]
    print("hinode spectrum changed by the extra atom:", not np.array_equal(base, out["hinode_stokes"]),
          np.abs(out["hinode_stokes"][0] / base[0] - 1).max())
    np.savez_compressed(GOLD / "atoms12.npz", **out)


if __name__ == "__main__":
    main()
