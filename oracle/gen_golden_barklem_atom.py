"""TEST INFRASTRUCTURE ONLY (oracle).  BARKLEM van der Waals broadening of neutral model-atom lines
(readatom.c:311-320 -> getBarklemactivecross, barklem.c:216-312 -> VanderWaals, broad.c:125-136): atoms.input with
MgI_6level.atom in place of Mg.atom (Mg b triplet, 3s3p 3P - 3s4s 3S: the s-p table) and CaI.atom added as a twelfth
atom (Ca I 422.7 nm, 4s2 1S - 4s4p 1P).  Benchmark column 1, both windows in one call each.
Output: tests/golden/barklem_atom.npz.   Usage: python -m oracle.gen_golden_barklem_atom
"""
from pathlib import Path

import numpy as np

from oracle import refdriver as rd
from oracle.gen_golden import GOLD


def stage():
    cwd = rd.make_workdir("benchmark", atoms_extra=(("CaI.atom", "PASSIVE"),))
    p = Path(cwd) / "atoms.input"
    p.write_text(p.read_text().replace("Mg.atom ", "MgI_6level.atom "))
    return cwd


def main():
    atm = np.load(GOLD / "synth70_c1.npz")["atmosphere"]
    cwd = stage()
    print(open(cwd + "/atoms.input").read())
    out = dict(atmosphere=atm)
    rd.rhf1d(atm, rd.hinode_wave(), cwd)
    for name, wave in (("Mgb", np.linspace(516.6, 518.6, 101)), ("CaI", np.linspace(422.5, 423.0, 81))):
        o = rd.rhf1d(atm, wave, cwd)
        o2 = rd.rhf1d(atm, wave, cwd)
        assert np.array_equal(o["I"], o2["I"])
        out[name + "_wave"] = wave
        out[name + "_stokes"] = np.array([o[s] for s in "IQUV"])
        print(f"[golden] barklem_atom/{name}: depth {1 - o['I'].min() / o['I'].max():.3f}")
    np.savez_compressed(GOLD / "barklem_atom.npz", **out)


if __name__ == "__main__":
    main()
