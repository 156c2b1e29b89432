"""TEST INFRASTRUCTURE ONLY (oracle).  The reference's own test, tests/test_compute1d.py: FAL-C (tests/falc.dat through
its spinor2multi, here the committed falc_base.npy) with B = 500 G, wave = linspace(630.25, 630.5, 100), mu = 1,
atm_scale = 0, run in the reference's tests/ directory (keyword.input with STOKES_MODE = NO_STOKES).  Also a moving
and inclined variant (v_z = 2 km/s, mu = 0.5).  Output: tests/golden/ref_test_compute1d.npz.
Usage: python -m oracle.gen_golden_ref_test
"""
import numpy as np

from oracle import refdriver as rd
from oracle.gen_golden import GOLD


def main():
    atm = np.load(GOLD / "falc_base.npy").copy()
    atm[5] = 500
    wave = np.linspace(630.25, 630.5, num=100)
    cwd = rd.make_workdir("tests")
    rd.rhf1d(atm, wave, cwd)
    o = rd.rhf1d(atm, wave, cwd, probe=rd.PROBE_SNAP)
    from oracle.gen_golden import recs_by_tag
    rl = np.array([d[:32] for m, d in sorted(recs_by_tag(o["records"])["rlk_line"], key=lambda x: x[0][0])])
    out = dict(atmosphere=atm, wave=wave, lam=o["lam"], I=o["I"], Q=o["Q"], U=o["U"], V=o["V"],
               rlk_cross=rl[:, 17], rlk_alpha=rl[:, 18], rlk_vdwaals=rl[:, 20])       # getBarklemcross output
    a2 = atm.copy()
    a2[3] = 2.0
    o2 = rd.rhf1d(a2, wave, cwd, mu=0.5)
    out.update(moving_atmosphere=a2, moving_I=o2["I"])
    print("[golden] ref_test_compute1d: depth", 1 - o["I"].min() / o["I"].max(), "Q any:", bool(np.any(o["Q"])),
          "| moving mu=0.5 depth", 1 - o2["I"].min() / o2["I"].max())
    np.savez_compressed(GOLD / "ref_test_compute1d.npz", **out)


if __name__ == "__main__":
    main()
