"""TEST INFRASTRUCTURE ONLY (oracle).  Golden vectors for the pyrh-unit entry point (compute1d boundary).

The reference's rhf1d() is run on the 70-depth benchmark base column 0 (fixture synth70_c0) with
  * atm_scale = 0 (log10 tau_500), 1 (log10 column mass [g cm^-2]) and 2 (height [km])
    (pyrh_compute1dray.c:230-246, convertScales: rhf1d/multiatmos.c:100-177), mu = 1;
  * atm_scale = 0 at mu = 0.6 (Bproject's inclined branch, rhf1d/project.c:60-77).
Recorded per case: the pyrh-unit input rows, the SI rows the reference derives (height, tau_ref, cmass, np,
cos_gamma, cos_2chi, sin_2chi), the abundance sums convertScales uses (abundance.c:219-221) and the spectrum.
Output: tests/golden/pyrh_scales.npz.   Usage: python -m oracle.gen_golden_scales
"""
import numpy as np

from oracle import refdriver as rd
from oracle.gen_golden import GOLD, recs_by_tag, one

ROWS = ("T", "ne", "vturb", "vel", "B", "cos_gamma", "cos_2chi", "sin_2chi", "nHtot", "np", "height", "tau_ref", "cmass")


def run(atm, wave, cwd, atm_scale, mu):
    o = rd.rhf1d(atm, wave, cwd, mu=mu, atm_scale=atm_scale, probe=rd.PROBE_SNAP)
    R = recs_by_tag(o["records"])
    d = {f: one(R, f) for f in ROWS}
    d["stokes"] = np.array([o[k] for k in "IQUV"])
    d["abund_sums"] = one(R, "abund_sums")
    d["lambda"] = one(R, "lambda")
    d["flags"] = one(R, "flags")
    return d


def main():
    g = np.load(GOLD / "synth70_c0.npz")
    atm0, wave = g["atmosphere"], g["wave"]
    cwd = rd.make_workdir("benchmark")
    out = dict(wave=wave)
    base = run(atm0, wave, cwd, 0, 1.0)
    assert np.array_equal(base["stokes"], g["stokes_scalar"]) and np.array_equal(base["height"], g["col_height"])
    cases = [("tau", atm0, 0, 1.0), ("tau_mu06", atm0, 0, 0.6)]
    a1 = atm0.copy()
    a1[0] = np.log10(base["cmass"] / (1.0e-3 / (1.0e-2 * 1.0e-2)))      # kg m^-2 -> g cm^-2
    cases.append(("cmass", a1, 1, 1.0))
    a2 = atm0.copy()
    a2[0] = base["height"] / 1.0e3
    cases.append(("height", a2, 2, 1.0))
    for name, atm, sc, mu in cases:
        d = run(atm, wave, cwd, sc, mu)
        out[name + "_atmosphere"] = atm
        out[name + "_scale_mu"] = np.array([sc, mu])
        for k, v in d.items():
            out[f"{name}_{k}"] = v
        print(f"[golden] pyrh_scales/{name}: atm_scale={sc} mu={mu} Ic={d['stokes'][0].max():.4e} "
              f"height[0]={d['height'][0]:.6e} wght_per_H={d['abund_sums'][0]:.10f}")
    np.savez_compressed(GOLD / "pyrh_scales.npz", **out)
    print(f"[golden] pyrh_scales -> {(GOLD / 'pyrh_scales.npz').stat().st_size/1e3:.0f} kB")


if __name__ == "__main__":
    main()
