"""TEST INFRASTRUCTURE ONLY (oracle).  pyrh.get_scales (rhf1d/pyrh_hse.c:402-553) of the unmodified reference on
benchmark column 0 for the three depth scales; lam_ref = 500 nm.  Output: tests/golden/get_scales.npz.
Usage: python -m oracle.gen_golden_get_scales
"""
import ctypes as C
import os

import numpy as np

from oracle import refdriver as rd
from oracle.gen_golden import GOLD

dp = C.POINTER(C.c_double)


def get_scales(atm, atm_scale, cwd, lam_ref=500.0, variant="scalar"):
    lib = rd.load(variant)
    lib.get_scales.restype = None
    lib.get_scales.argtypes = [C.c_char_p, C.c_int] + [dp] * 6 + [C.c_int, C.c_double] + [dp] * 3 + [C.c_int, C.c_void_p, C.c_void_p]
    a = np.ascontiguousarray(atm, np.float64).copy()
    rows = [np.ascontiguousarray(a[i]) for i in (0, 1, 2, 3, 4, 8)]
    n = a.shape[1]
    tau, height, cmass = np.full(n, np.nan), np.full(n, np.nan), np.full(n, np.nan)
    old = os.getcwd()
    os.chdir(cwd)
    try:
        lib.get_scales(str(cwd).encode(), n, *[r.ctypes.data_as(dp) for r in rows], int(atm_scale), float(lam_ref),
                       tau.ctypes.data_as(dp), height.ctypes.data_as(dp), cmass.ctypes.data_as(dp), 0, None, None)
    finally:
        os.chdir(old)
    return tau, height, cmass


def main():
    g = np.load(GOLD / "synth70_c0.npz")
    atm = g["atmosphere"]
    cwd = rd.make_workdir("benchmark")
    rd.rhf1d(atm, g["wave"], cwd)                       # warm-up call
    out = {}
    tau0, h0, c0 = get_scales(atm, 0, cwd)
    out["tau_atmosphere"], out["tau_height"], out["tau_cmass"] = atm, h0, c0          # tau is not returned for TAU500
    a1 = atm.copy(); a1[0] = np.log10(c0 / (1.0e-3 / (1.0e-2 * 1.0e-2)))
    t1, h1, c1 = get_scales(a1, 1, cwd)
    out["cmass_atmosphere"], out["cmass_height"], out["cmass_tau"] = a1, h1, t1
    a2 = atm.copy(); a2[0] = h0 / 1.0e3
    t2, h2, c2 = get_scales(a2, 2, cwd)
    out["height_atmosphere"], out["height_tau"], out["height_cmass"] = a2, t2, c2
    print("[golden] get_scales: height[0] =", h0[0], h1[0], "tau[-1] =", t1[-1], t2[-1], "cmass[-1] =", c0[-1], c2[-1])
    np.savez_compressed(GOLD / "get_scales.npz", **out)


if __name__ == "__main__":
    main()
