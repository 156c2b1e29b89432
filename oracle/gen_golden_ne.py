"""TEST INFRASTRUCTURE ONLY (oracle).  pyrh.get_ne_from_nH (rhf1d/pyrh_hse.c:555-677) and pyrh.hse (:67-400) of the
unmodified reference.  Inputs: T and nH of benchmark columns 0 and 1 (get_ne_from_nH); log tau500 and T of column 0
with pg_top = 0.1 ... (hse).  Output: tests/golden/ne_hse.npz.   Usage: python -m oracle.gen_golden_ne
"""
import ctypes as C
import os

import numpy as np

from oracle import refdriver as rd
from oracle.gen_golden import GOLD

dp = C.POINTER(C.c_double)


def _call(cwd, fn):
    old = os.getcwd()
    os.chdir(cwd)
    try:
        fn()
    finally:
        os.chdir(old)


def get_ne_from_nH(cwd, atm_scale, scale, T, nH, variant="scalar"):
    lib = rd.load(variant)
    lib.get_ne_from_nH.restype = None
    lib.get_ne_from_nH.argtypes = [C.c_char_p, C.c_int, C.c_int, dp, dp, dp, dp]
    scale, T, nH = (np.ascontiguousarray(x, np.float64).copy() for x in (scale, T, nH))
    ne = np.full(len(T), np.nan)
    _call(cwd, lambda: lib.get_ne_from_nH(str(cwd).encode(), int(atm_scale), len(T), scale.ctypes.data_as(dp),
                                          T.ctypes.data_as(dp), nH.ctypes.data_as(dp), ne.ctypes.data_as(dp)))
    return ne


def hse(cwd, atm_scale, scale, T, pg_top, fudge_wave=None, fudge_value=None, variant="scalar"):
    lib = rd.load(variant)
    lib.hse.restype = None
    lib.hse.argtypes = [C.c_char_p, C.c_int, dp, dp, dp, dp, dp, dp, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                        C.c_int, C.c_void_p, C.c_void_p]
    scale, T = (np.ascontiguousarray(x, np.float64).copy() for x in (scale, T))
    n = len(T)
    ne, nH, rho, pg = (np.full(n, np.nan) for _ in range(4))
    pg[0] = pg_top
    nf = 0 if fudge_wave is None else len(fudge_wave)
    fw = np.ascontiguousarray(fudge_wave if nf else [0.0], np.float64)
    fv = np.ascontiguousarray(fudge_value if nf else [0.0], np.float64)
    _call(cwd, lambda: lib.hse(str(cwd).encode(), n, scale.ctypes.data_as(dp), T.ctypes.data_as(dp),
                               ne.ctypes.data_as(dp), nH.ctypes.data_as(dp), rho.ctypes.data_as(dp),
                               pg.ctypes.data_as(dp), int(atm_scale), nf, fw.ctypes.data_as(dp) if nf else None,
                               fv.ctypes.data_as(dp) if nf else None, 0, None, None))
    return ne, nH, rho, pg


def main():
    cwd = rd.make_workdir("benchmark")
    out = {}
    for c in (0, 1):
        a = np.load(GOLD / f"synth70_c{c}.npz")["atmosphere"]
        out[f"c{c}_T"], out[f"c{c}_nH"] = a[1], a[8]
        out[f"c{c}_ne"] = get_ne_from_nH(cwd, 0, a[0], a[1], a[8])
        print(f"[golden] get_ne_from_nH c{c}: ne/ne_atmos at tau=1:", out[f"c{c}_ne"][56] / a[2][56])
    a = np.load(GOLD / "synth70_c0.npz")["atmosphere"]
    for name, pg_top in (("hse01", 0.1), ("hse1", 1.0)):
        ne, nH, rho, pg = hse(cwd, 0, a[0], a[1], pg_top)
        out[name + "_scale"], out[name + "_T"], out[name + "_pgtop"] = a[0], a[1], np.float64(pg_top)
        out[name + "_ne"], out[name + "_nH"], out[name + "_rho"], out[name + "_pg"] = ne, nH, rho, pg
        print(f"[golden] hse pg_top={pg_top}: pg[-1] = {pg[-1]:.5e} nH[-1] = {nH[-1]:.5e} ne[-1] = {ne[-1]:.5e}")
    fw = np.array([400.0, 500.0, 700.0])
    fv = np.array([[1.3, 1.2, 1.0], [1.0, 1.5, 1.0], [0.7, 1.4, 1.0]])
    ne, nH, rho, pg = hse(cwd, 0, a[0], a[1], 0.1, fw, fv)
    out.update(hsef_wave=fw, hsef_value=fv, hsef_ne=ne, hsef_nH=nH, hsef_rho=rho, hsef_pg=pg)
    print(f"[golden] hse with fudge: pg[-1] = {pg[-1]:.5e} (without: {out['hse01_pg'][-1]:.5e})")
    np.savez_compressed(GOLD / "ne_hse.npz", **out)


if __name__ == "__main__":
    main()
