"""TEST INFRASTRUCTURE ONLY (oracle).  get_atomic_rfs through the whole rhf1d() call: mySpectrum.rfs (pyrh_solveray.c:
144-147) = the analytic log gf response function of the emergent intensity (kurucz.c:696-699, bezier_1D.c:416-516,
formal.c:278-282) for
  ns      benchmark/fe6300, STOKES_MODE = NO_STOKES, both Fe I lines, mu = 1
  ns_mu   the same with RLK_SCATTER = TRUE at mu = 0.7
  l4016   benchmark/lines_4016, NO_STOKES, three of the 18 lines
  fs      FULL_STOKES: the polarised solver carries no dI, the reference returns zeros there
Every case is run twice and must repeat (the reference reads uninitialised dchi_c_lam where a registered line is
outside a wavelength's window; such entries are recorded in `*_defined` = False and not compared).
Output: tests/golden/loggf_rf.npz.   Usage: python -m oracle.gen_golden_loggf_rf
"""
from pathlib import Path

import numpy as np

from oracle import refdriver as rd
from oracle.gen_golden import GOLD


def run(atm, wave, cwd, mu, ids, vals):
    rd.rhf1d(atm, wave, cwd, mu=mu, loggf_ids=ids, loggf_values=vals, get_atomic_rfs=True)      # warm-up
    a = rd.rhf1d(atm, wave, cwd, mu=mu, loggf_ids=ids, loggf_values=vals, get_atomic_rfs=True)
    b = rd.rhf1d(atm, wave, cwd, mu=mu, loggf_ids=ids, loggf_values=vals, get_atomic_rfs=True)
    same = (a["rfs"] == b["rfs"]) | (np.isnan(a["rfs"]) & np.isnan(b["rfs"]))
    assert np.array_equal(a["I"], b["I"])
    return np.array([a[s] for s in "IQUV"]), a["rfs"], same


def main():
    g = np.load(GOLD / "synth70_c0.npz")
    atm = g["atmosphere"]
    wave = rd.air_to_vacuum(np.linspace(630.08, 630.32, 49))
    out = dict(atmosphere=atm, wave=wave)
    ids, vals = [0, 1], [-0.718, -0.968]
    cwd = rd.make_workdir("benchmark", keywords={"STOKES_MODE": "NO_STOKES"})
    st, rf, ok = run(atm, wave, cwd, 1.0, ids, vals)
    out.update(ns_stokes=st, ns_rfs=rf, ns_defined=ok, ids=np.array(ids, np.int32), vals=np.array(vals))
    print("[golden] loggf_rf/ns: repeatable", ok.all(), "max |rf|", np.abs(rf).max(), "zeros", (rf == 0).sum())
    cwd = rd.make_workdir("benchmark", keywords={"STOKES_MODE": "NO_STOKES", "RLK_SCATTER": "TRUE"})
    st, rf, ok = run(atm, wave, cwd, 0.7, ids, vals)
    out.update(ns_mu_stokes=st, ns_mu_rfs=rf, ns_mu_defined=ok)
    print("[golden] loggf_rf/ns_mu: repeatable", ok.all(), "max |rf|", np.abs(rf).max())
    cwd = rd.make_workdir("benchmark", keywords={"STOKES_MODE": "NO_STOKES"})
    (Path(cwd) / "kurucz.input").write_text("lines_4016\n")
    w2 = np.linspace(401.55, 401.85, 61)
    ids2, vals2 = [2, 7, 11], [-1.0, -0.5, -2.0]
    st, rf, ok = run(atm, w2, cwd, 1.0, ids2, vals2)
    out.update(l4016_wave=w2, l4016_stokes=st, l4016_rfs=rf, l4016_defined=ok, l4016_ids=np.array(ids2, np.int32),
               l4016_vals=np.array(vals2))
    print("[golden] loggf_rf/l4016: repeatable", ok.all(), "of", ok.size, "max |rf| per parameter", np.abs(rf).max(axis=0))
    cwd = rd.make_workdir("benchmark")
    st, rf, ok = run(atm, wave, cwd, 1.0, ids, vals)
    out.update(fs_stokes=st, fs_rfs=rf)
    print("[golden] loggf_rf/fs: all zero", not rf.any())
    np.savez_compressed(GOLD / "loggf_rf.npz", **out)
    print(f"-> {(GOLD / 'loggf_rf.npz').stat().st_size / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
