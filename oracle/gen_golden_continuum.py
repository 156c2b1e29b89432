"""TEST INFRASTRUCTURE ONLY (oracle).  Golden vectors for the background continuum.

Runs the compiled, unmodified reference on FAL-C WITHOUT Kurucz lines (empty kurucz.input) on a wavelength set
that crosses the branches of the continuum routines (UV to 2.3 micron, plus the Hinode window), and records
through oracle/probe.c (PROBE_CONT): the output of every contribution Background() sums (rh/background.c:343-465),
the totals chi_c / eta_c / sca_c it stores, and all inputs (level populations of the 11 model atoms, bound-free
continua with their tables, Rayleigh lines, molecular densities).  The function-static tables of hydrogen.c /
ohchbf.c are read from the reference sources by oracle/scrape_tables.py.
Output: tests/golden/falc_continuum.npz.   Usage: python -m oracle.gen_golden_continuum
"""
import numpy as np

from oracle import refdriver as rd
from oracle.gen_golden import GOLD, recs_by_tag, one, falc_case_atm
from oracle.scrape_tables import continuum_tables

NAMES = ["thomson", "hminus_bf", "hminus_ff", "oh_bf", "ch_bf", "h_bf", "h_ff", "rayleigh_h", "rayleigh_he",
         "h2plus_ff", "rayleigh_h2", "h2minus_ff", "metal_bf"]


def main():
    atm = falc_case_atm()
    wave = np.concatenate([[180.0, 250.0, 330.0, 364.0, 366.0, 420.0, 520.0], rd.air_to_vacuum(np.linspace(630.05, 630.35, 7)),
                           [820.0, 1100.0, 1500.0, 1640.0, 1700.0, 2300.0]])
    cwd = rd.make_workdir("benchmark", no_kurucz=True)
    o = rd.rhf1d(atm, wave, cwd, probe=rd.PROBE_CONT | rd.PROBE_BG | rd.PROBE_SNAP)
    R = recs_by_tag(o["records"])
    N = atm.shape[1]
    lam = one(R, "lambda")
    out = dict(atmosphere=atm, wave=wave, lam_spect=lam, names=np.array(NAMES))
    for k in ("hdr", "lev", "bf", "tab_lambda", "tab_alpha", "ray", "T", "ne", "nHmin", "nH2", "nOH", "nCH"):
        out["ct_" + k] = one(R, "ct_" + k)
    nlev = int(out["ct_hdr"][1])
    out["ct_lev"] = out["ct_lev"].reshape(nlev, 5)
    out["ct_bf"] = out["ct_bf"].reshape(-1, 10)
    out["ct_ray"] = out["ct_ray"].reshape(-1, 8)
    out["ct_n"] = one(R, "ct_n").reshape(nlev, N)
    out["ct_nstar"] = one(R, "ct_nstar").reshape(nlev, N)
    for k, v in continuum_tables().items():
        out["tab_" + k] = v
    # per-contribution outputs, ordered like lam_spect
    contrib = np.zeros((len(NAMES), len(lam), 2, N))
    okflag = np.zeros((len(NAMES), len(lam)), np.int32)
    for m, d in R["cont"]:
        if m[0] == 0:
            contrib[0, :, 0] = d[1:1 + N]; okflag[0] = 1
            continue
        l = int(np.flatnonzero(lam == d[0])[0])
        contrib[m[0], l] = d[1:].reshape(2, N)
        okflag[m[0], l] = m[1]
    out["contrib"], out["contrib_ok"] = contrib, okflag
    # totals: without lines chi_c = chi_ai etc. (one record per wavelength, written at the last mu/direction)
    tot = np.zeros((len(lam), 3, N))
    for m, d in R["bg"]:
        ns = m[3]
        tot[m[0], 0], tot[m[0], 1], tot[m[0], 2] = d[:N], d[ns * N:ns * N + N], d[2 * ns * N:2 * ns * N + N]
    out["total"] = tot
    # a background LINE (passive atom or molecule) adds to chi_c at some wavelengths: the totals are pure
    # continuum only where Background() found none
    out["hasline"] = one(R, "backgrflags").reshape(len(lam), 2)[:, 0].astype(np.int32)
    np.savez_compressed(GOLD / "falc_continuum.npz", **out)
    print(f"[golden] falc_continuum: {len(lam)} wavelengths, {nlev} levels, {len(out['ct_bf'])} continua, "
          f"contributions present per wavelength: {okflag.sum(axis=0).tolist()} "
          f"-> {(GOLD / 'falc_continuum.npz').stat().st_size/1e3:.0f} kB")


if __name__ == "__main__":
    main()
