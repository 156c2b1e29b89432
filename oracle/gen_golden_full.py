"""TEST INFRASTRUCTURE ONLY (oracle).  Inputs of the fused LTE path WITH the continuum on the device.

Same run as fixture falc_B1kG (FAL-C, B = 1 kG, Hinode grid, the two Fe I lines): records, next to what that
fixture already holds, the continuum model (all model atoms, bound-free edges, Rayleigh lines), what
ChemicalEquilibrium() did to the atomic populations (ntotal before / after, rh/chemequil.c:336-342), nHmin and
the molecular densities.  Expected outputs are falc_B1kG's (chi_ai, eta_ai, stokes_scalar).
Output: tests/golden/falc_full.npz.   Usage: python -m oracle.gen_golden_full
"""
import numpy as np

from oracle import refdriver as rd
from oracle.gen_golden import GOLD, recs_by_tag, one, falc_case_atm
from oracle.scrape_tables import continuum_tables


def main():
    atm = falc_case_atm()
    wave = rd.hinode_wave()
    cwd = rd.make_workdir("benchmark")
    o = rd.rhf1d(atm, wave, cwd, probe=rd.PROBE_CONT | rd.PROBE_SNAP)
    R = recs_by_tag(o["records"])
    N = atm.shape[1]
    out = dict(atmosphere=atm, wave=wave, lam_spect=one(R, "lambda"))
    for k in ("hdr", "lev", "bf", "tab_lambda", "tab_alpha", "ray", "T", "ne", "nHmin", "nH2", "nOH", "nCH"):
        out["ct_" + k] = one(R, "ct_" + k)
    nlev, natom = int(out["ct_hdr"][1]), int(out["ct_hdr"][0])
    out["ct_lev"] = out["ct_lev"].reshape(nlev, 5)
    out["ct_bf"] = out["ct_bf"].reshape(-1, 10)
    out["ct_ray"] = out["ct_ray"].reshape(-1, 8)
    out["ct_nstar"] = one(R, "ct_nstar").reshape(nlev, N)
    for k, v in continuum_tables().items():
        out["tab_" + k] = v
    pre, post = one(R, "ce_ntotal_pre").reshape(natom, N), one(R, "ce_ntotal_post").reshape(natom, N)
    out["abundance"] = one(R, "ce_abundance")
    out["fraction"] = post / pre                       # the reference's own division (chemequil.c:336)
    out["chem"] = np.concatenate([out["fraction"], [out["ct_nHmin"], out["ct_nH2"], out["ct_nOH"], out["ct_nCH"]]])
    out["stokes"] = np.array([o["I"], o["Q"], o["U"], o["V"]])
    # chemical network (for ChemicalEquilibrium on the device)
    out["ce_nuclei"] = one(R, "ce_nuclei").reshape(-1, 2)
    out["ce_mol"] = np.array([d for m, d in sorted(R["ce_mol"], key=lambda x: x[0][0])])
    out["col_nHtot"] = one(R, "nHtot")
    np.savez_compressed(GOLD / "falc_full.npz", **out)
    g = np.load(GOLD / "falc_B1kG.npz")
    print(f"[golden] falc_full: {natom} atoms, {nlev} levels; atoms rescaled by chemistry: "
          f"{int(np.sum(np.any(out['fraction'] != 1.0, axis=1)))}; spectrum equals falc_B1kG: "
          f"{np.array_equal(out['stokes'], g['stokes_scalar'])} -> {(GOLD / 'falc_full.npz').stat().st_size/1e3:.0f} kB")


def synth70():
    """chem inputs of the three 70-depth benchmark base columns (fixtures synth70_c0..2)."""
    from pyrh_b200 import synthetic
    base = np.load(GOLD / "falc_base.npy")
    wave = rd.hinode_wave()
    chems, ab = [], None
    for c in range(3):
        atm = synthetic.perturbed_batch(base, 1, first=c)[0]
        g = np.load(GOLD / f"synth70_c{c}.npz")
        assert np.array_equal(g["atmosphere"], atm)
        cwd = rd.make_workdir("benchmark")
        o = rd.rhf1d(atm, wave, cwd, probe=rd.PROBE_CONT | rd.PROBE_SNAP)
        R = recs_by_tag(o["records"])
        N = atm.shape[1]
        natom = int(one(R, "ct_hdr")[0])
        pre, post = one(R, "ce_ntotal_pre").reshape(natom, N), one(R, "ce_ntotal_post").reshape(natom, N)
        chems.append(np.concatenate([post / pre, [one(R, "ct_nHmin"), one(R, "ct_nH2"), one(R, "ct_nOH"), one(R, "ct_nCH")]]))
        ab = one(R, "ce_abundance")
        assert np.array_equal(np.array([o["I"], o["Q"], o["U"], o["V"]]), g["stokes_scalar"])
    np.savez_compressed(GOLD / "synth70_chem.npz", chem=np.array(chems), abundance=ab)
    print(f"[golden] synth70_chem: {np.array(chems).shape} -> {(GOLD / 'synth70_chem.npz').stat().st_size/1e3:.0f} kB")


if __name__ == "__main__":
    main()
    synth70()
