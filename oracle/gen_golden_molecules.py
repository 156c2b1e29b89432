"""TEST INFRASTRUCTURE ONLY (oracle).  Golden vectors for LTE molecular background lines.

Runs the compiled, unmodified reference on FAL-C (B = 1 kG, v_z != 0, FULL_STOKES) around 847.3 nm, where
the CN B-X list the reference ships (rh/Molecules/CN/CN_B-X_SCIP_CH1.asc, 99 lines) has lines, and
records every MolecularOpacity() call that found a line (rh/opacity.c:711-839) together with the
molecule's n, pf, vbroad and its line table.  Output: tests/golden/falc_molecules.npz.
Usage: python -m oracle.gen_golden_molecules
"""
import numpy as np

from oracle import refdriver as rd
from oracle import portdriver as pd
from oracle.gen_golden import GOLD, recs_by_tag, one, falc_case_atm

ML_NFIELD = 16


def main():
    atm = falc_case_atm()
    wave = np.linspace(846.9, 847.8, 46)
    cwd = rd.make_workdir("benchmark")
    o = rd.rhf1d(atm, wave, cwd, probe=rd.PROBE_RLK | rd.PROBE_SNAP)
    R = recs_by_tag(o["records"])
    N = atm.shape[1]
    mols = sorted(R["mol_col"], key=lambda x: x[0][0])
    mol_index = {m[0]: i for i, (m, _) in enumerate(mols)}
    rows, zq, zs, zt = [], [], [], []
    for m, d in sorted(R["mol_line"], key=lambda x: (x[0][0], x[0][1])):
        nc = m[2]
        r = np.zeros(ML_NFIELD)
        r[:9] = d[:9]
        r[9], r[10], r[11] = mol_index[m[0]], len(zq), nc
        zq += list(d[10:10 + nc].astype(int)); zs += list(d[10 + nc:10 + 2 * nc]); zt += list(d[10 + 2 * nc:10 + 3 * nc])
        rows.append(r)
    fl = one(R, "flags")
    out = dict(atmosphere=atm, wave=wave, lam_spect=one(R, "lambda"), muz=one(R, "muz"), flags=fl,
               mlines=np.array(rows), zq=np.array(zq, np.int32), zshift=np.array(zs), zstrength=np.array(zt),
               mol=np.array([d.reshape(3, N) for _, d in mols]),
               mol_meta=np.array([m[:6] for m, _ in R["mol"]], np.int32),       # nspect, mu, to_obs, hasline, ispol, ns
               mol_chi_eta=np.array([d.reshape(2, -1, N) for _, d in R["mol"]]),
               stokes=np.array([o["I"], o["Q"], o["U"], o["V"]]), lam_out=o["lam"])
    for f in ("T", "vel", "B", "cos_gamma", "cos_2chi", "sin_2chi", "ne", "vturb", "nHtot", "np", "height"):
        out["col_" + f] = one(R, f)
    np.savez_compressed(GOLD / "falc_molecules.npz", **out)
    ok = 0
    for m, d in zip(out["mol_meta"], out["mol_chi_eta"]):
        chi, eta, flg = pd.molecular_opacity(out["mlines"], out["zq"], out["zshift"], out["zstrength"], fl[6],
                                             out["lam_spect"][m[0]], float(out["muz"][m[1]]), bool(fl[0]), int(m[2]),
                                             out["col_T"], out["col_vel"], out["col_B"], out["col_cos_gamma"],
                                             out["col_cos_2chi"], out["col_sin_2chi"], out["mol"])
        ok += np.array_equal(chi, d[0]) and np.array_equal(eta, d[1]) and flg == (m[3] | (m[4] << 1))
    print(f"[golden] falc_molecules: {len(rows)} lines of {len(mols)} molecule(s), {len(out['mol_meta'])} calls with a line; "
          f"port exact {ok}/{len(out['mol_meta'])} -> {(GOLD / 'falc_molecules.npz').stat().st_size/1e3:.0f} kB")


def main_polarizable():
    """The same window with a polarizable copy of the CN list (refdriver.polarizable_cn_tree): MolZeeman patterns,
    MolProfile's Zeeman sum, Q/U/V background opacity from molecules.  Output: tests/golden/falc_molecules_pol.npz."""
    import os
    import tempfile
    atm = falc_case_atm()
    wave = np.linspace(846.9, 847.8, 46)
    cwd = rd.make_workdir("benchmark")
    rd.load()
    keep = os.environ["PYRH_PATH"]
    os.environ["PYRH_PATH"] = rd.polarizable_cn_tree(tempfile.mkdtemp(prefix="rhref_pyrhpath_"))
    try:
        o = rd.rhf1d(atm, wave, cwd, probe=rd.PROBE_RLK | rd.PROBE_SNAP)
    finally:
        os.environ["PYRH_PATH"] = keep
    R = recs_by_tag(o["records"])
    zq, zs, zt, rows = [], [], [], []
    for m, d in sorted(R["mol_line"], key=lambda x: (x[0][0], x[0][1])):
        nc = m[2]
        rows.append(np.concatenate([d[:9], [len(zq), nc]]))
        zq += list(d[10:10 + nc].astype(int)); zs += list(d[10 + nc:10 + 2 * nc]); zt += list(d[10 + 2 * nc:10 + 3 * nc])
    out = dict(atmosphere=atm, wave=wave, stokes=np.array([o["I"], o["Q"], o["U"], o["V"]]), lam_out=o["lam"],
               mlines=np.array(rows), zq=np.array(zq, np.int32), zshift=np.array(zs), zstrength=np.array(zt))
    np.savez_compressed(GOLD / "falc_molecules_pol.npz", **out)
    ref = dict(np.load(GOLD / "falc_molecules.npz"))["stokes"]
    print(f"[golden] falc_molecules_pol: {len(rows)} lines, {len(zq)} Zeeman components; max |V/I| {np.max(np.abs(out['stokes'][3] / out['stokes'][0])):.3e}; "
          f"differs from the unpolarizable run: {not np.array_equal(ref, out['stokes'])}")


if __name__ == "__main__":
    import sys
    if "--polarizable" in sys.argv:
        main_polarizable()
    else:
        main()
