"""TEST INFRASTRUCTURE ONLY (oracle).  Golden vectors for the reference's non-Bezier formal solvers.

Runs the compiled, unmodified reference (oracle/_ref) on FAL-C (B = 1 kG, v_z != 0) with
  S_INTERPOLATION        = S_LINEAR | S_PARABOLIC      (STOKES_MODE = NO_STOKES)   -> Piecewise_Linear_1D / Piecewise_1D
  S_INTERPOLATION_STOKES = DELO_PARABOLIC              (STOKES_MODE = FULL_STOKES) -> Piece_Stokes_1D
and records every solver call (rh/rhf1d/piecewise_1D.c:44,134; piecestokes_1D.c:49) with the --wrap
probe.  Output: tests/golden/falc_solvers.npz.   Usage: python -m oracle.gen_golden_solvers
"""
import numpy as np

from oracle import refdriver as rd
from oracle import portdriver as pd
from oracle.gen_golden import GOLD, recs_by_tag, one, falc_case_atm


def main():
    atm = falc_case_atm()
    wave = rd.air_to_vacuum(np.linspace(630.05, 630.35, 31))
    nd = atm.shape[1]
    out = dict(atmosphere=atm, wave=wave)
    for tag, kw in (("lin", {"STOKES_MODE": "NO_STOKES", "S_INTERPOLATION": "S_LINEAR"}),
                    ("par", {"STOKES_MODE": "NO_STOKES", "S_INTERPOLATION": "S_PARABOLIC"}),
                    ("pst", {"S_INTERPOLATION_STOKES": "DELO_PARABOLIC"})):
        cwd = rd.make_workdir("benchmark", keywords=kw)
        # the polarised run uses the grid of fixture falc_B1kG, whose recorded line table and background
        # are then the inputs of the fused device path with the other solver
        w = rd.hinode_wave() if tag == "pst" else wave
        full = rd.rhf1d(atm, w, cwd, variant="scalar", probe=rd.PROBE_ALL)
        R = recs_by_tag(full["records"])
        recs = R[tag]
        if tag == "pst":
            recs = [r for r in recs if r[0][2] == 1][::10] + [r for r in recs if r[0][2] == 0][::30]
            out["pst_lam_spect"] = one(R, "lambda")
        nrow = 13 if tag == "pst" else 4
        out[tag + "_meta"] = np.array([m[:4] for m, _ in recs], np.int32)      # nspect, mu, to_obs, has Psi
        out[tag] = np.array([d.reshape(nrow, nd) for _, d in recs])
        out[tag + "_spec"] = np.array([full["I"], full["Q"], full["U"], full["V"]])
        if "lam_spect" not in out:
            out["lam_spect"], out["muz"] = one(R, "lambda"), one(R, "muz")
            out["col_T"], out["col_height"] = one(R, "T"), one(R, "height")
            out["lam_out"] = full["lam"]
        print(f"[golden] falc_solvers/{tag}: {len(recs)} rays "
              f"(down {sum(1 for m, _ in recs if m[2] == 0)}, up {sum(1 for m, _ in recs if m[2] == 1)})")
    # Psi (approximate operator diagonal) is only requested in NLTE runs: two MALI iterations of the
    # CaII fixture with each scalar solver, a sample of the up- and down-ray calls that carry Psi
    from oracle.gen_golden_nlte import KW
    atm_n = rd.falc("tests")
    atm_n[5] = 500.0
    wave_n = np.linspace(630.25, 630.5, 21)
    for tag, interp in (("lin", "S_LINEAR"), ("par", "S_PARABOLIC")):
        kw = dict(KW, N_MAX_ITER=2, S_INTERPOLATION=interp)
        cwd = rd.make_workdir("tests", keywords=kw, atoms_extra=(("CaII.atom", "ACTIVE"),))
        o = rd.rhf1d(atm_n, wave_n, cwd, probe=rd.PROBE_BEZ | rd.PROBE_SNAP, get_populations=True)
        out[tag + "_nlte_n"] = o["pops"]["CA"]["n"]          # CaII populations after 2 MALI iterations
        R = recs_by_tag(o["records"])
        recs = [(m, d) for m, d in R[tag] if m[3] == 1][::97]
        ndn = atm_n.shape[1]
        out[tag + "psi_meta"] = np.array([m[:4] for m, _ in recs], np.int32)
        out[tag + "psi"] = np.array([d.reshape(4, ndn) for _, d in recs])
        if "n_lam_spect" not in out:
            out["n_lam_spect"], out["n_muz"] = one(R, "lambda"), one(R, "muz")
            out["n_T"], out["n_height"] = one(R, "T"), one(R, "height")
        print(f"[golden] falc_solvers/{tag}psi: {len(recs)} rays with Psi")
    np.savez_compressed(GOLD / "falc_solvers.npz", **out)
    print(f"[golden] falc_solvers -> {(GOLD / 'falc_solvers.npz').stat().st_size/1e6:.2f} MB")
    validate(out)


def validate(g):
    h, T, muz = g["col_height"], g["col_T"], g["muz"]
    for tag, kind in (("lin", "linear"), ("par", "parabolic")):
        ok = 0
        for m, d in zip(g[tag + "_meta"], g[tag]):
            I = pd.piecewise_scalar(kind, h, float(muz[m[1]]), int(m[2]), d[0], d[1], T, g["lam_spect"][m[0]])
            ok += np.array_equal(I, d[2])
        print(f"[port-vs-ref] {tag}: exact {ok}/{len(g[tag])}")
        ok = 0
        for m, d in zip(g[tag + "psi_meta"], g[tag + "psi"]):
            I, Psi = pd.piecewise_scalar(kind, g["n_height"], float(g["n_muz"][m[1]]), int(m[2]), d[0], d[1],
                                         g["n_T"], g["n_lam_spect"][m[0]], want_psi=True)
            ok += np.array_equal(I, d[2]) and np.array_equal(Psi, d[3])
        print(f"[port-vs-ref] {tag}psi: exact {ok}/{len(g[tag + 'psi'])}")
    ok = 0
    for m, d in zip(g["pst_meta"], g["pst"]):
        I = pd.stokes_parabolic(h, float(muz[m[1]]), int(m[2]), d[0], d[1:5], d[10:13], T, g["pst_lam_spect"][m[0]])
        ok += np.array_equal(I, d[5:9])
    print(f"[port-vs-ref] pst: exact {ok}/{len(g['pst'])}")


if __name__ == "__main__":
    main()
