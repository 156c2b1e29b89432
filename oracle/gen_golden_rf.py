"""TEST INFRASTRUCTURE ONLY (oracle).  Golden vectors for the analytic log gf response function.

Runs the compiled, unmodified reference with get_atomic_rfs = 1, loggf_ids = [0, 1] (both Fe I lines of
benchmark/fe6300) in NO_STOKES mode on FAL-C (v_z != 0) and records, per wavelength, the two
Piecewise_Bezier3_1D passes of Formal() (formal.c:167-283) together with spectrum.dchi_c_lam /
deta_c_lam (kurucz.c:696-699) and the propagated dI (bezier_1D.c:416-516), plus what rhf1d() returns in
mySpectrum.rfs (pyrh_solveray.c:144-147).  Output: tests/golden/falc_rf.npz.
Usage: python -m oracle.gen_golden_rf
"""
import numpy as np

from oracle import refdriver as rd
from oracle import portdriver as pd
from oracle.gen_golden import GOLD, recs_by_tag, one, falc_case_atm


def main():
    atm = falc_case_atm()
    wave = rd.air_to_vacuum(np.linspace(630.05, 630.35, 31))
    cwd = rd.make_workdir("benchmark", keywords={"STOKES_MODE": "NO_STOKES"})
    ids, vals = [0, 1], [-0.718, -0.968]
    o = rd.rhf1d(atm, wave, cwd, probe=rd.PROBE_ALL, loggf_ids=ids, loggf_values=vals, get_atomic_rfs=True)
    R = recs_by_tag(o["records"])
    N = atm.shape[1]
    bez = {(m[0], m[2]): d.reshape(4, N) for m, d in R["bez"]}
    ns_list, dn, up, Iin, dchi, deta, dI = [], [], [], [], [], [], []
    for m, d in R["bezrf"]:
        ns, npar = m[0], m[2]
        ns_list.append(ns)
        dn.append(bez[(ns, 0)][:3]); up.append(bez[(ns, 1)][:3])
        Iin.append(d[:N])
        dchi.append(d[N:N + N * npar].reshape(N, npar))
        deta.append(d[N + N * npar:N + 2 * N * npar].reshape(N, npar))
        dI.append(d[N + 2 * N * npar:].reshape(N, npar))
    lam = one(R, "lambda")
    out = dict(atmosphere=atm, wave=wave, loggf_ids=np.array(ids, np.int32), loggf_values=np.array(vals),
               lam_spect=lam, ns=np.array(ns_list, np.int32), muz=one(R, "muz"), col_T=one(R, "T"),
               col_height=one(R, "height"), down=np.array(dn), up=np.array(up), I_in=np.array(Iin),
               dchi=np.array(dchi), deta=np.array(deta), dI=np.array(dI), rfs=o["rfs"], I_spec=o["I"],
               lam_out=o["lam"])
    np.savez_compressed(GOLD / "falc_rf.npz", **out)
    ok = 0
    for i, ns in enumerate(ns_list):
        I, d = pd.bezier3_scalar_rf(out["col_height"], float(out["muz"][0]), out["up"][i, 0], out["up"][i, 1],
                                    out["col_T"], lam[ns], out["I_in"][i], out["dchi"][i], out["deta"][i])
        ok += np.array_equal(I, out["up"][i, 2]) and np.array_equal(d, out["dI"][i])
    keep = lam != 500.0
    print(f"[golden] falc_rf: {len(ns_list)} wavelengths x {out['dchi'].shape[2]} parameters; port exact {ok}/{len(ns_list)}; "
          f"rfs == dI[:, 0]: {np.array_equal(out['dI'][keep[out['ns']]][:, 0], o['rfs'])}; "
          f"stale I == down-ray: {np.array_equal(out['I_in'], out['down'][:, 2])} "
          f"-> {(GOLD / 'falc_rf.npz').stat().st_size/1e3:.0f} kB")


if __name__ == "__main__":
    main()
