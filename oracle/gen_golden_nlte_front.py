"""oracle/gen_golden_nlte_front.py -- TEST INFRASTRUCTURE ONLY.

Golden vectors for NLTE through the drop-in call on PERTURBED columns: the UNMODIFIED reference's rhf1d() with
get_populations on 8 synthetic perturbed FAL-C columns (70 depths, SURVEY 8(d) recipe, pyrh_b200.synthetic) for
    caii_r3 / caii_r5       Ca II (5 levels + continuum) ACTIVE, hydrogen PASSIVE in LTE, NRAYS = 3 / 5, Ca II 8542 window
    h_caii_r3 / h_caii_r5   H (6 levels) + Ca II ACTIVE (BASELINE config 4), NRAYS = 3 / 5, Hinode window
CRD (PRD_N_MAX_ITER = 0), Ng order 2, ITER_LIMIT 1e-4, NO_STOKES.  Recorded: spectrum, populations n / nstar of the
ACTIVE atoms, and the number of MALI iterations (counted from the probe's updatePopulations records).
    caii_r3_ff / h_caii_r5_ff   the same with STOKES_MODE = FIELD_FREE: field-free iterations, then adjustStokesMode()
                                and the full Stokes solution (I, Q, U, V recorded; the columns carry B up to 2.5 kG)
    caii_r3_pf                  STOKES_MODE = POLARIZATION_FREE: Zeeman-broadened profiles, scalar iterations, then full Stokes
    caii_r3_fs / h_caii_r3_fs   STOKES_MODE = FULL_STOKES: polarised profiles, rays and I_eff in every MALI iteration
    caii_r3_prd / caii_r5_prd1  angle-averaged PRD in Ca II H & K: PRD_N_MAX_ITER 3 / 1 (the shipped
                                keyword.input.NLTE has 1), user grid inside Ca II K for the first

    python -m oracle.gen_golden_nlte_front
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import refdriver as rd  # noqa: E402
from oracle.gen_golden import recs_by_tag  # noqa: E402
from pyrh_b200 import synthetic  # noqa: E402

GOLD = ROOT / "tests" / "golden"
KW = {"N_MAX_SCATTER": 2, "N_MAX_ITER": 100, "NG_ORDER": 2, "NG_DELAY": 10, "NG_PERIOD": 3,
      "ITER_LIMIT": "1.0E-4", "PRD_N_MAX_ITER": 0, "STOKES_MODE": "NO_STOKES"}
CASES = {
    "caii_r3": dict(kw=dict(KW, NRAYS=3, HYDROGEN_LTE="TRUE"), active=(), wave=(854.2, 854.7, 41), keys=("CA",), mu=1.0),
    "caii_r5": dict(kw=dict(KW, NRAYS=5, HYDROGEN_LTE="TRUE"), active=(), wave=(854.2, 854.7, 41), keys=("CA",), mu=0.8),
    "h_caii_r3": dict(kw=dict(KW, NRAYS=3, HYDROGEN_LTE="FALSE"), active=("H_6.atom",), wave=(630.25, 630.5, 21),
                      keys=("H ", "CA"), mu=1.0),
    "h_caii_r5": dict(kw=dict(KW, NRAYS=5, HYDROGEN_LTE="FALSE"), active=("H_6.atom",), wave=(630.25, 630.5, 21),
                      keys=("H ", "CA"), mu=0.8),
    "caii_r3_ff": dict(kw=dict(KW, NRAYS=3, HYDROGEN_LTE="TRUE", STOKES_MODE="FIELD_FREE"), active=(), wave=(854.2, 854.7, 41),
                       keys=("CA",), mu=1.0),
    "caii_r3_fs": dict(kw=dict(KW, NRAYS=3, HYDROGEN_LTE="TRUE", STOKES_MODE="FULL_STOKES"), active=(), wave=(854.2, 854.7, 41),
                       keys=("CA",), mu=0.9),
    "h_caii_r3_fs": dict(kw=dict(KW, NRAYS=3, HYDROGEN_LTE="FALSE", STOKES_MODE="FULL_STOKES"), active=("H_6.atom",),
                         wave=(630.25, 630.5, 21), keys=("H ", "CA"), mu=1.0),
    "caii_r3_prd": dict(kw=dict(KW, NRAYS=3, HYDROGEN_LTE="TRUE", PRD_N_MAX_ITER=3, PRD_ITER_LIMIT="1.0E-2"), active=(),
                        wave=(393.2, 393.5, 31), keys=("CA",), mu=1.0),
    "caii_r5_prd1": dict(kw=dict(KW, NRAYS=5, HYDROGEN_LTE="TRUE", PRD_N_MAX_ITER=1, PRD_ITER_LIMIT="1.0E-2"), active=(),
                         wave=(854.2, 854.7, 41), keys=("CA",), mu=0.8),
    "caii_r3_pf": dict(kw=dict(KW, NRAYS=3, HYDROGEN_LTE="TRUE", STOKES_MODE="POLARIZATION_FREE"), active=(), wave=(854.2, 854.7, 41),
                       keys=("CA",), mu=1.0),
    "h_caii_r5_ff": dict(kw=dict(KW, NRAYS=5, HYDROGEN_LTE="FALSE", STOKES_MODE="FIELD_FREE"), active=("H_6.atom",),
                         wave=(630.25, 630.5, 21), keys=("H ", "CA"), mu=0.8),
}
# columns of the synthetic batch; column 4 is left out: its MALI iteration does not converge in the reference (100
# iterations at NRAYS = 3) and ends in "Singular matrix" -> exit() at NRAYS = 5, so there is nothing to be identical to
COLUMNS = (0, 1, 2, 3, 5, 6, 7, 8)
NCOL = len(COLUMNS)


def workdir(case):
    c = CASES[case]
    return rd.make_workdir("tests", keywords=c["kw"], atoms_active=c["active"], atoms_extra=(("CaII.atom", "ACTIVE"),))


def one_column(case, col, path):
    """child process: one column of one case through the reference (it exit()s on some columns: "Singular matrix")"""
    c = CASES[case]
    base = np.load(GOLD / "falc_base.npy")
    atm = synthetic.perturbed_batch(base, max(COLUMNS) + 1, ndep=70)[list(COLUMNS)]
    o = rd.rhf1d(atm[col], np.linspace(*c["wave"]), workdir(case), mu=c["mu"], probe=rd.PROBE_NLTE, get_populations=True)
    R = recs_by_tag(o["records"])
    np.savez(path, nit=len({m[1] for m, _ in R["up_n"]}), I=o["I"], quv=np.array([o["Q"], o["U"], o["V"]]), lam=o["lam"],
             n=np.concatenate([o["pops"][k]["n"] for k in c["keys"]]), ns=np.concatenate([o["pops"][k]["nstar"] for k in c["keys"]]))


def main():
    import subprocess
    import tempfile
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        return one_column(sys.argv[2], int(sys.argv[3]), sys.argv[4])
    base = np.load(GOLD / "falc_base.npy")
    atm = synthetic.perturbed_batch(base, max(COLUMNS) + 1, ndep=70)[list(COLUMNS)]
    out = dict(atmosphere=atm)
    only = [a for a in sys.argv[1:] if a in CASES]
    if only:                                       # add / refresh the named cases, keep the rest of the fixture
        out = dict(np.load(GOLD / "nlte_front.npz"))
        assert np.array_equal(out["atmosphere"], atm)
    for case, c in CASES.items():
        if only and case not in only:
            continue
        wave = np.linspace(*c["wave"])
        I, n, ns, nit, quv, ok = [], [], [], [], [], []
        for col in range(NCOL):
            with tempfile.TemporaryDirectory() as td:
                f = str(Path(td) / "col.npz")
                r = subprocess.run([sys.executable, "-m", "oracle.gen_golden_nlte_front", "--one", case, str(col), f],
                                   cwd=str(ROOT), capture_output=True, text=True)
                if r.returncode != 0 or not Path(f).exists():
                    ok.append(0)
                    print(f"[golden] {case} column {col}: the reference exited ({r.stdout.strip()[-60:]!r})", flush=True)
                    continue
                o = dict(np.load(f))
            ok.append(1)
            nit.append(int(o["nit"])); I.append(o["I"]); quv.append(o["quv"]); n.append(o["n"]); ns.append(o["ns"]); lam = o["lam"]
            print(f"[golden] {case} column {col}: {nit[-1]} iterations, {len(lam)} wavelengths", flush=True)
        ok = np.array(ok, bool)

        def full(rows):                              # columns the reference aborted on are NaN / -1
            a = np.full((NCOL,) + rows[0].shape, np.nan)
            a[ok] = np.array(rows)
            return a
        nit_all = np.full(NCOL, -1, np.int32); nit_all[ok] = nit
        out.update({f"{case}_wave": wave, f"{case}_lam": lam, f"{case}_I": full(I), f"{case}_n": full(n),
                    f"{case}_nstar": full(ns), f"{case}_niter": nit_all, f"{case}_mu": np.float64(c["mu"])})
        if c["kw"]["STOKES_MODE"] != "NO_STOKES":
            out[f"{case}_QUV"] = full(quv)
    out["cases"] = np.array(json.dumps({k: dict(kw=v["kw"], active=list(v["active"]), keys=list(v["keys"])) for k, v in CASES.items()}))
    np.savez_compressed(GOLD / "nlte_front.npz", **out)
    print(f"[golden] nlte_front.npz: {(GOLD / 'nlte_front.npz').stat().st_size/1e6:.2f} MB")


if __name__ == "__main__":
    main()
