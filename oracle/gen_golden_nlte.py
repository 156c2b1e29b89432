"""oracle/gen_golden_nlte.py -- TEST INFRASTRUCTURE ONLY.

NLTE golden fixture: runs the UNMODIFIED reference on FAL-C with an ACTIVE model atom
(CaII.atom, CRD: PRD_N_MAX_ITER = 0; NRAYS = 3; Ng order 2; ITER_LIMIT 1e-4; NO_STOKES) and
records, through oracle/probe.c, the complete NLTE problem the reference is about to iterate
(state at the entry of Iterate(), rh/iterate.c:48) plus per-iteration Gamma / rates /
populations (updatePopulations, rh/statequil.c:177), a sample of SolveLinearEq calls and the
converged populations.  Flat layout = the one pyrh_b200.nlte.NlteProblem consumes.

    python -m oracle.gen_golden_nlte
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import refdriver as rd  # noqa: E402
from oracle.gen_golden import recs_by_tag, one  # noqa: E402

GOLD = ROOT / "tests" / "golden"
KW = {"NRAYS": 3, "N_MAX_SCATTER": 2, "N_MAX_ITER": 50, "NG_ORDER": 2, "NG_DELAY": 10, "NG_PERIOD": 3,
      "ITER_LIMIT": "1.0E-4", "PRD_N_MAX_ITER": 0, "STOKES_MODE": "NO_STOKES", "HYDROGEN_LTE": "TRUE"}

# transition row fields (doubles)
(TR_ATOM, TR_TYPE, TR_I, TR_J, TR_NBLUE, TR_NLAMBDA, TR_AJI, TR_BJI, TR_BIJ, TR_ISOFRAC, TR_WOFF,
 TR_PHIROW, TR_KR, TR_LINEIDX) = range(14)
TR_NFIELD = 16


def flatten(R):
    hdr = one(R, "nl_hdr")
    Ns, Nrays, Natom, N = int(hdr[0]), int(hdr[1]), int(hdr[2]), int(hdr[3])
    lam = one(R, "nl_lambda")
    out = dict(hdr=hdr, lam=lam, muz=one(R, "nl_muz"), wmu=one(R, "nl_wmu"), T=one(R, "nl_T"),
               height=one(R, "nl_height"), J0=one(R, "nl_J").reshape(Ns, N),
               bg=one(R, "nl_bg").reshape(3, Ns, N), bgflags=one(R, "nl_bgflags").reshape(Ns, 2).astype(np.int32))
    nlevel, n0, nstar, ntotal, C = [], [], [], [], []
    for m, d in R["nl_atom"]:
        nlevel.append(m[1])
    out["atom_nlevel"] = np.array(nlevel, np.int32)
    for tag, lst in (("nl_n", n0), ("nl_nstar", nstar), ("nl_ntotal", ntotal), ("nl_C", C)):
        for m, d in sorted(R[tag], key=lambda x: x[0][0]):
            lst.append(d.reshape(-1, N))
    out["n0"], out["nstar"] = np.concatenate(n0), np.concatenate(nstar)
    out["ntotal"], out["C"] = np.concatenate(ntotal), np.concatenate(C)
    # transitions: per atom lines first then continua (index used by the active sets)
    rows, wl, wlam, alpha, phis, wphi = [], [], [], [], [], []
    tr_index = {}
    phirow = 0
    for a in range(Natom):
        for m, d in sorted([x for x in R["nl_line"] if x[0][0] == a], key=lambda x: x[0][1]):
            kr, i, j, Nla, Nblue = m[1], m[2], m[3], m[4], m[5]
            r = np.zeros(TR_NFIELD)
            r[[TR_ATOM, TR_TYPE, TR_I, TR_J, TR_NBLUE, TR_NLAMBDA]] = [a, 0, i, j, Nblue, Nla]
            r[[TR_AJI, TR_BJI, TR_BIJ, TR_ISOFRAC]] = [d[1], d[2], d[3], d[4]]
            assert d[6] == 0, "PRD line in a CRD fixture"
            r[TR_WOFF], r[TR_PHIROW], r[TR_KR], r[TR_LINEIDX] = len(wl), phirow, kr, len(wphi)
            wl += list(d[8:8 + Nla]); wlam += list(d[8 + Nla:8 + 2 * Nla]); alpha += [0.0] * Nla
            wphi.append(d[8 + 2 * Nla:8 + 2 * Nla + N])
            ph = [x for x in R["nl_phi"] if x[0][0] == a and x[0][1] == kr][0]
            nrow = ph[0][2]
            assert nrow == 2 * Nrays * Nla, "static atmosphere fixture not supported"
            phis.append(ph[1].reshape(nrow, N)); phirow += nrow
            tr_index[(a, 0, kr)] = len(rows); rows.append(r)
        for m, d in sorted([x for x in R["nl_cont"] if x[0][0] == a], key=lambda x: x[0][1]):
            kr, i, j, Nla, Nblue = m[1], m[2], m[3], m[4], m[5]
            r = np.zeros(TR_NFIELD)
            r[[TR_ATOM, TR_TYPE, TR_I, TR_J, TR_NBLUE, TR_NLAMBDA]] = [a, 1, i, j, Nblue, Nla]
            r[TR_WOFF], r[TR_PHIROW], r[TR_KR], r[TR_LINEIDX] = len(wl), -1, kr, -1
            wl += list(d[4:4 + Nla]); alpha += list(d[4 + Nla:4 + 2 * Nla]); wlam += list(d[4 + 2 * Nla:4 + 3 * Nla])
            tr_index[(a, 1, kr)] = len(rows); rows.append(r)
    out["trans"] = np.array(rows)
    out["tr_lambda"], out["tr_wlambda"], out["tr_alpha"] = np.array(wl), np.array(wlam), np.array(alpha)
    out["phi"], out["wphi"] = np.concatenate(phis), np.array(wphi)
    # inputs of Profile(): damping parameter per line, Doppler width per atom, line-of-sight velocity
    ad, l0, comp_ok = [], [], True
    for a in range(Natom):
        for m, d in sorted([x for x in R["nl_adamp"] if x[0][0] == a], key=lambda x: x[0][1]):
            ad.append(d)
            comp_ok &= (m[3] == 1)
        for m, d in sorted([x for x in R["nl_line"] if x[0][0] == a], key=lambda x: x[0][1]):
            l0.append(d[0])
        for m, d in R["nl_comp"]:
            comp_ok &= (m[2] == 1 and d[0] == 0.0 and d[1] == 1.0)
    assert comp_ok, "multi-component lines are not covered by the fixture"
    out["adamp"], out["line_lambda0"] = np.array(ad), np.array(l0)
    out["vbroad"] = np.array([d for m, d in sorted(R["nl_vbroad"], key=lambda x: x[0][0])])
    out["vel"] = one(R, "nl_vel")
    first, lst = [0], []
    for m, d in sorted(R["nl_as"], key=lambda x: x[0][0]):
        cnt = m[1]
        for t in range(cnt):
            lst.append(tr_index[(int(d[3 * t]), int(d[3 * t + 1]), int(d[3 * t + 2]))])
        first.append(len(lst))
    out["as_first"], out["as_trans"] = np.array(first, np.int32), np.array(lst, np.int32)
    return out


def build(name, keywords=KW, atoms_active=(), atoms_extra=(("CaII.atom", "ACTIVE"),), pops_keys=("CA",), atm=None,
          wave=None, mu=1.0, drop=()):
    """Run the reference and write tests/golden/<name>.npz.  Several ACTIVE atoms are concatenated in the
    order of atmos.activeatoms (levels, Gamma blocks, transitions), which is the device layout."""
    if atm is None:
        atm = rd.falc("tests")
        atm[5] = 500.0
    if wave is None:
        wave = np.linspace(630.25, 630.5, 21)
    cwd = rd.make_workdir("tests", keywords=keywords, atoms_active=atoms_active, atoms_extra=atoms_extra)
    o = rd.rhf1d(atm, wave, cwd, mu=mu, probe=rd.PROBE_NLTE, get_populations=True)
    R = recs_by_tag(o["records"])
    g = flatten(R)
    N, Natom = int(g["hdr"][3]), int(g["hdr"][2])

    def per_iter(tag):
        """[iteration][atom-concatenated rows]: records carry (atom, niter) in their meta"""
        its = sorted({m[1] for m, _ in R[tag]})
        return its, [[d for m, d in sorted(R[tag], key=lambda x: x[0][0]) if m[1] == it] for it in its]

    its, gam = per_iter("up_gamma")
    g["niter"] = np.int32(len(its))
    keep = [0, 1, len(its) - 1]
    g["iter_keep"] = np.array(keep, np.int32)
    g["gamma_iter"] = np.array([np.concatenate([d.reshape(-1, N) for d in gam[i]]) for i in keep])
    _, rates = per_iter("up_rates")
    g["rates_iter"] = np.array([np.concatenate([d.reshape(-1, N) for d in rates[i]]) for i in keep])
    _, upn = per_iter("up_n")
    g["n_iter"] = np.array([np.concatenate([d[:-1].reshape(-1, N) for d in row]) for row in upn])   # all iterations
    g["dpops_iter"] = np.array([row[0][-1] for row in upn])
    g["n_final"] = np.concatenate([d.reshape(-1, N) for m, d in sorted(R["nl_n_final"], key=lambda x: x[0][0])])
    g["J_final"] = one(R, "nl_J_final").reshape(-1, N)
    # SolveLinearEq samples: the statEquil systems (N == Nlevel) of iteration 1 and the Ng systems
    nl = int(g["atom_nlevel"][0])
    lus = [(m, d) for m, d in R["lu"] if m[0] == nl][:N] + [(m, d) for m, d in R["lu"] if m[0] == 2][-6:]
    g["lu_n"] = np.array([m[0] for m, _ in lus], np.int32)
    g["lu_data"] = np.concatenate([d for _, d in lus])
    # final single-mu pass of _solveray(): recomputed background + profiles, emergent intensities
    Ns = len(g["lam"])
    g["fs_bg"] = one(R, "fs_bg").reshape(3, Ns, N)
    g["fs_bgflags"] = one(R, "fs_bgflags").reshape(Ns, 2).astype(np.int32)
    g["fs_muz"], g["fs_wmu"] = one(R, "fs_muz"), one(R, "fs_wmu")
    g["fs_phi"] = np.concatenate([d.reshape(-1, N) for m, d in sorted(R["fs_phi"], key=lambda x: (x[0][0], x[0][1]))])
    g["fs_wphi"] = np.array([d for m, d in sorted(R["fs_wphi"], key=lambda x: (x[0][0], x[0][1]))])
    g["fs_I"] = one(R, "fs_I")
    g["fs_adamp"] = np.array([d for m, d in sorted(R["fs_adamp"], key=lambda x: (x[0][0], x[0][1]))])
    g["fs_vbroad"] = np.array([d for m, d in sorted(R["fs_vbroad"], key=lambda x: x[0][0])])
    g["fs_vel"] = one(R, "fs_vel")
    g["spec_lam"], g["spec_I"] = o["lam"], o["I"]
    g["pops_final"] = np.concatenate([o["pops"][k]["n"] for k in pops_keys])
    g["atmosphere"], g["wave"], g["mu"] = np.array(atm), np.array(wave), np.float64(mu)
    for k in drop:
        g.pop(k, None)
    np.savez_compressed(GOLD / f"{name}.npz", **g)
    print(f"[golden] {name}: Natom={Natom} Nspect={len(g['lam'])} Ntrans={len(g['trans'])} iterations={int(g['niter'])} "
          f"dpops_last={g['dpops_iter'][-1]:.3e} -> {(GOLD / (name + '.npz')).stat().st_size/1e6:.2f} MB")
    return g


def main():
    if "--perturbed" in sys.argv:
        # one PERTURBED, moving 70-depth column (pyrh_b200.synthetic) with the user grid inside the ACTIVE
        # Ca II 8542 line and the final pass at mu = 0.8: every input of Iterate() and of _solveray()'s pass
        from pyrh_b200 import synthetic
        atm = synthetic.perturbed_batch(np.load(GOLD / "falc_base.npy"), 1, ndep=70)[0]
        build("nlte_caii_pert", atm=atm, wave=np.linspace(854.2, 854.7, 41), mu=0.8,
              drop=("phi", "gamma_iter", "rates_iter", "n_iter", "lu_n", "lu_data", "J0"))
        return
    if "--two-atom-only" not in sys.argv:
        build("nlte_caii")
    # BASELINE config 4: H (6 levels, 10 lines, 5 continua) + Ca II, both ACTIVE, CRD
    build("nlte_h_caii", keywords=dict(KW, HYDROGEN_LTE="FALSE"), atoms_active=("H_6.atom",), pops_keys=("H ", "CA"))


if __name__ == "__main__":
    main()
