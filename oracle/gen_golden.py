"""oracle/gen_golden.py -- TEST INFRASTRUCTURE ONLY.

Generates the committed golden fixtures under tests/golden/ by running the
UNMODIFIED reference (oracle/_ref/liboracle_scalar.so, built from
/root/reference by oracle/build_ref.sh) through its own entry point rhf1d()
on the reference's own input sets, and recording the boundary data of the
hot path with oracle/probe.c.  Runs only where /root/reference exists.

    python -m oracle.gen_golden            # writes tests/golden/*.npz

Each fixture holds, for one atmosphere column:
  inputs   atmosphere rows in pyrh units, wavelength grid
  column   SI arrays the reference holds when Formal() starts (T, ne, vturb,
           vel, B, cos_gamma, cos_2chi, sin_2chi, nHtot, np, height)
  tables   Kurucz line rows + Zeeman patterns + element/partition data
  chi_ai / eta_ai / sca_ai   angle-independent background, from a second
           reference run with an empty Kurucz file list (bit-exactly the `chi_ai`
           of background.c:343-465: verified chi_ai + rlk == total)
  elem_n   LTEpops_elem output
  rlk_*    rlk_opacity() outputs (up-ray) at a subset of wavelengths
  delo_*   Piece_Stokes_Bezier3_1D inputs/outputs at the same subset
  stokes_scalar / stokes_simd   rhf1d() result of both reference builds
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import refdriver as rd  # noqa: E402
from pyrh_b200 import linelist as ll  # noqa: E402

GOLD = ROOT / "tests" / "golden"


def recs_by_tag(records):
    out = {}
    for tag, meta, data in records:
        out.setdefault(tag, []).append((meta, data))
    return out


def tables_from_records(R, vmicro_char):
    """RLK_Line / Element dumps (probe.c:snapshot) -> LineTable."""
    lines_raw = sorted(R["rlk_line"], key=lambda md: md[0][0])
    elem_ids, elems, pfrows, elem_rows = [], [], [], {}
    for meta, d in R["elem"]:
        pt = meta[0]
        if pt in elem_rows:
            continue
        nst = meta[1]
        row = np.zeros(ll.RE_NFIELD)
        row[ll.RE_WEIGHT], row[ll.RE_ABUND], row[ll.RE_NSTAGE] = d[0], d[1], nst
        row[ll.RE_PFROW] = len(pfrows)
        assert nst <= ll.RE_MAXSTAGE
        row[ll.RE_IONPOT0: ll.RE_IONPOT0 + nst] = d[8:8 + nst]
        pf = {m[1]: dd for m, dd in R["elem_pf"] if m[0] == pt}
        for i in range(nst):
            pfrows.append(pf[i])
        elem_rows[pt] = len(elems)
        elems.append(row)
        elem_ids.append(pt)
    rows, zq, zs, zst = [], [], [], []
    for meta, d in lines_raw:
        r = np.zeros(ll.RL_NFIELD)
        (r[ll.RL_LAMBDA0], r[ll.RL_GI], r[ll.RL_GJ], r[ll.RL_EI], r[ll.RL_EJ], r[ll.RL_BJI],
         r[ll.RL_AJI], r[ll.RL_BIJ]) = d[0:8]
        r[ll.RL_GRAD], r[ll.RL_GSTARK], r[ll.RL_GVDW] = d[10], d[11], d[12]
        r[ll.RL_HFS_FRAC], r[ll.RL_ISO_FRAC] = d[13], d[14]
        r[ll.RL_CROSS], r[ll.RL_ALPHA] = d[17], d[18]
        r[ll.RL_POLARIZABLE], r[ll.RL_VDWAALS] = d[19], d[20]
        r[ll.RL_ELEM], r[ll.RL_STAGE] = elem_rows[int(d[21])], d[22]
        nc = int(d[25])
        r[ll.RL_ZOFF], r[ll.RL_NCOMP] = len(zq), nc
        comp = d[32:32 + 3 * nc].reshape(nc, 3)
        zq += [int(x) for x in comp[:, 0]]
        zs += list(comp[:, 1])
        zst += list(comp[:, 2])
        rows.append(r)
    lt = ll.LineTable(lines=np.array(rows), zq=np.array(zq, np.int32), zshift=np.array(zs),
                      zstrength=np.array(zst), elems=np.array(elems), pf=np.array(pfrows),
                      Tpf=R["Tpf"][0][1], vmicro_char=vmicro_char)
    lt.validate()
    return lt, elem_ids


def one(R, tag):
    return R[tag][0][1]


def make_case(name, atm, wave, kind="benchmark", subset_step=10, keywords=None):
    cwd = rd.make_workdir(kind, keywords=keywords)
    full = rd.rhf1d(atm, wave, cwd, variant="scalar", probe=rd.PROBE_ALL)
    simd = rd.rhf1d(atm, wave, cwd, variant="simd")
    cwd0 = rd.make_workdir(kind, keywords=keywords, no_kurucz=True)
    cont = rd.rhf1d(atm, wave, cwd0, variant="scalar", probe=rd.PROBE_BG | rd.PROBE_SNAP)

    R, R0 = recs_by_tag(full["records"]), recs_by_tag(cont["records"])
    flags = one(R, "flags")
    lam = one(R, "lambda")
    ns, nd = len(lam), atm.shape[1]
    assert np.array_equal(lam, one(R0, "lambda")), "wavelength sets differ between runs"
    for f in ("height", "np", "T", "ne", "nHtot"):
        assert np.array_equal(one(R, f), one(R0, f)), f

    # angle-independent background from the line-free run
    chi_ai, eta_ai, sca_ai = np.zeros((ns, nd)), np.zeros((ns, nd)), np.zeros((ns, nd))
    for meta, d in R0["bg"]:
        n, nst = meta[0], meta[3]
        assert nst == 1
        chi_ai[n], eta_ai[n], sca_ai[n] = d[:nd], d[nd:2 * nd], d[2 * nd:3 * nd]
    # totals (with lines), last write wins (readj.c:319-345): mu = Nrays-1, to_obs = 1
    tot = {}
    for meta, d in R["bg"]:
        tot[meta[0]] = (meta[3], d)
    rlk_up = {m[0]: (m, d) for m, d in R["rlk"] if m[2] == 1}
    nbad = 0
    for n in range(ns):
        nst, d = tot[n]
        m, r = rlk_up[n]
        if m[3]:
            chi_l, eta_l = r[:4 * nd].reshape(4, nd), r[4 * nd:].reshape(4, nd)
            ok = (np.array_equal(chi_ai[n] + chi_l[0], d[:nd]) and
                  np.array_equal(eta_ai[n] + eta_l[0], d[nst * nd:(nst + 1) * nd]))
        else:
            ok = np.array_equal(chi_ai[n], d[:nd])
        nbad += (not ok)
    assert nbad == 0, f"{nbad} wavelengths where chi_ai + rlk != total (passive_bb/molecular lines?)"

    lt, elem_ids = tables_from_records(R, vmicro_char=float(flags[6]))
    elem_n = np.zeros((lt.nelem, ll.RE_MAXSTAGE, nd))
    for m, d in R["elem_n"]:
        elem_n[elem_ids.index(m[0]), m[1]] = d

    # wavelengths that belong to the user grid (lambda_ref removed, pyrh_solveray.c:130-150)
    keep = lam != flags[7]
    assert np.array_equal(lam[keep], full["lam"])
    sub = np.arange(0, ns, subset_step)
    delo_up = {m[0]: d for m, d in R["delo"] if m[2] == 1}
    sub = np.array([n for n in sub if n in delo_up])
    out = dict(
        atmosphere=atm, wave=wave, lam_spect=lam, lam_keep=keep, flags=flags,
        chi_ai=chi_ai, eta_ai=eta_ai, sca_ai=sca_ai, elem_n=elem_n,
        backgrflags=one(R, "backgrflags").reshape(ns, 2).astype(np.int32),
        sub=sub,
        rlk_chi=np.array([rlk_up[n][1][:4 * nd].reshape(4, nd) for n in sub]),
        rlk_eta=np.array([rlk_up[n][1][4 * nd:].reshape(4, nd) for n in sub]),
        delo=np.array([delo_up[n].reshape(13, nd) for n in sub]),
        stokes_scalar=np.array([full[k] for k in "IQUV"]),
        stokes_simd=np.array([simd[k] for k in "IQUV"]),
        muz=one(R, "muz"),
    )
    for f in ("T", "ne", "vturb", "vel", "B", "cos_gamma", "cos_2chi", "sin_2chi",
              "nHtot", "np", "height", "tau_ref"):
        out["col_" + f] = one(R, f)
    out.update(lt.to_npz_dict())
    # all up-ray DELO records, kept only for the in-container port validation below
    extra = dict(delo_all=delo_up, rlk_all=rlk_up)
    GOLD.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(GOLD / f"{name}.npz", **out)
    print(f"[golden] {name}: Nspect={ns} Ndep={nd} lines={lt.nline} "
          f"-> {(GOLD / (name + '.npz')).stat().st_size/1e6:.2f} MB")
    return out, extra


def validate_port(name, g, extra):
    """Pin the C restatement against every record of the reference run."""
    from oracle import portdriver as pd
    lt = ll.LineTable.from_npz(g)
    tab = pd.PortTables(lt)
    col = pd.PortColumn(muz=float(g["muz"][0]), moving=bool(g["flags"][0]),
                        **{k: g["col_" + k] for k in pd.PortColumn.FIELDS})
    nd = col.Ndep
    en = pd.elem_pops(tab, col)
    d_n = np.abs(en - g["elem_n"]).max() / np.abs(g["elem_n"]).max()
    lam = g["lam_spect"]
    worst = dict(rlk=0.0, delo=0.0)
    nexact = dict(rlk=0, delo=0)
    for n, (m, r) in extra["rlk_all"].items():
        fl, chi, eta = pd.rlk_opacity(tab, col, g["elem_n"], lam[n], 1)
        assert (fl & 1) == m[3] and ((fl >> 1) & 1) == m[4], (n, fl, m)
        if not m[3]:
            continue
        ref = r.reshape(2, 4, nd)
        got = np.array([chi, eta])
        nexact["rlk"] += np.array_equal(ref, got)
        scale = np.abs(ref[:, 0]).max(axis=1)[:, None, None]
        worst["rlk"] = max(worst["rlk"], (np.abs(got - ref) / scale).max())
    for n, d in extra["delo_all"].items():
        d = d.reshape(13, nd)
        I = pd.stokes_bezier3(g["col_height"], col.c.muz, 1, d[0], d[1:5], d[10:13],
                              g["col_T"], lam[n])
        nexact["delo"] += np.array_equal(I, d[5:9])
        worst["delo"] = max(worst["delo"], np.abs(I - d[5:9]).max() / np.abs(d[5]).max())
    keep = g["lam_keep"]
    st = pd.lte_stokes_column(tab, col, lam[keep], g["chi_ai"][keep], g["eta_ai"][keep])
    ref = g["stokes_scalar"]
    Ic = ref[0].max()
    print(f"[port-vs-ref] {name}: LTEpops rel {d_n:.2e}; rlk exact {nexact['rlk']}/{len(extra['rlk_all'])} "
          f"worst {worst['rlk']:.2e}; delo exact {nexact['delo']}/{len(extra['delo_all'])} "
          f"worst {worst['delo']:.2e}; spectrum: I rel {np.abs(st[0]/ref[0]-1).max():.2e} "
          f"QUV/Ic {np.abs(st[1:]-ref[1:]).max()/Ic:.2e} exact={np.array_equal(st, ref)}")


def make_scalar_case(name, atm, wave):
    """NO_STOKES LTE run on a grid wider than the line windows: the reference then solves
    line-window wavelengths with Piecewise_Bezier3_1D (formal.c:234-235) and line-free ones with
    Feautrier (formal.c:289-309).  Records every such call (inputs + outputs)."""
    cwd = rd.make_workdir("benchmark", keywords={"STOKES_MODE": "NO_STOKES"})
    full = rd.rhf1d(atm, wave, cwd, variant="scalar", probe=rd.PROBE_ALL)
    R = recs_by_tag(full["records"])
    lam, flags = one(R, "lambda"), one(R, "flags")
    nd = atm.shape[1]
    bez = [(m, d.reshape(4, nd)) for m, d in R.get("bez", [])]
    feau = [(m, d) for m, d in R.get("feau", [])]
    out = dict(atmosphere=atm, wave=wave, lam_spect=lam, flags=flags, muz=one(R, "muz"),
               backgrflags=one(R, "backgrflags").reshape(len(lam), 2).astype(np.int32),
               bez_meta=np.array([m[:4] for m, _ in bez], np.int32), bez=np.array([d for _, d in bez]),
               feau_meta=np.array([m[:4] for m, _ in feau], np.int32),
               feau=np.array([d[:4 * nd].reshape(4, nd) for _, d in feau]),
               feau_Iem=np.array([d[4 * nd] for _, d in feau]),
               I_scalar=full["I"], lam_out=full["lam"])
    for f in ("T", "height"):
        out["col_" + f] = one(R, f)
    np.savez_compressed(GOLD / f"{name}.npz", **out)
    print(f"[golden] {name}: {len(bez)} Bezier3 rays, {len(feau)} Feautrier rays "
          f"-> {(GOLD / (name + '.npz')).stat().st_size/1e6:.2f} MB")
    return out


def validate_scalar_port(name, g):
    from oracle import portdriver as pd
    nb = nf = 0
    for m, d in zip(g["bez_meta"], g["bez"]):
        I, Psi = pd.bezier3_scalar(g["col_height"], float(g["muz"][m[1]]), int(m[2]), d[0], d[1], g["col_T"],
                                   g["lam_spect"][m[0]], want_psi=True)
        nb += np.array_equal(I, d[2]) and (not m[3] or np.array_equal(Psi, d[3]))
    for m, d, Iem in zip(g["feau_meta"], g["feau"], g["feau_Iem"]):
        P, Psi, I0 = pd.feautrier(g["col_height"], float(g["muz"][m[1]]), d[0], d[1], g["col_T"],
                                  g["lam_spect"][m[0]])
        nf += np.array_equal(P, d[2]) and I0 == Iem and (not m[3] or np.array_equal(Psi, d[3]))
    print(f"[port-vs-ref] {name}: Bezier3 exact {nb}/{len(g['bez'])}, Feautrier exact {nf}/{len(g['feau'])}")


def make_voigt_armstrong():
    """Voigt(a, v, NULL, ARMSTRONG) of the reference itself (rh/voigt.c:83,126-243) on random (a, v)
    covering the K1 / K2 / K3 branches."""
    import ctypes as C
    from oracle import portdriver as pd
    ref = rd.load("scalar")
    ref.Voigt.restype = C.c_double
    ref.Voigt.argtypes = [C.c_double, C.c_double, C.POINTER(C.c_double), C.c_int]
    port = pd.lib()
    port.rp_voigt_armstrong.restype = C.c_double
    port.rp_voigt_armstrong.argtypes = [C.c_double, C.c_double]
    port.rp_armstrong_region.argtypes = [C.c_double, C.c_double]
    rng = np.random.default_rng(5)
    n = 20000
    a = 10 ** rng.uniform(-5, 1.2, n)
    v = rng.uniform(-12, 12, n)
    v[:3000] = rng.uniform(-300, 300, 3000)
    H = np.array([ref.Voigt(a[i], v[i], None, 0) for i in range(n)])          # enum ARMSTRONG = 0 (rh.h)
    P = np.array([port.rp_voigt_armstrong(a[i], v[i]) for i in range(n)])
    reg = np.array([port.rp_armstrong_region(a[i], v[i]) for i in range(n)], np.int8)
    np.savez_compressed(GOLD / "voigt_armstrong.npz", a=a, v=v, H=H, region=reg)
    print(f"[golden] voigt_armstrong: {n} points, regions {np.bincount(reg)[1:]}; port exact = {np.array_equal(H, P)}")


def falc_case_atm():
    atm = rd.falc("benchmark")
    atm[5] = 1000.0                                   # B [G]; gamma, chi from falc.dat (pi/4, pi/3)
    atm[3] = 0.5 * np.sin(np.linspace(0, 3, atm.shape[1]))   # v_z [km/s]
    return atm


def main():
    from pyrh_b200 import synthetic
    wave = rd.hinode_wave()
    g, ex = make_case("falc_B1kG", falc_case_atm(), wave)
    validate_port("falc_B1kG", g, ex)
    make_voigt_armstrong()
    gs = make_scalar_case("falc_scalar", falc_case_atm(), rd.air_to_vacuum(np.linspace(629.7, 630.7, 101)))
    validate_scalar_port("falc_scalar", gs)
    # FULL_STOKES on the same wide grid: DELO inside the line windows, Feautrier outside
    g, ex = make_case("falc_B1kG_wide", falc_case_atm(), rd.air_to_vacuum(np.linspace(629.7, 630.7, 101)),
                      subset_step=25)
    validate_port("falc_B1kG_wide", g, ex)
    if "--scalar-only" in sys.argv:
        return
    # BASELINE config 2 columns: perturbed FAL-C resampled to 70 depths
    base = rd.falc("tests")
    np.save(GOLD / "falc_base.npy", base)
    for c in range(3):
        atm = synthetic.perturbed_batch(base, 1, first=c)[0]
        g, ex = make_case(f"synth70_c{c}", atm, wave, subset_step=60)
        validate_port(f"synth70_c{c}", g, ex)


if __name__ == "__main__":
    main()
