"""pyrh_b200.host: the text inputs of a pyrh working directory parsed in Python (readInput / readAbundance /
readKuruczLines of the reference) must give, bit for bit, the tables recorded from the reference's own parsed
state (fixtures synth70_c0, pyrh_scales, synth70_chem).  Needs the reference's data files staged under
oracle/_ref (make -C oracle ref); they travel to the GPU box with the snapshot."""
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLD

ROOT = Path(__file__).resolve().parent.parent
CWD = ROOT / "oracle" / "_ref" / "inputs" / "benchmark"
PYRH_PATH = ROOT / "oracle" / "_ref" / "pyrh_path"

pytestmark = pytest.mark.skipif(not (CWD / "keyword.input").exists() or not (PYRH_PATH / "rh" / "Atoms").exists(),
                                reason="reference input files not staged (oracle/_ref)")


def test_keywords_and_elements():
    from pyrh_b200 import host
    kw = host.read_keywords(CWD)
    assert kw["KURUCZ_DATA"] == "kurucz.input" and kw["STOKES_MODE"] == "FULL_STOKES" and float(kw["VMICRO_CHAR"]) == 5.0
    el = host.read_elements(PYRH_PATH, kw)
    sums = np.load(GOLD / "pyrh_scales.npz")["tau_abund_sums"]          # abundance.c:219-221 of the reference run
    assert el.wght_per_H == sums[0] and el.totalAbund == sums[1] and el.avgMolWght == sums[2]
    bg = np.load(ROOT / "pyrh_b200" / "data" / "background_falc11.npz")
    ab = np.array([el.abund[int(p) - 1] for p in bg["atom_pt_index"]])
    assert np.array_equal(ab, np.load(GOLD / "synth70_chem.npz")["abundance"])
    assert len(el.ID) == 99 and el.ID[25] == "FE" and el.nstage[25] == 6


def test_kurucz_lines_equal_reference_tables():
    from pyrh_b200 import host, linelist as ll
    kw = host.read_keywords(CWD)
    el = host.read_elements(PYRH_PATH, kw)
    lt = host.read_kurucz_lines(CWD, kw, el)
    ref = ll.LineTable.from_npz(np.load(GOLD / "synth70_c0.npz"))
    cols = [i for i in range(ll.RL_NFIELD) if i != ll.RL_ALPHA]      # alpha is uninitialised memory for Unsold lines
    assert np.array_equal(lt.lines[:, cols], ref.lines[:, cols])
    assert np.array_equal(lt.zq, ref.zq) and np.array_equal(lt.zshift, ref.zshift)
    assert np.array_equal(lt.zstrength, ref.zstrength)
    assert np.array_equal(lt.pf, ref.pf) and np.array_equal(lt.Tpf, ref.Tpf) and lt.vmicro_char == ref.vmicro_char
    n = int(ref.elems[0, ll.RE_NSTAGE])
    assert np.array_equal(lt.elems[:, :ll.RE_IONPOT0 + n - 1], ref.elems[:, :ll.RE_IONPOT0 + n - 1])


def test_loggf_and_wavelength_overrides():
    """loggf_ids / lam_ids of pyrh.compute1d (kurucz.c:224-234, 247-257) address lines by their position in the list."""
    from pyrh_b200 import host, linelist as ll
    kw = host.read_keywords(CWD)
    el = host.read_elements(PYRH_PATH, kw)
    a = host.read_kurucz_lines(CWD, kw, el)
    b = host.read_kurucz_lines(CWD, kw, el, loggf_ids=[1], loggf_values=[-0.5], lam_ids=[0], lam_values=[0.001])
    assert b.lines[0, ll.RL_LAMBDA0] > a.lines[0, ll.RL_LAMBDA0] and b.lines[1, ll.RL_LAMBDA0] == a.lines[1, ll.RL_LAMBDA0]
    assert b.lines[1, ll.RL_AJI] / a.lines[1, ll.RL_AJI] == pytest.approx(10 ** (-0.5 + 0.968), rel=1e-12)


def test_kurucz_table_remembers_file_positions(tmp_path):
    """get_atomic_rfs addresses lines by their position in the Kurucz files while the device table is sorted by
    wavelength: LineTable.file_index maps one onto the other (RLK_Line.loggf_rf_ind, kurucz.c:250-259)."""
    import shutil
    from pyrh_b200 import host, linelist as ll
    for f in Path(CWD).iterdir():
        if f.is_file() and f.suffix not in (".fits", ".spec", ".py"):
            shutil.copy(f, tmp_path / f.name)
    (tmp_path / "kurucz.input").write_text("lines_4016\n")
    kw = host.read_keywords(tmp_path)
    el = host.read_elements(PYRH_PATH, kw)
    lt = host.read_kurucz_lines(tmp_path, kw, el)
    assert sorted(lt.file_index) == list(range(len(lt.lines)))
    lam_file = [float(r.split()[0]) for r in host.read_kurucz_records(tmp_path, kw["KURUCZ_DATA"])]
    assert np.allclose(lt.lines[:, ll.RL_LAMBDA0], [host.air_to_vacuum(lam_file[i]) for i in lt.file_index], rtol=1e-12)


def test_abundance_override_and_sort_lambda():
    from pyrh_b200 import host
    kw = host.read_keywords(CWD)
    a = host.read_elements(PYRH_PATH, kw)
    b = host.read_elements(PYRH_PATH, kw, atomic_number=[26], atomic_abundance=[7.60])
    assert b.abund[25] == host.POW10(7.60 - 12.0) and b.abund[25] != a.abund[25] and b.abund[24] == a.abund[24]
    lam = host.sort_lambda([630.2, 630.1, 630.2], 500.0)
    assert list(lam) == [500.0, 630.1, 630.2]


def test_unsupported_inputs_are_refused(tmp_path):
    import shutil
    from pyrh_b200 import host
    for f in CWD.iterdir():
        if f.is_file() and f.stat().st_size < 1_000_000:
            shutil.copy(f, tmp_path / f.name)
    kwf = tmp_path / "keyword.input"
    kwf.write_text(kwf.read_text().replace("MAGNETO_OPTICAL = FALSE", "MAGNETO_OPTICAL = TRUE"))
    with pytest.raises(NotImplementedError, match="MAGNETO_OPTICAL"):
        host.Session(tmp_path, [630.1, 630.2], path=PYRH_PATH)


def test_molecular_line_list_equals_reference_tables():
    """readMolecularLines in Python (KURUCZ_NEW format): the 99 CN B-X lines the reference ships equal, bit for bit,
    the line table recorded from the reference's parsed state (fixture falc_molecules)."""
    from pyrh_b200 import host
    kw = host.read_keywords(CWD)
    el = host.read_elements(PYRH_PATH, kw)
    rows, sel, zee = host.molecular_line_table(CWD, kw, el, PYRH_PATH)
    assert len(zee[0]) == 0 and not rows[:, host.ML_POLARIZABLE].any()      # the shipped CN list carries no Hund's-case data
    g = np.load(GOLD / "falc_molecules.npz")
    assert rows.shape == (99, host.ML_NFIELD) and np.array_equal(rows[:, :9], g["mlines"][:, :9])
    assert sel.shape == (1, host.MS_NFIELD) and sel[0, 0] == 7 and sel[0, 1] == 12.01 + 14.01       # CN: 8th molecule


def test_polarizable_molecular_lines_get_molzeeman_patterns(tmp_path):
    """Hund's-case columns behind column 71 of a molecular line list (readmolecule.c:859-912) make the lines polarizable;
    MolZeeman (molzeeman.c:196-319: case-b Lande factors, anomalous pattern, strengths normalised per q) in Python equals
    the reference's patterns bit for bit -- 99 lines, 9080 components recorded from the reference on a polarizable copy
    of the shipped CN list (oracle.refdriver.polarizable_cn_tree, fixture falc_molecules_pol).  A list without the
    subbranch digit where its format expects it is refused (the reference reads uninitialised memory there)."""
    from pyrh_b200 import host
    from oracle import refdriver as rd
    if not rd.available():
        pytest.skip("reference data files not staged (oracle/_ref)")
    pp = rd.polarizable_cn_tree(str(tmp_path / "pyrh_path"))
    kw = host.read_keywords(CWD)
    el = host.read_elements(pp, kw)
    rows, sel, (zq, zs, zt) = host.molecular_line_table(CWD, kw, el, pp)
    g = np.load(GOLD / "falc_molecules_pol.npz")
    assert rows.shape == (99, host.ML_NFIELD) and rows[:, host.ML_POLARIZABLE].all()
    assert np.array_equal(rows[:, :9], g["mlines"][:, :9])
    assert np.array_equal(rows[:, host.ML_ZOFF], g["mlines"][:, 9]) and np.array_equal(rows[:, host.ML_NCOMP], g["mlines"][:, 10])
    assert np.array_equal(zq, g["zq"]) and np.array_equal(zs, g["zshift"]) and np.array_equal(zt, g["zstrength"])
    for q in (-1, 0, 1):                                          # normalised per q, line by line
        o, n = int(rows[0, host.ML_ZOFF]), int(rows[0, host.ML_NCOMP])
        assert abs(zt[o:o + n][zq[o:o + n] == q].sum() - 1.0) < 1e-14
    lst = Path(pp) / "rh" / "Molecules" / "CN" / "CN_B-X_polarizable.asc"
    lst.write_text(lst.read_text().replace("KURUCZ_CD18", "KURUCZ_NEW"))
    with pytest.raises(ValueError, match="subbranch"):
        host.molecular_line_table(CWD, kw, el, pp)


def test_background_model_from_atom_and_molecule_files():
    """read_background_model (readAtom / readMolecule in Python) reproduces, bit for bit, the model recorded from the
    reference's parsed state: 221 levels, 157 bound-free edges with 6368 table points, the Rayleigh lines, the chemical
    network of 4 nuclei and 12 molecules.  (The reference leaves the alpha table of HYDROGENIC edges uninitialised.)"""
    from pyrh_b200 import host
    kw = host.read_keywords(CWD)
    el = host.read_elements(PYRH_PATH, kw)
    m = host.read_background_model(CWD, kw, el, PYRH_PATH)
    ref = dict(np.load(ROOT / "pyrh_b200" / "data" / "background_falc11.npz"))
    for k in ("ct_hdr", "ct_lev", "ct_bf", "ct_tab_lambda", "ct_ray", "ce_nuclei", "ce_mol", "atom_pt_index"):
        assert np.array_equal(np.asarray(m[k]), ref[k]), k
    explicit = np.zeros(len(ref["ct_tab_alpha"]), bool)
    for b in ref["ct_bf"]:
        if b[5] == 0.0:
            explicit[int(b[8]):int(b[8]) + int(b[7])] = True
    assert explicit.sum() > 4000
    assert np.array_equal(m["ct_tab_alpha"][explicit], ref["ct_tab_alpha"][explicit])


def test_passive_line_table_shapes():
    from pyrh_b200 import host
    kw = host.read_keywords(CWD)
    el = host.read_elements(PYRH_PATH, kw)
    bg = host.read_background_model(CWD, kw, el, PYRH_PATH)
    first = [int(np.flatnonzero(bg["ct_lev"][:, 0] == a)[0]) for a in range(11)]
    rows, cs, cf = host.passive_line_table(CWD, kw, el, first, PYRH_PATH)
    assert rows.shape == (202, host.PL_NFIELD) and len(cs) == len(cf) == 202
    ha = rows[np.argmin(np.abs(rows[:, host.PL_LAMBDA0] - 656.47))]
    assert ha[host.PL_IS_H] == 1 and ha[host.PL_LINSTARK_C] > 0 and ha[host.PL_VDW_TYPE] == host.VDW_UNSOLD_A


def test_barklem_tables_and_cubic_convolution():
    """tests/fe6300 of the reference gives the orbital quantum numbers of the Fe I pair, so readKuruczLines takes the
    Anstee-Barklem-O'Mara s-p table (getBarklemcross, barklem.c:139-196, with cubeconvol.c): cross-section and
    velocity exponent equal the reference's parsed values bit for bit."""
    from pyrh_b200 import host, linelist as ll
    cwd = ROOT / "oracle" / "_ref" / "inputs" / "tests"
    if not (cwd / "keyword.input").exists():
        pytest.skip("reference tests/ inputs not staged")
    g = np.load(GOLD / "ref_test_compute1d.npz")
    kw = host.read_keywords(cwd)
    el = host.read_elements(PYRH_PATH, kw)
    lt = host.read_kurucz_lines(cwd, kw, el, path=PYRH_PATH)
    assert np.array_equal(lt.lines[:, ll.RL_VDWAALS], g["rlk_vdwaals"]) and lt.lines[0, ll.RL_VDWAALS] == ll.VDW_BARKLEM
    assert np.array_equal(lt.lines[:, ll.RL_CROSS], g["rlk_cross"])
    assert np.array_equal(lt.lines[:, ll.RL_ALPHA], g["rlk_alpha"])


def test_barklem_cross_section_of_model_atom_lines():
    """getBarklemactivecross (barklem.c:216-312) on the host: the Mg b lines of MgI_6level.atom (3s3p 3P - 3s4s 3S)
    and Ca I 422.7 nm take the s-p table; the intercombination line keeps UNSOLD; lines of ions fall back to UNSOLD."""
    from pyrh_b200 import host
    kw = host.read_keywords(CWD)
    el = host.read_elements(PYRH_PATH, kw)
    atoms = PYRH_PATH / "rh" / "Atoms"
    mg = host.read_atom(atoms / "MgI_6level.atom")
    assert mg["abo_level"] == [-1] * 6
    w = el.weight[el.ID.index("MG")]
    hits = [host.barklem_active_cross(mg, ln, w, el.weight[0], PYRH_PATH) for ln in mg["lines"] if "BARKLEM" in ln["vdw"]]
    assert len(hits) == 3 and all(h is not None for h in hits)
    for cross, alpha in hits:
        assert 0.2 < alpha < 0.35 and 1e-15 < cross < 1e-13          # ABO: sigma ~ 300-700 a0^2, alpha ~ 0.25-0.3
    ca2 = host.read_atom(atoms / "CaII.atom")
    assert all("BARKLEM" not in ln["vdw"] for ln in ca2["lines"])      # ions: readatom.c:313-319 -> UNSOLD


@pytest.mark.parametrize("fixture,kw,active", [("nlte_caii", {}, ()), ("nlte_h_caii", {"HYDROGEN_LTE": "FALSE"}, ("H_6.atom",))])
def test_nlte_plan_equals_reference_parsed_state(fixture, kw, active):
    """pyrh_b200.nlte_host: readAtom of ACTIVE atoms + getLambda + SortLambda (merged grid, Nblue / Nlambda with Hunt's
    starting-guess behaviour, active sets in the reference's order, hydrogenic and tabulated bound-free cross-sections,
    quadrature weights) give the flat plan the probe recorded inside the reference, bit for bit."""
    import os
    from oracle import refdriver as rd
    from oracle.gen_golden_nlte import KW
    from pyrh_b200 import host, nlte_host as nh
    os.environ["PYRH_PATH"] = str(PYRH_PATH)
    g = np.load(GOLD / f"{fixture}.npz")
    cwd = rd.make_workdir("tests", keywords=dict(KW, **kw), atoms_active=active, atoms_extra=(("CaII.atom", "ACTIVE"),))
    k = host.read_keywords(cwd)
    root = PYRH_PATH / "rh" / "Atoms"
    atoms = [nh.read_active_atom(root / f, k) for f, s in host._atoms_listed(cwd, k) if s == "ACTIVE"]
    p = nh.build_plan(atoms, np.linspace(630.25, 630.5, 21), float(k["LAMBDA_REF"]), int(k["NRAYS"]), k)
    for key in ("lam", "muz", "wmu", "atom_nlevel", "tr_lambda", "tr_wlambda", "tr_alpha", "as_first", "as_trans"):
        assert np.array_equal(p[key], g[key]), key
    gt = g["trans"].copy()
    lines = gt[:, nh.TR_TYPE] == 0
    gt[lines, nh.TR_LAMBDA0] = g["line_lambda0"][gt[lines, nh.TR_LINEIDX].astype(int)]
    assert np.array_equal(p["trans"], gt)
    assert p["nphirow"] == g["phi"].shape[0] and p["nline"] == len(g["wphi"])
    p1 = nh.single_mu_plan(p, 1.0)
    assert p1["nphirow"] == g["fs_phi"].shape[0] and list(p1["muz"]) == [1.0]
    coll, T, C, M = nh.collision_table(atoms)
    assert len(coll) == sum(1 for at in atoms for ln in at["coll"] if ln.split()[0] in ("OMEGA", "CE", "CI"))
    assert len(T) == len(C) == len(M) == int(coll[:, nh.CO_NT].sum())


def test_hunt_reproduces_the_reference_quirk():
    """Hunt() (hunt.c:17-78) is not Locate(): hunting DOWN onto an exact table value returns the index below it.
    SortLambda's Nred inherits this, so nlte_host.hunt must too (H 2-4 of H_6.atom loses its last wavelength that way)."""
    from pyrh_b200 import nlte_host as nh
    arr = [float(x) for x in range(10)]
    assert nh.locate(arr, 4.0) == 4 and nh.hunt(arr, 4.0, 0) == 4           # no usable guess: bisection
    assert nh.hunt(arr, 4.0, 2) == 4                                         # hunting up: exact hit kept
    assert nh.hunt(arr, 4.0, 6) == 3                                         # hunting down lands ON the value: one below
    assert nh.hunt(arr, 4.0, 7) == 4                                         # ... steps over it: bracket still contains it
    assert nh.hunt(arr, 4.5, 7) == 4 and nh.hunt(arr, 9.0, 9) == 9 and nh.hunt(arr, 0.0, 5) == 0
    rng = np.random.default_rng(3)
    tab = np.sort(rng.uniform(0, 100, 200))
    for v, guess in zip(rng.uniform(-5, 105, 500), rng.integers(0, 200, 500)):
        lo = nh.locate(tab, v)
        assert nh.hunt(tab, v, int(guess)) == lo                              # generic values: same bracket


def test_gauss_legendre_and_spline_helpers():
    from pyrh_b200 import nlte_host as nh
    x, w = nh.gauss_leg(0.0, 1.0, 5)
    xs, ws = np.polynomial.legendre.leggauss(5)
    assert np.allclose(x, 0.5 * (xs + 1), atol=1e-14) and np.allclose(w, 0.5 * ws, atol=1e-14)
    t = [3000.0, 5000.0, 7000.0, 15000.0, 50000.0]
    y = [1.0, 2.0, 1.5, 4.0, 3.0]
    M = nh.spline_coef(t, y)
    assert M[0] == 0.0 and M[-1] == 0.0
    assert nh.spline_eval(t, y, M, [5000.0, 2000.0, 60000.0]) == [2.0, 1.0, 3.0]
    from scipy.interpolate import CubicSpline
    cs = CubicSpline(t, y, bc_type="natural")
    assert np.allclose(nh.spline_eval(t, y, M, [4000.0, 9000.0, 30000.0]), cs([4000.0, 9000.0, 30000.0]), rtol=1e-12)


def test_active_atom_prd_and_polarizable_flags():
    """readAtom for an ACTIVE atom (readatom.c:247-258, 352-368): Ca II H & K carry the shape string PRD and become PRD
    lines only when PRD_N_MAX_ITER > 0; every Ca II line has term labels determinate() can read (|dJ| <= 1), so all
    five are polarizable -- their Zeeman patterns (zeeman.c:186-281) have 4 / 6 (J 1/2-1/2, 1/2-3/2) ... components
    with strengths normalised per q."""
    from pyrh_b200 import nlte_host, zeeman
    root = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "pyrh_path" / "rh" / "Atoms"
    if not (root / "CaII.atom").exists():
        pytest.skip("reference atom files not staged (oracle/_ref)")
    kw = dict(PRD_N_MAX_ITER="3", VMICRO_CHAR="5.0", B_STRENGTH_CHAR="0.0")
    at = nlte_host.read_active_atom(root / "CaII.atom", kw)
    prd = {(ln["j"], ln["i"]): ln["PRD"] for ln in at["lines"]}
    assert prd == {(3, 0): True, (4, 0): True, (3, 1): False, (4, 1): False, (4, 2): False}
    assert all(ln["polarizable"] for ln in at["lines"])
    at0 = nlte_host.read_active_atom(root / "CaII.atom", dict(kw, PRD_N_MAX_ITER="0"))
    assert not any(ln["PRD"] for ln in at0["lines"])
    for ln in at["lines"]:
        q, sh, st = zeeman.zeeman(at["label"][ln["i"]], at["g"][ln["i"]], at["label"][ln["j"]], at["g"][ln["j"]], ln["g_Lande_eff"])
        assert len(q) > 0 and set(q) <= {-1, 0, 1}
        for comp in (-1, 0, 1):
            assert abs(sum(s_ for q_, s_ in zip(q, st) if q_ == comp) - 1.0) < 1e-12


def test_unported_keywords_are_refused_not_ignored():
    """Keywords that change rhf1d()'s result and that the device path does not implement raise instead of being ignored."""
    from pyrh_b200 import host
    host.refuse_unported_keywords({"HYDROSTATIC": "FALSE", "BACKGROUND_POLARIZATION": "FALSE", "ATMOS_ITOP": "none"})
    for bad in ({"HYDROSTATIC": "TRUE"}, {"BACKGROUND_POLARIZATION": "TRUE"}, {"ATMOS_ITOP": "itop.dat"}):
        with pytest.raises(NotImplementedError):
            host.refuse_unported_keywords(bad)
