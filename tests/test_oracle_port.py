"""CPU tests: the C restatement (oracle/port) against golden vectors recorded
from the unmodified reference (oracle/gen_golden.py).  Bit-exact is the bar:
the port follows the reference's arithmetic order and is built without FMA."""
import numpy as np
import pytest

from conftest import port_objects
from oracle import portdriver as pd


def test_ltepops_elem_bit_exact(golden):
    name, g = golden
    lt, tab, col = port_objects(g)
    assert np.array_equal(pd.elem_pops(tab, col), g["elem_n"])


def test_rlk_opacity_bit_exact(golden):
    name, g = golden
    lt, tab, col = port_objects(g)
    for i, n in enumerate(g["sub"]):
        fl, chi, eta = pd.rlk_opacity(tab, col, g["elem_n"], g["lam_spect"][n], 1)
        assert fl == 3
        assert np.array_equal(chi, g["rlk_chi"][i])
        assert np.array_equal(eta, g["rlk_eta"][i])


def test_rlk_opacity_outside_window(golden_falc):
    g = golden_falc
    lt, tab, col = port_objects(g)
    fl, chi, eta = pd.rlk_opacity(tab, col, g["elem_n"], 500.0, 1)   # lambda_ref: no line within Q_WING
    assert fl == 0


def test_delo_bezier3_bit_exact(golden):
    name, g = golden
    for i, n in enumerate(g["sub"]):
        d = g["delo"][i]
        I, Psi = pd.stokes_bezier3(g["col_height"], float(g["muz"][0]), 1, d[0], d[1:5], d[10:13],
                                   g["col_T"], g["lam_spect"][n], want_psi=True)
        assert np.array_equal(I, d[5:9])
        assert np.array_equal(Psi, d[9])      # eval_operator=TRUE in Iterate -> Psi recorded


def test_lte_spectrum_bit_exact_scalar(golden):
    name, g = golden
    lt, tab, col = port_objects(g)
    k = g["lam_keep"]
    st = pd.lte_stokes_column(tab, col, g["lam_spect"][k], g["chi_ai"][k], g["eta_ai"][k])
    assert np.array_equal(st, g["stokes_scalar"])


def test_lte_spectrum_simd_variant_close(golden_falc):
    """-DSIMDON build: rcpss is CPU-defined, so only closeness is asserted here;
    the two reference builds themselves differ by ~3e-6 (hard part 1)."""
    g = golden_falc
    lt, tab, col = port_objects(g, matinv_simd=True)
    k = g["lam_keep"]
    st = pd.lte_stokes_column(tab, col, g["lam_spect"][k], g["chi_ai"][k], g["eta_ai"][k])
    ref = g["stokes_simd"]
    assert np.abs(st[0] / ref[0] - 1).max() < 1e-5
    d_builds = np.abs(g["stokes_simd"][0] / g["stokes_scalar"][0] - 1).max()
    assert 1e-8 < d_builds < 1e-4


def test_humlicek_regions_and_symmetry():
    rng = np.random.default_rng(1)
    for a, v in zip(10 ** rng.uniform(-4, 1.3, 200), rng.uniform(-30, 30, 200)):
        H, F = pd.voigt(a, v)
        H2, F2 = pd.voigt(a, -v)
        assert H == H2 and F == -F2
        assert H > 0
    # pure Doppler core and Lorentz wing limits (Humlicek's 1e-4 relative accuracy)
    H, _ = pd.voigt(1e-6, 0.0)
    assert abs(H - 1.0) < 2e-4
    H, _ = pd.voigt(0.1, 20.0)
    assert abs(H - 0.1 / (np.sqrt(np.pi) * 400.0)) / H < 1e-2


def test_matinv_scalar_inverts():
    import ctypes as C
    rng = np.random.default_rng(0)
    for _ in range(50):
        m = (np.eye(4) + 0.3 * rng.standard_normal((4, 4))).astype(np.float32)
        inv = m.copy()
        pd.lib().rp_matinv_scalar(inv.ctypes.data_as(C.POINTER(C.c_float)))
        assert np.abs(inv.astype(np.float64) @ m.astype(np.float64) - np.eye(4)).max() < 1e-4


def test_scalar_bezier3_and_feautrier_bit_exact():
    """NO_STOKES reference run on a grid wider than the line windows: Piecewise_Bezier3_1D inside,
    Feautrier outside (golden falc_scalar)."""
    from conftest import GOLD
    g = dict(np.load(GOLD / "falc_scalar.npz"))
    assert len(g["bez"]) > 50 and len(g["feau"]) > 20
    for m, d in zip(g["bez_meta"], g["bez"]):
        I, Psi = pd.bezier3_scalar(g["col_height"], float(g["muz"][m[1]]), int(m[2]), d[0], d[1],
                                   g["col_T"], g["lam_spect"][m[0]], want_psi=True)
        assert np.array_equal(I, d[2])
        if m[3]:
            assert np.array_equal(Psi, d[3])
    for m, d, Iem in zip(g["feau_meta"], g["feau"], g["feau_Iem"]):
        P, Psi, I0 = pd.feautrier(g["col_height"], float(g["muz"][m[1]]), d[0], d[1], g["col_T"],
                                  g["lam_spect"][m[0]])
        assert np.array_equal(P, d[2]) and I0 == Iem
        if m[3]:
            assert np.array_equal(Psi, d[3])


@pytest.mark.parametrize("fixture,niter", [("nlte_caii", 31), ("nlte_h_caii", 34)])
def test_nlte_port_bit_exact_all_iterations(fixture, niter):
    """MALI iteration (Opacity, addtoGamma/Coupling/Rates, statEquil, Ng) of the C restatement vs the
    reference's recorded Gamma, rates and populations: CaII 6-level atom on FAL-C (31 iterations) and
    BASELINE config 4, H 6-level + CaII both ACTIVE (895 wavelengths, 25 transitions, 34 iterations)."""
    from conftest import GOLD
    g = dict(np.load(GOLD / f"{fixture}.npz"))
    assert int(g["niter"]) == niter
    P = pd.PortNlte(g)
    it, nh, gh, rh, dh = P.iterate(int(g["hdr"][9]), float(g["hdr"][10]))
    assert it == int(g["niter"])
    assert np.array_equal(dh, g["dpops_iter"])
    assert np.array_equal(nh, g["n_iter"])
    ntr = P.ntr
    for idx, i in enumerate(g["iter_keep"]):
        assert np.array_equal(gh[i], g["gamma_iter"][idx])
        assert np.array_equal(rh[i][:ntr], g["rates_iter"][idx][0::2])
        assert np.array_equal(rh[i][ntr:], g["rates_iter"][idx][1::2])
    assert np.array_equal(nh[-1], g["n_final"]) and np.array_equal(nh[-1], g["pops_final"])
    assert np.array_equal(P.a["J"], g["J_final"])


def test_solve_linear_eq_bit_exact():
    from conftest import GOLD
    g = dict(np.load(GOLD / "nlte_caii.npz"))
    off = 0
    for N in g["lu_n"]:
        d = g["lu_data"][off:off + N * N + 2 * N]
        off += N * N + 2 * N
        x = pd.solve_linear_eq(d[:N * N].reshape(N, N), d[N * N:N * N + N])
        assert np.array_equal(x, d[N * N + N:])


def test_voigt_armstrong_port_bit_exact():
    import ctypes as C
    from conftest import GOLD
    g = dict(np.load(GOLD / "voigt_armstrong.npz"))
    f = pd.lib().rp_voigt_armstrong
    f.restype = C.c_double
    f.argtypes = [C.c_double, C.c_double]
    H = np.array([f(a, v) for a, v in zip(g["a"], g["v"])])
    assert np.array_equal(H, g["H"])
    assert set(np.unique(g["region"])) == {1, 2, 3}


def test_profile_port_bit_exact_both_angle_sets():
    """Profile() (profile.c:67) for the 3-ray MALI angle set and for the single-mu set _solveray()
    re-evaluates it on (pyrh_solveray.c:75-106; Damping() there is 1 ulp off the first call's)."""
    from conftest import GOLD
    g = dict(np.load(GOLD / "nlte_caii.npz"))
    for adamp, muz, wmu, phis, wphis in ((g["adamp"], g["muz"], g["wmu"], g["phi"], g["wphi"]),
                                         (g["fs_adamp"], g["fs_muz"], g["fs_wmu"], g["fs_phi"], g["fs_wphi"])):
        row, li = 0, 0
        for t in g["trans"]:
            if t[1] != 0:
                continue
            Nla, woff = int(t[5]), int(t[10])
            phi, wphi = pd.profile_line(g["line_lambda0"][li], g["tr_lambda"][woff:woff + Nla],
                                        g["tr_wlambda"][woff:woff + Nla], adamp[li], g["vbroad"][0], g["vel"],
                                        muz, wmu)
            assert np.array_equal(phi, phis[row:row + phi.shape[0]]) and np.array_equal(wphi, wphis[li])
            row += phi.shape[0]
            li += 1
        assert li == 5



def test_piecewise_and_delo_parabolic_port_bit_exact():
    """Piecewise_Linear_1D, Piecewise_1D (piecewise_1D.c:44,134) and Piece_Stokes_1D (piecestokes_1D.c:49)
    restatements vs calls recorded from the reference (LTE up/down rays; NLTE rays with Psi)."""
    from conftest import GOLD
    g = dict(np.load(GOLD / "falc_solvers.npz"))
    h, T, muz = g["col_height"], g["col_T"], g["muz"]
    for tag, kind in (("lin", "linear"), ("par", "parabolic")):
        assert {0, 1} <= set(g[tag + "_meta"][:, 2])
        for m, d in zip(g[tag + "_meta"], g[tag]):
            I = pd.piecewise_scalar(kind, h, float(muz[m[1]]), int(m[2]), d[0], d[1], T, g["lam_spect"][m[0]])
            assert np.array_equal(I, d[2])
        for m, d in zip(g[tag + "psi_meta"], g[tag + "psi"]):
            I, Psi = pd.piecewise_scalar(kind, g["n_height"], float(g["n_muz"][m[1]]), int(m[2]), d[0], d[1],
                                         g["n_T"], g["n_lam_spect"][m[0]], want_psi=True)
            assert np.array_equal(I, d[2]) and np.array_equal(Psi, d[3])
    for m, d in zip(g["pst_meta"], g["pst"]):
        I = pd.stokes_parabolic(h, float(muz[m[1]]), int(m[2]), d[0], d[1:5], d[10:13], T,
                                g["pst_lam_spect"][m[0]])
        assert np.array_equal(I, d[5:9])


def test_loggf_response_function_port_bit_exact():
    """Up-ray of Piecewise_Bezier3_1D with get_atomic_rfs (bezier_1D.c:416-516) vs the reference's recorded
    dI; the reference's returned rfs are dI at the top of the atmosphere (formal.c:278-282)."""
    from conftest import GOLD
    g = dict(np.load(GOLD / "falc_rf.npz"))
    assert np.array_equal(g["I_in"], g["down"][:, 2])          # the RF branch sees the down-ray solution
    for i, ns in enumerate(g["ns"]):
        I, dI = pd.bezier3_scalar_rf(g["col_height"], float(g["muz"][0]), g["up"][i, 0], g["up"][i, 1], g["col_T"],
                                     g["lam_spect"][ns], g["I_in"][i], g["dchi"][i], g["deta"][i])
        assert np.array_equal(I, g["up"][i, 2]) and np.array_equal(dI, g["dI"][i])
    keep = (g["lam_spect"] != 500.0)[g["ns"]]
    assert np.array_equal(g["dI"][keep][:, 0], g["rfs"])


def test_molecular_opacity_port_bit_exact():
    """MolecularOpacity + MolProfile (opacity.c:711-916) vs the reference's recorded calls: CN B-X lines
    around 847.3 nm, FAL-C with B = 1 kG (the shipped list is not polarizable -> VoigtArmstrong)."""
    from conftest import GOLD
    g = dict(np.load(GOLD / "falc_molecules.npz"))
    fl = g["flags"]
    assert len(g["mol_meta"]) == 12 and set(g["mol_meta"][:, 2]) == {0, 1}
    for m, d in zip(g["mol_meta"], g["mol_chi_eta"]):
        chi, eta, flg = pd.molecular_opacity(g["mlines"], g["zq"], g["zshift"], g["zstrength"], fl[6],
                                             g["lam_spect"][m[0]], float(g["muz"][m[1]]), bool(fl[0]), int(m[2]),
                                             g["col_T"], g["col_vel"], g["col_B"], g["col_cos_gamma"],
                                             g["col_cos_2chi"], g["col_sin_2chi"], g["mol"])
        assert np.array_equal(chi, d[0]) and np.array_equal(eta, d[1])
        assert flg == (m[3] | (m[4] << 1))


def test_passive_bb_port_bit_exact():
    """passive_bb (metal.c:174-344) vs the reference's recorded calls: Na I D and H-alpha of the PASSIVE
    Na.atom / H_6.atom on FAL-C, both directions."""
    from conftest import GOLD
    g = dict(np.load(GOLD / "falc_passive_bb.npz"))
    fl = g["flags"]
    assert len(g["pbb"]) == 36 and len(g["plines"]) == 2
    for m, d in zip(g["pbb_meta"], g["pbb"]):
        chi, eta, has = pd.passive_bb(g["plines"], g["c_shift"], g["c_fraction"], fl[6], g["lam_spect"][m[0]],
                                      float(g["muz"][m[1]]), bool(fl[0]), int(m[2]), g["col_vel"], g["pcol"])
        assert has == 1 and np.array_equal(chi, d[0]) and np.array_equal(eta, d[1])
