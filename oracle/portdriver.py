"""oracle/portdriver.py -- TEST INFRASTRUCTURE ONLY.

ctypes bindings for our plain-C restatement of the reference hot path
(oracle/port/*.c -> oracle/_build/liboracle_port.so).  Used by tests/ as the
checker for the CUDA path, by __graft_entry__.smoke(), and by bench.py's
cpu_baseline leg.  Never imported by pyrh_b200.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "liboracle_port.so"

RL_NFIELD, RE_NFIELD, RE_MAXSTAGE = 24, 16, 12
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


class RpLineTable(C.Structure):
    _fields_ = [("nline", C.c_int), ("nelem", C.c_int), ("npf", C.c_int), ("matinv_simd", C.c_int),
                ("lines", dp), ("zq", ip), ("zshift", dp), ("zstrength", dp),
                ("elems", dp), ("pf", dp), ("Tpf", dp),
                ("vmicro_char", C.c_double), ("magneto_optical", C.c_int)]


class RpColumn(C.Structure):
    _fields_ = [("Ndep", C.c_int)] + [(n, dp) for n in (
        "T", "ne", "vturb", "vel", "B", "cos_gamma", "cos_2chi", "sin_2chi",
        "nHtot", "np", "height")] + [("muz", C.c_double), ("moving", C.c_int)]


def build(force: bool = False):
    if force or not LIB.exists():
        subprocess.check_call(["make", "-C", str(HERE), "port"], stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(LIB))
        _lib.rp_voigt_humlicek.restype = C.c_double
        _lib.rp_voigt_humlicek.argtypes = [C.c_double, C.c_double, dp]
        _lib.rp_humlicek_region.argtypes = [C.c_double, C.c_double]
        _lib.rp_planck.restype = C.c_double
        _lib.rp_planck.argtypes = [C.c_double, C.c_double]
        _lib.rp_cent_deriv.restype = C.c_double
        _lib.rp_cent_deriv.argtypes = [C.c_double] * 5
        _lib.rp_rlk_opacity.restype = C.c_int
    return _lib


def _d(a):
    return a.ctypes.data_as(dp)


class PortTables:
    """Keeps numpy arrays alive behind an rp_linetable."""

    def __init__(self, lt, matinv_simd=False):
        # lt: pyrh_b200.linelist.LineTable-like (attributes lines, zq, zshift, zstrength, elems, pf, Tpf, vmicro_char)
        self.a = dict(lines=np.ascontiguousarray(lt.lines, np.float64),
                      zq=np.ascontiguousarray(lt.zq, np.int32),
                      zshift=np.ascontiguousarray(lt.zshift, np.float64),
                      zstrength=np.ascontiguousarray(lt.zstrength, np.float64),
                      elems=np.ascontiguousarray(lt.elems, np.float64),
                      pf=np.ascontiguousarray(lt.pf, np.float64),
                      Tpf=np.ascontiguousarray(lt.Tpf, np.float64))
        a = self.a
        self.c = RpLineTable(a["lines"].shape[0], a["elems"].shape[0], a["Tpf"].shape[0],
                             int(matinv_simd), _d(a["lines"]), a["zq"].ctypes.data_as(ip),
                             _d(a["zshift"]), _d(a["zstrength"]), _d(a["elems"]), _d(a["pf"]),
                             _d(a["Tpf"]), float(lt.vmicro_char), 0)


class PortColumn:
    FIELDS = ("T", "ne", "vturb", "vel", "B", "cos_gamma", "cos_2chi", "sin_2chi",
              "nHtot", "np", "height")

    def __init__(self, muz=1.0, moving=True, **arrs):
        self.a = {k: np.ascontiguousarray(arrs[k], np.float64) for k in self.FIELDS}
        n = self.a["T"].shape[0]
        self.c = RpColumn(n, *[_d(self.a[k]) for k in self.FIELDS], float(muz), int(moving))
        self.Ndep = n


def ltepops_elem(tab: PortTables, col: PortColumn, ielem: int) -> np.ndarray:
    nst = int(tab.a["elems"][ielem, 2])
    out = np.zeros((nst, col.Ndep))
    lib().rp_ltepops_elem(C.byref(tab.c), ielem, C.byref(col.c), _d(out))
    return out


def elem_pops(tab: PortTables, col: PortColumn) -> np.ndarray:
    ne = tab.a["elems"].shape[0]
    out = np.zeros((ne, RE_MAXSTAGE, col.Ndep))
    for ie in range(ne):
        n = ltepops_elem(tab, col, ie)
        out[ie, : n.shape[0]] = n
    return out


def rlk_opacity(tab: PortTables, col: PortColumn, elem_n: np.ndarray, lam: float, to_obs: int = 1):
    chi = np.zeros((4, col.Ndep))
    eta = np.zeros((4, col.Ndep))
    elem_n = np.ascontiguousarray(elem_n)
    fl = lib().rp_rlk_opacity(C.byref(tab.c), C.byref(col.c), _d(elem_n), C.c_double(lam),
                              int(to_obs), _d(chi), _d(eta))
    return fl, chi, eta


def stokes_bezier3(height, muz, to_obs, chi, S, chiQUV, T, lam, bc_top=1, bc_bottom=2,
                   matinv_simd=False, want_psi=False):
    n = len(chi)
    arr = [np.ascontiguousarray(x, np.float64) for x in (height, chi, S, chiQUV, T)]
    I = np.zeros((4, n))
    Psi = np.zeros(n)
    lib().rp_stokes_bezier3(n, _d(arr[0]), C.c_double(muz), int(to_obs), _d(arr[1]), _d(arr[2]),
                            _d(arr[3]), _d(arr[4]), C.c_double(lam), int(bc_top), int(bc_bottom),
                            int(matinv_simd), _d(I), _d(Psi) if want_psi else None)
    return (I, Psi) if want_psi else I


def bezier3_scalar(height, muz, to_obs, chi, S, T, lam, bc_top=1, bc_bottom=2, want_psi=False):
    n = len(chi)
    arr = [np.ascontiguousarray(x, np.float64) for x in (height, chi, S, T)]
    I = np.zeros(n)
    Psi = np.zeros(n)
    lib().rp_bezier3_scalar(n, _d(arr[0]), C.c_double(muz), int(to_obs), _d(arr[1]), _d(arr[2]),
                            _d(arr[3]), C.c_double(lam), int(bc_top), int(bc_bottom), _d(I),
                            _d(Psi) if want_psi else None)
    return (I, Psi) if want_psi else I


def bezier3_scalar_rf(height, muz, chi, S, T, lam, I_in, dchi, deta, bc_top=1, bc_bottom=2):
    """Up-ray of Piecewise_Bezier3_1D with the log gf response function; I_in = the down-ray solution
    left in the buffer.  dchi, deta: [Ndep, npar].  Returns (I, dI[Ndep, npar])."""
    n = len(chi)
    arr = [np.ascontiguousarray(x, np.float64) for x in (height, chi, S, T, dchi, deta)]
    npar = arr[4].shape[1]
    I = np.ascontiguousarray(I_in, np.float64).copy()
    dI = np.zeros((n, npar))
    f = lib().rp_bezier3_scalar_rf
    f.restype = None
    f(n, _d(arr[0]), C.c_double(muz), 1, _d(arr[1]), _d(arr[2]), _d(arr[3]), C.c_double(lam),
      int(bc_top), int(bc_bottom), _d(I), None, int(npar), _d(arr[4]), _d(arr[5]), _d(dI))
    return I, dI


def piecewise_scalar(kind, height, muz, to_obs, chi, S, T, lam, bc_top=1, bc_bottom=2, want_psi=False):
    """kind: 'linear' (Piecewise_Linear_1D) | 'parabolic' (Piecewise_1D), piecewise_1D.c:44,134."""
    n = len(chi)
    arr = [np.ascontiguousarray(x, np.float64) for x in (height, chi, S, T)]
    I = np.zeros(n)
    Psi = np.zeros(n)
    f = {"linear": lib().rp_piecewise_linear, "parabolic": lib().rp_piecewise_parabolic}[kind]
    f(n, _d(arr[0]), C.c_double(muz), int(to_obs), _d(arr[1]), _d(arr[2]), _d(arr[3]), C.c_double(lam),
      int(bc_top), int(bc_bottom), _d(I), _d(Psi) if want_psi else None)
    return (I, Psi) if want_psi else I


def stokes_parabolic(height, muz, to_obs, chi, S, chiQUV, T, lam, bc_top=1, bc_bottom=2, want_psi=False):
    """Piece_Stokes_1D, piecestokes_1D.c:49-174."""
    n = len(chi)
    arr = [np.ascontiguousarray(x, np.float64) for x in (height, chi, S, chiQUV, T)]
    I = np.zeros((4, n))
    Psi = np.zeros(n)
    lib().rp_stokes_parabolic(n, _d(arr[0]), C.c_double(muz), int(to_obs), _d(arr[1]), _d(arr[2]),
                              _d(arr[3]), _d(arr[4]), C.c_double(lam), int(bc_top), int(bc_bottom),
                              _d(I), _d(Psi) if want_psi else None)
    return (I, Psi) if want_psi else I


def feautrier(height, muz, chi, S, T, lam, bc_top=1, bc_bottom=2):
    n = len(chi)
    arr = [np.ascontiguousarray(x, np.float64) for x in (height, chi, S, T)]
    P, Psi = np.zeros(n), np.zeros(n)
    f = lib().rp_feautrier
    f.restype = C.c_double
    I0 = f(n, _d(arr[0]), C.c_double(muz), _d(arr[1]), _d(arr[2]), _d(arr[3]), C.c_double(lam),
           int(bc_top), int(bc_bottom), _d(P), _d(Psi))
    return P, Psi, I0


def lte_stokes_column(tab: PortTables, col: PortColumn, lam, chi_ai, eta_ai, bc_top=1, bc_bottom=2):
    lam = np.ascontiguousarray(lam, np.float64)
    chi_ai = np.ascontiguousarray(chi_ai, np.float64)
    eta_ai = np.ascontiguousarray(eta_ai, np.float64)
    out = np.zeros((4, len(lam)))
    lib().rp_lte_stokes_column(C.byref(tab.c), C.byref(col.c), len(lam), _d(lam), _d(chi_ai),
                               _d(eta_ai), int(bc_top), int(bc_bottom), _d(out))
    return out


def voigt(a, v):
    F = C.c_double()
    H = lib().rp_voigt_humlicek(float(a), float(v), C.byref(F))
    return H, F.value


def region_hist(tab: PortTables, col: PortColumn, lam, to_obs: int = 1):
    lam = np.ascontiguousarray(lam, np.float64)
    h = (C.c_long * 5)()
    lib().rp_region_hist(C.byref(tab.c), C.byref(col.c), len(lam), _d(lam), int(to_obs), h)
    return np.array(list(h))


# ------------------------------------------------------------------ NLTE port
class RpNlte(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ("Nspect", "Nrays", "Ndep", "Natom", "Ntrans", "moving", "Ngorder",
                                         "Ngdelay", "Ngperiod", "isum", "bc_top", "bc_bottom")] +
                [(n, dp) for n in ("lam", "muz", "wmu", "T", "height")] + [("atom_nlevel", ip)] +
                [(n, dp) for n in ("trans", "tr_lambda", "tr_wlambda", "tr_alpha")] +
                [(n, ip) for n in ("as_first", "as_trans", "bg_hasline")] +
                [(n, dp) for n in ("nstar", "ntotal", "C", "phi", "wphi", "chi_c", "eta_c", "sca_c",
                                   "n", "J", "Gamma", "Rij", "Rji")] + [("updateJ", C.c_int), ("Iem", dp)])


class PortNlte:
    """Builds an rp_nlte from the flat golden layout (oracle/gen_golden_nlte.py) and keeps the arrays alive."""

    def __init__(self, g):
        hdr = g["hdr"]
        f64 = lambda x: np.ascontiguousarray(x, np.float64)   # noqa: E731
        i32 = lambda x: np.ascontiguousarray(x, np.int32)     # noqa: E731
        N = int(hdr[3])
        a = dict(lam=f64(g["lam"]), muz=f64(g["muz"]), wmu=f64(g["wmu"]), T=f64(g["T"]), height=f64(g["height"]),
                 atom_nlevel=i32(g["atom_nlevel"]), trans=f64(g["trans"]), tr_lambda=f64(g["tr_lambda"]),
                 tr_wlambda=f64(g["tr_wlambda"]), tr_alpha=f64(g["tr_alpha"]), as_first=i32(g["as_first"]),
                 as_trans=i32(g["as_trans"]), bg_hasline=i32(g["bgflags"][:, 0]), nstar=f64(g["nstar"]),
                 ntotal=f64(g["ntotal"]), C=f64(g["C"]), phi=f64(g["phi"]), wphi=f64(g["wphi"]),
                 chi_c=f64(g["bg"][0]), eta_c=f64(g["bg"][1]), sca_c=f64(g["bg"][2]),
                 n=f64(g["n0"]).copy(), J=f64(g["J0"]).copy())
        ntr = a["trans"].shape[0]
        a["Gamma"] = np.zeros_like(a["C"])
        a["Rij"], a["Rji"] = np.zeros((ntr, N)), np.zeros((ntr, N))
        a["Iem"] = np.zeros((int(hdr[0]), len(a["muz"])))
        self.a = a
        ptr = lambda k: a[k].ctypes.data_as(ip if a[k].dtype == np.int32 else dp)   # noqa: E731
        self.c = RpNlte(int(hdr[0]), len(a["muz"]), N, int(hdr[2]), ntr, int(hdr[4]), int(hdr[5]), int(hdr[6]),
                        int(hdr[7]), int(hdr[8]), int(hdr[13]), int(hdr[14]),
                        *[ptr(k) for k in ("lam", "muz", "wmu", "T", "height", "atom_nlevel", "trans", "tr_lambda",
                                           "tr_wlambda", "tr_alpha", "as_first", "as_trans", "bg_hasline", "nstar",
                                           "ntotal", "C", "phi", "wphi", "chi_c", "eta_c", "sca_c", "n", "J",
                                           "Gamma", "Rij", "Rji")], 1, ptr("Iem"))
        self.N, self.ntr = N, ntr

    def solve_spectrum(self, eval_operator=False, updateJ=True):
        f = lib().rp_nlte_solve_spectrum
        f.restype = C.c_double
        self.c.updateJ = int(updateJ)
        d = f(C.byref(self.c), int(eval_operator))
        self.c.updateJ = 1
        return d

    def iterate(self, nmax, limit):
        a = self.a
        nh = np.zeros((nmax,) + a["n"].shape)
        gh = np.zeros((nmax,) + a["C"].shape)
        rh = np.zeros((nmax, 2 * self.ntr, self.N))
        dh = np.zeros(nmax)
        f = lib().rp_nlte_iterate
        f.restype = C.c_int
        it = f(C.byref(self.c), int(nmax), C.c_double(limit), _d(nh), _d(gh), _d(rh), _d(dh))
        return it, nh[:it], gh[:it], rh[:it], dh[:it]


def solve_linear_eq(A, b, improve=True):
    A = np.ascontiguousarray(A, np.float64).copy()
    b = np.ascontiguousarray(b, np.float64).copy()
    lib().rp_solve_linear_eq(A.shape[0], _d(A), _d(b), int(improve))
    return b


def stat_equil(Gamma, ntotal, n, isum=-1):
    n = np.ascontiguousarray(n, np.float64).copy()
    Gamma = np.ascontiguousarray(Gamma, np.float64)
    ntotal = np.ascontiguousarray(ntotal, np.float64)
    lib().rp_stat_equil(n.shape[0], n.shape[1], _d(Gamma), _d(ntotal), int(isum), _d(n))
    return n


def profile_line(lambda0, lam, wlam, adamp, vbroad, vel, muz, wmu):
    """Profile() (profile.c:67-372) of one un-polarised single-component line: phi [2*Nrays*Nla][N], wphi [N]."""
    L = lib()
    dp = C.POINTER(C.c_double)
    L.rp_profile_line.restype = None
    L.rp_profile_line.argtypes = [C.c_int] * 3 + [C.c_double] + [dp] * 9
    arrs = [np.ascontiguousarray(x, np.float64) for x in (lam, wlam, adamp, vbroad, vel, muz, wmu)]
    N, Nrays, Nla = len(arrs[2]), len(arrs[5]), len(arrs[0])
    phi, wphi = np.zeros((2 * Nrays * Nla, N)), np.zeros(N)
    L.rp_profile_line(N, Nrays, Nla, float(lambda0), *[_d(a) for a in arrs], _d(phi), _d(wphi))
    return phi, wphi


def molecular_opacity(mlines, zq, zshift, zstrength, vmicro_char, lam, muz, moving, to_obs,
                      T, vel, B, cos_gamma, cos_2chi, sin_2chi, mol):
    """MolecularOpacity() (opacity.c:711-839) of one column at one wavelength: (chi[4,N], eta[4,N], flags)."""
    f64 = lambda x: np.ascontiguousarray(x, np.float64)   # noqa: E731
    ml, mol = f64(mlines), f64(mol)
    N = len(T)
    zq_ = np.ascontiguousarray(zq if len(zq) else [0], np.int32)
    zs, zt = f64(zshift if len(zshift) else [0.0]), f64(zstrength if len(zstrength) else [0.0])
    arrs = [f64(x) for x in (T, vel, B, cos_gamma, cos_2chi, sin_2chi)]
    chi, eta = np.zeros((4, N)), np.zeros((4, N))
    fn = lib().rp_molecular_opacity
    fn.restype = C.c_int
    fl = fn(N, mol.shape[0], ml.shape[0], _d(ml), zq_.ctypes.data_as(C.POINTER(C.c_int)), _d(zs), _d(zt),
            C.c_double(vmicro_char), C.c_double(lam), C.c_double(muz), int(moving), int(to_obs),
            *[_d(a) for a in arrs], _d(mol), _d(chi), _d(eta))
    return chi, eta, fl


def passive_bb(plines, c_shift, c_fraction, vmicro_char, lam, muz, moving, to_obs, vel, pcol):
    """passive_bb() (metal.c:174-344) of one column at one wavelength: (chi[N], eta[N], hasline)."""
    f64 = lambda x: np.ascontiguousarray(x, np.float64)   # noqa: E731
    pl, pc, cs, cf, vel = f64(plines), f64(pcol), f64(c_shift), f64(c_fraction), f64(vel)
    N = len(vel)
    chi, eta = np.zeros(N), np.zeros(N)
    fn = lib().rp_passive_bb
    fn.restype = C.c_int
    has = fn(N, pl.shape[0], _d(pl), _d(cs), _d(cf), C.c_double(vmicro_char), C.c_double(lam), C.c_double(muz),
             int(moving), int(to_obs), _d(vel), _d(pc), _d(chi), _d(eta))
    return chi, eta, has

