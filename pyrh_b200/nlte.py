"""Host-side mirror of the reference's NLTE interface: the MALI iteration of ``Iterate()``
(rh/iterate.c:48-143) for ACTIVE atoms (CRD, unpolarised), batched over independent columns.

The RH host keeps what it already does before ``Iterate`` (readAtomicModels, SortLambda active
sets, collisional rates, LTE populations, line profiles ``getProfiles``, background, ``initScatter``)
and hands the flat problem to ``rhb200_nlte_iterate``; it gets back converged populations and J.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib

TR_NFIELD = 16
(TR_ATOM, TR_TYPE, TR_I, TR_J, TR_NBLUE, TR_NLAMBDA, TR_AJI, TR_BJI, TR_BIJ, TR_ISOFRAC, TR_WOFF,
 TR_PHIROW, TR_KR, TR_LINEIDX, TR_LAMBDA0) = range(15)

dp, ip = _lib.dp, _lib.ip


class PlanStruct(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ("Nspect", "Nrays", "Ndep", "Natom", "Ntrans", "moving", "Ngorder",
                                         "Ngdelay", "Ngperiod", "isum", "bc_top", "bc_bottom", "ntrl",
                                         "nphirow", "nline")] +
                [("lam", dp), ("muz", dp), ("wmu", dp), ("atom_nlevel", ip), ("trans", dp),
                 ("tr_lambda", dp), ("tr_wlambda", dp), ("tr_alpha", dp), ("as_first", ip),
                 ("as_trans", ip), ("bg_hasline", ip)])


class ColumnsStruct(C.Structure):
    _fields_ = [(n, dp) for n in ("T", "height", "nstar", "ntotal", "C", "phi", "wphi", "adamp", "vbroad",
                                  "vel", "chi_c", "eta_c", "sca_c", "n", "J")]


class FrontStruct(C.Structure):
    """rhb200_nlte_front (include/rhb200.h)."""
    _fields_ = [("atom_model", ip), ("ncoll", C.c_int), ("ncolltab", C.c_int), ("coll", dp), ("coll_T", dp),
                ("coll_coef", dp), ("coll_M", dp), ("line_rows", dp), ("NmaxScatter", C.c_int), ("NmaxIter", C.c_int),
                ("iterLimit", C.c_double), ("plan1", C.POINTER(PlanStruct)),
                ("stokes", C.c_int), ("line_pol", ip), ("line_zoff", ip), ("zq", ip), ("zshift", dp), ("zstrength", dp),
                ("line_prd", ip), ("PRD_NmaxIter", C.c_int), ("PRDiterLimit", C.c_double)]


@dataclass
class NlteProblem:
    """Flat description of one NLTE problem (layout of include/rhb200.h, rhb200_nlte_plan/_columns).
    Per-column arrays carry a leading ``ncol`` axis."""
    hdr: dict
    lam: np.ndarray
    muz: np.ndarray
    wmu: np.ndarray
    atom_nlevel: np.ndarray
    trans: np.ndarray
    tr_lambda: np.ndarray
    tr_wlambda: np.ndarray
    tr_alpha: np.ndarray
    as_first: np.ndarray
    as_trans: np.ndarray
    bg_hasline: np.ndarray
    T: np.ndarray
    height: np.ndarray
    nstar: np.ndarray
    ntotal: np.ndarray
    C: np.ndarray
    phi: np.ndarray
    wphi: np.ndarray
    chi_c: np.ndarray
    eta_c: np.ndarray
    sca_c: np.ndarray
    n0: np.ndarray
    J0: np.ndarray
    adamp: np.ndarray | None = None      # [ncol, nline, Ndep]  Damping() output (host)
    vbroad: np.ndarray | None = None     # [ncol, Natom, Ndep]
    vel: np.ndarray | None = None        # [ncol, Ndep]

    @classmethod
    def from_golden(cls, g, ncol: int = 1) -> "NlteProblem":
        """Build from the fixture written by oracle/gen_golden_nlte.py, replicated to `ncol` columns."""
        h = g["hdr"]
        hdr = dict(Nspect=int(h[0]), Nrays=int(h[1]), Natom=int(h[2]), Ndep=int(h[3]), moving=int(h[4]),
                   Ngorder=int(h[5]), Ngdelay=int(h[6]), Ngperiod=int(h[7]), isum=int(h[8]),
                   NmaxIter=int(h[9]), iterLimit=float(h[10]), bc_top=int(h[13]), bc_bottom=int(h[14]))
        rep = lambda x: np.ascontiguousarray(np.broadcast_to(x, (ncol,) + x.shape), np.float64)   # noqa: E731
        trans = np.array(g["trans"], np.float64)
        if "line_lambda0" in g:
            lines = trans[:, TR_TYPE] == 0
            trans[lines, TR_LAMBDA0] = g["line_lambda0"][trans[lines, TR_LINEIDX].astype(int)]
        extra = {}
        if "adamp" in g:
            extra = dict(adamp=rep(g["adamp"]), vbroad=rep(g["vbroad"]), vel=rep(g["vel"]))
        return cls(**extra, hdr=hdr, lam=g["lam"], muz=g["muz"], wmu=g["wmu"], atom_nlevel=g["atom_nlevel"],
                   trans=trans, tr_lambda=g["tr_lambda"], tr_wlambda=g["tr_wlambda"],
                   tr_alpha=g["tr_alpha"], as_first=g["as_first"], as_trans=g["as_trans"],
                   bg_hasline=g["bgflags"][:, 0], T=rep(g["T"]), height=rep(g["height"]),
                   nstar=rep(g["nstar"]), ntotal=rep(g["ntotal"]), C=rep(g["C"]), phi=rep(g["phi"]),
                   wphi=rep(g["wphi"]), chi_c=rep(g["bg"][0]), eta_c=rep(g["bg"][1]), sca_c=rep(g["bg"][2]),
                   n0=rep(g["n0"]), J0=rep(g["J0"]))


def _structs(prob: NlteProblem, device_profiles: bool):
    f64 = lambda x: np.ascontiguousarray(x, np.float64)   # noqa: E731
    i32 = lambda x: np.ascontiguousarray(x, np.int32)     # noqa: E731
    h = prob.hdr
    keep = dict(lam=f64(prob.lam), muz=f64(prob.muz), wmu=f64(prob.wmu), atom_nlevel=i32(prob.atom_nlevel),
                trans=f64(prob.trans), tr_lambda=f64(prob.tr_lambda), tr_wlambda=f64(prob.tr_wlambda),
                tr_alpha=f64(prob.tr_alpha), as_first=i32(prob.as_first), as_trans=i32(prob.as_trans),
                bg_hasline=i32(prob.bg_hasline))
    ptr = lambda a: a.ctypes.data_as(ip if a.dtype == np.int32 else dp)   # noqa: E731
    plan = PlanStruct(h["Nspect"], len(keep["muz"]), h["Ndep"], h["Natom"], keep["trans"].shape[0], h["moving"],
                      h["Ngorder"], h["Ngdelay"], h["Ngperiod"], h["isum"], h["bc_top"], h["bc_bottom"],
                      len(keep["tr_lambda"]), prob.phi.shape[1], prob.wphi.shape[1],
                      *[ptr(keep[k]) for k in ("lam", "muz", "wmu", "atom_nlevel", "trans", "tr_lambda",
                                               "tr_wlambda", "tr_alpha", "as_first", "as_trans", "bg_hasline")])
    n, J = f64(prob.n0).copy(), f64(prob.J0).copy()
    names = ("T", "height", "nstar", "ntotal", "C", "phi", "wphi", "adamp", "vbroad", "vel",
             "chi_c", "eta_c", "sca_c")
    skip = {"phi", "wphi"} if device_profiles else {"adamp", "vbroad", "vel"}
    cols_keep = {k: f64(getattr(prob, k)) for k in names if k not in skip and getattr(prob, k) is not None}
    cols = ColumnsStruct(*[ptr(cols_keep[k]) if k in cols_keep else None for k in names], ptr(n), ptr(J))
    return plan, cols, n, J, (keep, cols_keep), ptr


def formal(ctx, prob: NlteProblem, npass: int = 1, update_J: int = 0, limit: float = 0.0,
           device_profiles: bool = False):
    """``solveSpectrum(FALSE, FALSE)`` repeated: initScatter (update_J) or the final pass of _solveray.
    Returns dict(J, Iem [ncol, Nspect, Nrays], npass)."""
    plan, cols, n, J, keep, ptr = _structs(prob, device_profiles)
    ncol = prob.T.shape[0]
    Iem = np.zeros((ncol, prob.hdr["Nspect"], len(prob.muz)))
    done = np.zeros(ncol, np.int32)
    _lib.check(ctx.lib.rhb200_nlte_formal(ctx.h, C.byref(plan), ncol, C.byref(cols), int(npass), int(update_J),
                                          float(limit), ptr(Iem), done.ctypes.data_as(ip)))
    return dict(J=J, Iem=Iem, npass=done)


def single_mu_problem(g, mu: float | None = None, ncol: int = 1) -> NlteProblem:
    """The problem _solveray() solves after convergence (pyrh_solveray.c:75-106): one angle,
    profiles and background recomputed for it (fixture keys fs_*), converged n and J."""
    prob = NlteProblem.from_golden(g, ncol=ncol)
    rep = lambda x: np.ascontiguousarray(np.broadcast_to(x, (ncol,) + x.shape), np.float64)   # noqa: E731
    prob.muz = np.array(g["fs_muz"] if mu is None else [mu], np.float64)
    prob.wmu = np.array(g["fs_wmu"], np.float64)
    prob.hdr = dict(prob.hdr, Nrays=1)
    tr = prob.trans.copy()
    row = 0
    for t in tr:
        if t[TR_TYPE] == 0:
            t[TR_PHIROW] = row
            row += 2 * int(t[TR_NLAMBDA])
    prob.trans = tr
    prob.phi, prob.wphi = rep(g["fs_phi"]), rep(g["fs_wphi"])
    if "fs_adamp" in g:       # Damping() re-evaluated after the second Background(): differs by 1 ulp
        prob.adamp = rep(g["fs_adamp"])
    prob.chi_c, prob.eta_c, prob.sca_c = rep(g["fs_bg"][0]), rep(g["fs_bg"][1]), rep(g["fs_bg"][2])
    prob.bg_hasline = g["fs_bgflags"][:, 0]
    prob.n0, prob.J0 = rep(g["n_final"]), rep(g["J_final"])
    return prob


def iterate(ctx, prob: NlteProblem, nmax: int | None = None, limit: float | None = None,
            dump_iter: int = 0, device_profiles: bool = False, nscatter: int = 0):
    """Run the MALI iteration on the GPU.  Returns dict(n, J, niter, dpops[, gamma, rij, rji]).
    With ``device_profiles`` the line profiles are evaluated on the device from adamp/vbroad/vel
    (``Profile()``) instead of being passed in; they are returned as ``phi``/``wphi``."""
    f64 = lambda x: np.ascontiguousarray(x, np.float64)   # noqa: E731
    i32 = lambda x: np.ascontiguousarray(x, np.int32)     # noqa: E731
    h = prob.hdr
    nmax = h["NmaxIter"] if nmax is None else nmax
    limit = h["iterLimit"] if limit is None else limit
    keep = dict(lam=f64(prob.lam), muz=f64(prob.muz), wmu=f64(prob.wmu), atom_nlevel=i32(prob.atom_nlevel),
                trans=f64(prob.trans), tr_lambda=f64(prob.tr_lambda), tr_wlambda=f64(prob.tr_wlambda),
                tr_alpha=f64(prob.tr_alpha), as_first=i32(prob.as_first), as_trans=i32(prob.as_trans),
                bg_hasline=i32(prob.bg_hasline))
    ptr = lambda a: a.ctypes.data_as(ip if a.dtype == np.int32 else dp)   # noqa: E731
    plan = PlanStruct(h["Nspect"], h["Nrays"], h["Ndep"], h["Natom"], keep["trans"].shape[0], h["moving"],
                      h["Ngorder"], h["Ngdelay"], h["Ngperiod"], h["isum"], h["bc_top"], h["bc_bottom"],
                      len(keep["tr_lambda"]), prob.phi.shape[1], prob.wphi.shape[1],
                      *[ptr(keep[k]) for k in ("lam", "muz", "wmu", "atom_nlevel", "trans", "tr_lambda",
                                               "tr_wlambda", "tr_alpha", "as_first", "as_trans", "bg_hasline")])
    ncol = prob.T.shape[0]
    n, J = f64(prob.n0).copy(), f64(prob.J0).copy()
    names = ("T", "height", "nstar", "ntotal", "C", "phi", "wphi", "adamp", "vbroad", "vel",
             "chi_c", "eta_c", "sca_c")
    skip = {"phi", "wphi"} if device_profiles else {"adamp", "vbroad", "vel"}
    cols_keep = {k: f64(getattr(prob, k)) for k in names if k not in skip and getattr(prob, k) is not None}
    cols = ColumnsStruct(*[ptr(cols_keep[k]) if k in cols_keep else None for k in names], ptr(n), ptr(J))
    phi_out = np.zeros(prob.phi.shape) if device_profiles else None
    wphi_out = np.zeros(prob.wphi.shape) if device_profiles else None
    niter = np.zeros(ncol, np.int32)
    dpops = np.zeros((ncol, max(nmax, 1)))
    ngam = int(np.sum(np.asarray(prob.atom_nlevel) ** 2))
    ntr = keep["trans"].shape[0]
    gam = np.zeros((ncol, ngam, h["Ndep"])) if dump_iter else None
    rates = np.zeros((2, ncol, ntr, h["Ndep"])) if dump_iter else None
    lib = ctx.lib
    _lib.check(lib.rhb200_nlte_iterate(ctx.h, C.byref(plan), ncol, C.byref(cols), int(nscatter), int(nmax), float(limit),
                                       niter.ctypes.data_as(ip), ptr(dpops), int(dump_iter),
                                       ptr(gam) if dump_iter else None, ptr(rates) if dump_iter else None,
                                       ptr(phi_out) if device_profiles else None,
                                       ptr(wphi_out) if device_profiles else None))
    out = dict(n=n, J=J, niter=niter, dpops=dpops)
    if device_profiles:
        out.update(phi=phi_out, wphi=wphi_out)
    if dump_iter:
        out.update(gamma=gam, rij=rates[0], rji=rates[1])
    return out


def solve_linear_eq(ctx, A, b, improve=True):
    A = np.ascontiguousarray(A, np.float64)
    b = np.ascontiguousarray(b, np.float64).copy()
    nsys, N = b.shape
    _lib.check(ctx.lib.rhb200_solve_linear_eq_batch(ctx.h, nsys, N, A.ctypes.data_as(dp), b.ctypes.data_as(dp),
                                                    int(improve)))
    return b


def shard_range(prob: NlteProblem, rank: int, nrank: int) -> tuple[int, int]:
    """Wavelength chunk [ns_lo, ns_hi) of ``rank`` when one atmosphere is split over ``nrank`` GPUs
    (rhb200_nlte_shard_range; host-only, balanced by ray count)."""
    plan, _cols, _n, _J, _keep, _ptr = _structs(prob, False)
    lo, hi = C.c_int(0), C.c_int(0)
    _lib.check(_lib.load().rhb200_nlte_shard_range(C.byref(plan), int(rank), int(nrank), C.byref(lo), C.byref(hi)))
    return lo.value, hi.value
