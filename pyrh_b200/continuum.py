"""Host-side mirror of the background-continuum interface (include/rhb200.h, rhb200_continuum_model).

The RH host fills the model once (level table, bound-free continua with their cross-section tables, the lines
Rayleigh() sums, the published opacity tables it holds) and hands per-column populations to
``continuum_batch``; reference: Background() rh/background.c:343-465 and the routines it calls.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

dp = _lib.dp

_PTRS = ("lev", "bf", "tab_lambda", "tab_alpha", "ray")
_TABLES = (("hmbf_lambda", "hmbf_alpha"), ("hmff_lambda", "hmff_theta", "hmff_kappa"),
           ("h2mff_lambda", "h2mff_theta", "h2mff_kappa"), ("h2pff_lambda", "h2pff_temp", "h2pff_kappa"),
           ("rh2_a", "rh2_lambda", "rh2_sigma"), ("oh_T", "oh_E", "oh_cross"), ("ch_T", "ch_E", "ch_cross"))


class ModelStruct(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ("natom", "nlev", "ncont", "ntab", "nray")] +
                [(n, dp) for n in _PTRS] +
                [(n, C.c_int) for n in ("nlev_H", "atom_He", "H_active", "has_OH", "has_CH", "has_H2", "solve_NLTE",
                                        "do_fudge")] +
                [("vmicro_char", C.c_double)] +
                [("hmbf_lambda", dp), ("hmbf_alpha", dp), ("n_hmbf", C.c_int),
                 ("hmff_lambda", dp), ("hmff_theta", dp), ("hmff_kappa", dp), ("n_hmff_lambda", C.c_int), ("n_hmff_theta", C.c_int),
                 ("h2mff_lambda", dp), ("h2mff_theta", dp), ("h2mff_kappa", dp), ("n_h2mff_lambda", C.c_int), ("n_h2mff_theta", C.c_int),
                 ("h2pff_lambda", dp), ("h2pff_temp", dp), ("h2pff_kappa", dp), ("n_h2pff_lambda", C.c_int), ("n_h2pff_temp", C.c_int),
                 ("rh2_a", dp), ("rh2_lambda", dp), ("rh2_sigma", dp), ("n_rh2", C.c_int),
                 ("oh_T", dp), ("oh_E", dp), ("oh_cross", dp), ("n_oh_T", C.c_int), ("n_oh_E", C.c_int),
                 ("ch_T", dp), ("ch_E", dp), ("ch_cross", dp), ("n_ch_T", C.c_int), ("n_ch_E", C.c_int),
                 ("n_fudge", C.c_int), ("fudge_lambda", dp), ("fudge", dp)])


class ContinuumModel:
    """Keeps the numpy arrays alive and exposes the C struct.  ``g`` is a mapping with the keys of the fixture
    tests/golden/falc_continuum.npz (the flat form of atmos.atoms[] + the reference's static tables)."""

    def __init__(self, g, fudge_wave=None, fudge_value=None):
        f64 = lambda x: np.ascontiguousarray(x, np.float64)   # noqa: E731
        h = g["ct_hdr"]
        self.a = {k: f64(g["ct_" + k]) for k in _PTRS}
        for grp in _TABLES:
            for k in grp:
                self.a[k] = f64(g["tab_" + k])
        a = self.a
        s = ModelStruct()
        s.natom, s.nlev, s.ncont, s.ntab, s.nray = int(h[0]), int(h[1]), int(h[2]), int(h[3]), int(h[4])
        for k in _PTRS:
            setattr(s, k, a[k].ctypes.data_as(dp))
        s.nlev_H, s.atom_He = int(h[13]), (1 if h[6] else -1)
        s.H_active, s.has_OH, s.has_CH, s.has_H2 = int(h[5]), int(h[7]), int(h[8]), int(h[9])
        s.solve_NLTE, s.do_fudge, s.vmicro_char = int(h[10]), int(h[12]), float(h[11])
        for grp in _TABLES:
            for k in grp:
                setattr(s, k, a[k].ctypes.data_as(dp))
        s.n_hmbf = len(a["hmbf_lambda"])
        s.n_hmff_lambda, s.n_hmff_theta = len(a["hmff_lambda"]), len(a["hmff_theta"])
        s.n_h2mff_lambda, s.n_h2mff_theta = len(a["h2mff_lambda"]), len(a["h2mff_theta"])
        s.n_h2pff_lambda, s.n_h2pff_temp = len(a["h2pff_lambda"]), len(a["h2pff_temp"])
        s.n_rh2 = len(a["rh2_lambda"])
        s.n_oh_T, s.n_oh_E = len(a["oh_T"]), len(a["oh_E"])
        s.n_ch_T, s.n_ch_E = len(a["ch_T"]), len(a["ch_E"])
        s.n_fudge = 0
        if fudge_wave is not None:                      # pyrh.compute1d's fudge_wave [n] / fudge_value [3, n]
            self.a["fudge_lambda"] = f64(fudge_wave)
            self.a["fudge"] = f64(fudge_value)
            if self.a["fudge"].shape != (3, len(self.a["fudge_lambda"])):
                raise ValueError("fudge_value must be [3, len(fudge_wave)] (H-, scattering, metals)")
            s.do_fudge, s.n_fudge = 1, len(self.a["fudge_lambda"])
            s.fudge_lambda, s.fudge = self.a["fudge_lambda"].ctypes.data_as(dp), self.a["fudge"].ctypes.data_as(dp)
        self.struct = s
        self.nlev = s.nlev


def continuum_batch(ctx, model: ContinuumModel, lam, T, ne, nHmin, nH2, nOH, nCH, n, nstar=None, contrib=False):
    """chi_ai, eta_ai, sca_ai [ncol, nlambda, ndep] for per-column inputs T, ne, nHmin, nH2, nOH, nCH [ncol, ndep]
    and level populations n (and nstar, default = n) [ncol, nlev, ndep]."""
    f64 = lambda x: None if x is None else np.ascontiguousarray(x, np.float64)   # noqa: E731
    lam, T, ne, nHmin, nH2, nOH, nCH, n = map(f64, (lam, T, ne, nHmin, nH2, nOH, nCH, n))
    nstar = n if nstar is None else f64(nstar)
    ncol, ndep = T.shape
    assert n.shape == (ncol, model.nlev, ndep)
    out = [np.zeros((ncol, len(lam), ndep)) for _ in range(3)]
    P = lambda x: None if x is None else x.ctypes.data_as(dp)   # noqa: E731
    con = np.zeros((ncol, len(lam), 13, 2, ndep)) if contrib else None
    _lib.check(ctx.lib.rhb200_continuum_batch(ctx.h, C.byref(model.struct), len(lam), P(lam), ncol, ndep, P(T), P(ne),
                                              P(nHmin), P(nH2), P(nOH), P(nCH), P(n), P(nstar), *[P(o) for o in out],
                                              P(con)))
    return tuple(out) + ((con,) if contrib else ())
