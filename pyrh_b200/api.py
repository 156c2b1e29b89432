"""Host-side mirror of the reference's operator interface for the LTE Stokes path.

Reference call stack being replaced: ``pyrh.compute1d`` (pyrh.pyx:537-668)
-> ``rhf1d`` (rh/rhf1d/pyrh_compute1dray.c:112-389) -> Background/rlk_opacity -> Iterate ->
Formal -> Piece_Stokes_Bezier3_1D -> _solveray.  The RH host keeps everything that is not
on the hot path (input parsing, LTE/chemical equilibrium, continuum opacities, tau->height):
it hands this module the per-column arrays it already holds when ``Formal`` starts, and gets
the emergent Stokes spectra back.  Argument meaning and units follow ``compute1d``.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from .linelist import LineTable

KM_TO_M = 1.0E+03
CM_TO_M = 1.0E-02
_CUBE_CM_TO_M = (CM_TO_M * CM_TO_M) * CM_TO_M         # CUBE(CM_TO_M), rh.h:48

AT = dict(T=0, ne=1, vturb=2, vel=3, B=4, cos_gamma=5, cos_2chi=6, sin_2chi=7, nHtot=8, np=9, height=10)


def _dp(a: np.ndarray):
    return a.ctypes.data_as(_lib.dp)


def _vp(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


_cos = np.frompyfunc(math.cos, 1, 1)     # glibc libm, like the reference's Bproject()
_sin = np.frompyfunc(math.sin, 1, 1)


def atmos_rows_from_pyrh(atmosphere: np.ndarray, np_: np.ndarray, height: np.ndarray,
                         mu: float = 1.0) -> np.ndarray:
    """pyrh-unit atmosphere(s) ``[..., 9+, Ndep]`` (pyrh.pyx:621-625) -> device rows
    ``[..., RHB200_AT_NFIELD, Ndep]`` in SI, applying exactly the in-place conversion of
    pyrh_compute1dray.c:263-270 and ``Bproject`` for mu = 1 (rhf1d/project.c:52-58).
    ``np_`` [m^-3] and ``height`` [m] come from the RH host (LTE populations of hydrogen,
    convertScales)."""
    a = np.asarray(atmosphere, dtype=np.float64)
    if mu != 1.0:
        raise NotImplementedError("Bproject for mu != 1 (project.c:60-77) stays on the RH host: "
                                  "pass cos_gamma/cos_2chi/sin_2chi rows directly")
    out = np.empty(a.shape[:-2] + (_lib.AT_NFIELD, a.shape[-1]))
    out[..., AT["T"], :] = a[..., 1, :]
    out[..., AT["ne"], :] = a[..., 2, :] / _CUBE_CM_TO_M
    out[..., AT["vel"], :] = a[..., 3, :] * KM_TO_M
    out[..., AT["vturb"], :] = a[..., 4, :] * KM_TO_M
    out[..., AT["B"], :] = a[..., 5, :] / 1e4
    out[..., AT["cos_gamma"], :] = _cos(a[..., 6, :]).astype(np.float64)
    out[..., AT["cos_2chi"], :] = _cos(2.0 * a[..., 7, :]).astype(np.float64)
    out[..., AT["sin_2chi"], :] = _sin(2.0 * a[..., 7, :]).astype(np.float64)
    out[..., AT["nHtot"], :] = a[..., 8, :] / _CUBE_CM_TO_M
    out[..., AT["np"], :] = np_
    out[..., AT["height"], :] = height
    return out


def is_moving(atmosphere: np.ndarray, vmacro_tresh: float = 0.0) -> bool:
    """atmos.moving, pyrh_compute1dray.c:261-278 (per column)."""
    return bool(np.any(np.abs(np.asarray(atmosphere)[..., 3, :] * KM_TO_M) >= vmacro_tresh))


# enum S_interpol / S_interpol_stokes of the reference (inputs.h:26-27), keyed by the keyword.input spelling
S_INTERPOLATION = {"S_LINEAR": 0, "S_PARABOLIC": 1, "S_BEZIER3": 2}
S_INTERPOLATION_STOKES = {"DELO_PARABOLIC": 0, "DELO_BEZIER3": 1}


class Context:
    """One GPU context (``rhb200_open``): shared line tables + wavelength grid."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        if self.lib.rhb200_device_count() <= 0:
            raise _lib.RHB200Error("no CUDA device visible: pyrh_b200 has no CPU fallback")
        self.h = self.lib.rhb200_open(int(device))
        if not self.h:
            raise _lib.RHB200Error(self.lib.rhb200_last_error().decode())
        self.device = device
        self.nlambda = 0
        self.lt: LineTable | None = None

    def _staging(self, name, shape):
        """Page-locked result buffer kept with the context (grown on demand): D2H copies into pageable numpy memory
        run at a fraction of the PCIe rate and do not overlap the kernels of the next chunk."""
        cache = self.__dict__.setdefault("_stage", {})
        n = int(np.prod(shape))
        buf = cache.get(name)
        if buf is None or buf.size < n:
            if buf is not None:                                # grown: the old page-locked block goes back first
                self.lib.rhb200_host_free_pinned(C.c_void_p(buf.ctypes.data))
            buf = cache[name] = pinned_empty((max(n, 1),))
        return buf[:n].reshape(shape)

    def close(self):
        if getattr(self, "h", None):
            self.lib.rhb200_close(self.h)                      # synchronises the device first
            self.h = None
            for buf in self.__dict__.pop("_stage", {}).values():   # results were always handed out as copies
                self.lib.rhb200_host_free_pinned(C.c_void_p(buf.ctypes.data))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- tables ------------------------------------------------------------
    def set_lines(self, lt: LineTable, magneto_optical: bool = False, rlkscatter: bool = False):
        lt.validate()
        lines = np.ascontiguousarray(lt.lines, np.float64)
        zq = np.ascontiguousarray(lt.zq, np.int32)
        zs = np.ascontiguousarray(lt.zshift, np.float64)
        zt = np.ascontiguousarray(lt.zstrength, np.float64)
        el = np.ascontiguousarray(lt.elems, np.float64)
        pf = np.ascontiguousarray(lt.pf, np.float64)
        tp = np.ascontiguousarray(lt.Tpf, np.float64)
        self.lrf_npar = 0
        _lib.check(self.lib.rhb200_set_lines(self.h, lines.shape[0], _dp(lines), len(zq),
                                             zq.ctypes.data_as(_lib.ip), _dp(zs), _dp(zt),
                                             el.shape[0], _dp(el), pf.shape[0], pf.shape[1], _dp(pf),
                                             _dp(tp), float(lt.vmicro_char), int(magneto_optical),
                                             int(rlkscatter)))
        self.lt = lt
        self.nlambda = 0

    def set_model_lines(self, rows):
        """Lines of the PASSIVE model atoms that switch off Kurucz lines of the same element and stage inside their
        wing windows (kurucz.c:617-633): rows [n, 4] = element row, lower-level stage, lambda0 [nm], qwing."""
        rows = np.ascontiguousarray(rows, np.float64).reshape(-1, 4)
        _lib.check(self.lib.rhb200_set_model_lines(self.h, len(rows), _dp(rows)))

    def set_molecular_lines(self, mlines, molecules, zq=None, zshift=None, zstrength=None):
        """Lines of PASSIVE molecules for the fused path: ``mlines [n, 16]``, ``molecules [nmol, 16]`` (after
        set_continuum / set_chemistry, before set_wavelengths); polarizable rows index their MolZeeman components in
        ``zq / zshift / zstrength``."""
        ml = np.ascontiguousarray(mlines, np.float64).reshape(-1, 16)
        ms = np.ascontiguousarray(molecules, np.float64).reshape(-1, 16)
        zq = np.ascontiguousarray(zq if zq is not None else [], np.int32)
        zs = np.ascontiguousarray(zshift if zshift is not None else [], np.float64)
        zt = np.ascontiguousarray(zstrength if zstrength is not None else [], np.float64)
        _lib.check(self.lib.rhb200_set_molecular_lines_zeeman(self.h, len(ml), _dp(ml), len(ms), _dp(ms), len(zq),
                                                              zq.ctypes.data_as(_lib.ip), _dp(zs), _dp(zt)))

    def set_scatter(self, n_max_scatter=0, iter_limit=1.0e-2):
        """keywords N_MAX_SCATTER / ITER_LIMIT in LTE: scattering passes of the line-free wavelengths."""
        _lib.check(self.lib.rhb200_set_scatter(self.h, int(n_max_scatter), float(iter_limit)))

    def set_stokes_mode(self, mode="FULL_STOKES"):
        """keyword STOKES_MODE: FULL_STOKES or NO_STOKES (call before set_wavelengths)."""
        if mode not in ("FULL_STOKES", "NO_STOKES"):
            raise NotImplementedError(f"STOKES_MODE = {mode}: only FULL_STOKES and NO_STOKES are implemented")
        _lib.check(self.lib.rhb200_set_stokes_mode(self.h, int(mode == "FULL_STOKES")))

    def set_passive_lines(self, rows, c_shift, c_fraction):
        """Bound-bound lines of the PASSIVE model atoms for the fused path (passive_bb): rows [n, RHB200_PL_NFIELD]."""
        rows = np.ascontiguousarray(rows, np.float64).reshape(-1, 28)
        cs, cf = np.ascontiguousarray(c_shift, np.float64), np.ascontiguousarray(c_fraction, np.float64)
        _lib.check(self.lib.rhb200_set_passive_lines(self.h, len(rows), _dp(rows), len(cs), _dp(cs), _dp(cf)))

    def set_wavelengths(self, lam):
        lam = np.ascontiguousarray(lam, np.float64)
        _lib.check(self.lib.rhb200_set_wavelengths(self.h, len(lam), _dp(lam)))
        self.nlambda = len(lam)
        self.lam = lam

    def line_windows(self):
        first = np.zeros(self.nlambda, np.int32)
        count = np.zeros(self.nlambda, np.int32)
        n = C.c_int()
        _lib.check(self.lib.rhb200_get_line_windows(self.h, first.ctypes.data_as(_lib.ip),
                                                    count.ctypes.data_as(_lib.ip), None, 0, C.byref(n)))
        idx = np.zeros(max(n.value, 1), np.int32)
        _lib.check(self.lib.rhb200_get_line_windows(self.h, None, None, idx.ctypes.data_as(_lib.ip),
                                                    len(idx), C.byref(n)))
        return first, count, idx[: n.value]

    # -- the hot path --------------------------------------------------------
    def lte_stokes_batch(self, atmos_rows, chi_ai, eta_ai, mu=1.0, moving=True,
                         bc_top=_lib.BC_ZERO, bc_bottom=_lib.BC_THERMALIZED, out=None):
        """atmos_rows [ncol, AT_NFIELD, ndep], chi_ai/eta_ai [ncol, nlambda, ndep] (host arrays)
        -> stokes [ncol, 4, nlambda]."""
        at = np.ascontiguousarray(atmos_rows, np.float64)
        ca = np.ascontiguousarray(chi_ai, np.float64)
        ea = np.ascontiguousarray(eta_ai, np.float64)
        ncol, nf, ndep = at.shape
        assert nf == _lib.AT_NFIELD
        assert ca.shape == (ncol, self.nlambda, ndep) and ea.shape == ca.shape
        if out is None:
            out = np.empty((ncol, 4, self.nlambda))
        _lib.check(self.lib.rhb200_lte_stokes_batch(self.h, ncol, ndep, float(mu), int(moving),
                                                    int(bc_top), int(bc_bottom), _vp(at), _vp(ca),
                                                    _vp(ea), _vp(out)))
        return out

    def set_continuum(self, model, abundance):
        """Background continuum on the device for the wavelengths set before (rhb200_set_continuum).
        ``model``: pyrh_b200.continuum.ContinuumModel; ``abundance`` [natom] = atom->abundance."""
        self._cont_model = model                       # the struct points into its arrays
        ab = np.ascontiguousarray(abundance, np.float64)
        _lib.check(self.lib.rhb200_set_continuum(self.h, C.byref(model.struct), _dp(ab)))

    def lte_stokes_batch_pops(self, atmos_rows, chem, mu=1.0, moving=True, bc_top=_lib.BC_ZERO,
                              bc_bottom=_lib.BC_THERMALIZED, out=None):
        """``lte_stokes_batch`` with LTE populations, continuum, line opacity and formal solution all on the
        device: chem [ncol, natom+4, ndep] = ChemicalEquilibrium's population factor per model atom, nHmin, nH2,
        nOH, nCH."""
        at = np.ascontiguousarray(atmos_rows, np.float64)
        ch = np.ascontiguousarray(chem, np.float64)
        ncol, _, ndep = at.shape
        nl = self.nlambda
        st = np.empty((ncol, 4, nl)) if out is None else out
        _lib.check(self.lib.rhb200_lte_stokes_batch_pops(self.h, ncol, ndep, float(mu), int(moving), int(bc_top),
                                                         int(bc_bottom), _vp(at), _vp(ch), _vp(st)))
        return st

    def set_chemistry(self, nucleus_atom, mol):
        """ChemicalEquilibrium on the device: nucleus_atom [nnuclei] = model-atom index of each nucleus, mol
        [nmol, 32] = per-molecule fit records (include/rhb200.h)."""
        na = np.ascontiguousarray(nucleus_atom, np.int32)
        mo = np.ascontiguousarray(mol, np.float64)
        _lib.check(self.lib.rhb200_set_chemistry(self.h, len(na), na.ctypes.data_as(_lib.ip), mo.shape[0], _dp(mo)))
        self._natom_lev = None

    def chemistry(self, atmos_rows, natom, nlev):
        """(chem [ncol, natom+4, ndep], pops [ncol, nlev, ndep]) from LTEpops + ChemicalEquilibrium on the device."""
        at = np.ascontiguousarray(atmos_rows, np.float64)
        ncol, _, ndep = at.shape
        chem = np.zeros((ncol, natom + 4, ndep))
        pops = np.zeros((ncol, nlev, ndep))
        _lib.check(self.lib.rhb200_chemistry_batch(self.h, ncol, ndep, _dp(at), _dp(chem), _dp(pops)))
        return chem, pops

    def lte_stokes_batch_atmos(self, atmos_rows, mu=1.0, moving=True, bc_top=_lib.BC_ZERO,
                               bc_bottom=_lib.BC_THERMALIZED, out=None):
        """The whole LTE column from the atmosphere rows alone (set_lines, set_wavelengths, set_continuum and
        set_chemistry done before)."""
        at = np.ascontiguousarray(atmos_rows, np.float64)
        ncol, _, ndep = at.shape
        st = np.empty((ncol, 4, self.nlambda)) if out is None else out
        _lib.check(self.lib.rhb200_lte_stokes_batch_atmos(self.h, ncol, ndep, float(mu), int(moving), int(bc_top),
                                                          int(bc_bottom), _vp(at), _vp(st)))
        return st

    def compute1d_batch(self, atmosphere, mu=1.0, atm_scale=0, lambda_ref=500.0, wght_per_H=0.0, vmacro_tresh=0.0,
                        bc_top=_lib.BC_ZERO, bc_bottom=_lib.BC_THERMALIZED, out=None, get_scales=False,
                        keep_lambda_ref=False):
        """``pyrh.compute1d`` (pyrh.pyx:537-668) for a batch of columns in LTE: ``atmosphere[ncol, 9+, ndep]`` in
        pyrh's own units and scales (``atm_scale`` 0 log tau500, 1 log column mass, 2 height km) -> Stokes spectra
        ``[ncol, 4, nlambda]``; nothing is derived on the host.  The context's wavelength grid must be
        ``spectrum.lambda``, i.e. contain ``lambda_ref``; its column is dropped from the result like ``_solveray``
        does (pyrh_solveray.c:130-150) unless ``keep_lambda_ref``.  ``vmacro_tresh`` in km/s like the keyword."""
        a = np.ascontiguousarray(atmosphere, np.float64)
        if a.ndim != 3 or a.shape[1] < 9:
            raise ValueError("atmosphere must be [ncol, >=9, ndep] (pyrh.pyx:621-625)")
        ncol, nrow, ndep = a.shape
        lam = np.asarray(self.lam)
        hit = np.nonzero(lam == lambda_ref)[0]
        if len(hit) != 1:
            raise ValueError("the wavelength grid must contain lambda_ref once (sortlambda.c adds it to spectrum.lambda)")
        iref = int(hit[0])
        st = self._staging("stokes", (ncol, 4, self.nlambda)) if out is None else out
        sc = np.empty((ncol, 3, ndep)) if get_scales else None
        _lib.check(self.lib.rhb200_compute1d_batch(self.h, ncol, ndep, nrow, float(mu), int(atm_scale), _vp(a), iref,
                                                   float(wght_per_H), float(vmacro_tresh) * KM_TO_M, int(bc_top),
                                                   int(bc_bottom), _vp(st), _vp(sc) if sc is not None else None))
        if keep_lambda_ref:
            res = st if out is not None else st.copy()           # never hand out the context's staging buffer
        else:
            res = np.delete(st, iref, axis=2)
        return (res, sc) if get_scales else res

    def set_loggf_rf(self, line_rows):
        """``get_atomic_rfs``: rows of the line table (sorted by wavelength) whose log gf the analytic response
        function is taken for; parameter p <-> ``line_rows[p]``.  ``set_lines`` clears it."""
        rows = np.ascontiguousarray(line_rows, np.int32)
        _lib.check(self.lib.rhb200_set_loggf_rf(self.h, len(rows), rows.ctypes.data_as(_lib.ip)))
        self.lrf_npar = len(rows)

    def compute1d_rf_batch(self, atmosphere, mu=1.0, atm_scale=0, lambda_ref=500.0, wght_per_H=0.0, vmacro_tresh=0.0,
                           bc_top=_lib.BC_ZERO, bc_bottom=_lib.BC_THERMALIZED, keep_lambda_ref=False):
        """``compute1d_batch`` with ``get_atomic_rfs``: ``(stokes [ncol, 4, nlambda], rfs [ncol, nlambda, npar])`` --
        ``mySpectrum.rfs`` (pyrh_solveray.c:144-147) for the lines registered with ``set_loggf_rf``."""
        a = np.ascontiguousarray(atmosphere, np.float64)
        if a.ndim != 3 or a.shape[1] < 9:
            raise ValueError("atmosphere must be [ncol, >=9, ndep] (pyrh.pyx:621-625)")
        ncol, nrow, ndep = a.shape
        hit = np.nonzero(np.asarray(self.lam) == lambda_ref)[0]
        if len(hit) != 1:
            raise ValueError("the wavelength grid must contain lambda_ref once (sortlambda.c adds it to spectrum.lambda)")
        iref = int(hit[0])
        npar = getattr(self, "lrf_npar", 0)
        st = self._staging("stokes", (ncol, 4, self.nlambda))
        rf = self._staging("rfs", (ncol, self.nlambda, max(npar, 1)))
        _lib.check(self.lib.rhb200_compute1d_rf_batch(self.h, ncol, ndep, nrow, float(mu), int(atm_scale), _vp(a), iref,
                                                      float(wght_per_H), float(vmacro_tresh) * KM_TO_M, int(bc_top),
                                                      int(bc_bottom), _vp(st), None, _vp(rf)))
        rf = rf[:, :, :npar]
        if keep_lambda_ref:
            return st.copy(), rf.copy()
        return np.delete(st, iref, axis=2), np.delete(rf, iref, axis=1)

    def set_elements(self, elems, pf, Tpf):
        """atmos.elements[] for Solve_ne: ``elems[nelem, RE_NFIELD]`` (hydrogen first), ``pf[rows, npf]`` = ln U."""
        elems = np.ascontiguousarray(elems, np.float64)
        pf = np.ascontiguousarray(pf, np.float64)
        Tpf = np.ascontiguousarray(Tpf, np.float64)
        _lib.check(self.lib.rhb200_set_elements(self.h, len(elems), _dp(elems), len(pf), pf.shape[1], _dp(pf), _dp(Tpf)))

    def solve_ne(self, T, nHtot, ne0=None):
        """Solve_ne (solvene.c:55-140) at every point of ``T`` / ``nHtot`` [K, m^-3] -> ne [m^-3]; ``ne0`` = starting
        guess (else from scratch: ionisation of hydrogen alone)."""
        T = np.ascontiguousarray(T, np.float64)
        nH = np.ascontiguousarray(nHtot, np.float64)
        ne = np.zeros(T.shape) if ne0 is None else np.array(ne0, np.float64, copy=True)
        _lib.check(self.lib.rhb200_solve_ne_batch(self.h, T.size, _dp(T), _dp(nH), _dp(ne), int(ne0 is None)))
        return ne

    def hse_batch(self, scale, T, pg_top, atm_scale=0, wght_per_H=0.0, total_abund=0.0, gravity=None):
        """``pyrh.hse`` for a batch: ``scale``, ``T`` [ncol, ndep], ``pg_top`` [ncol] (Pa) -> ne, nHtot, rho, pg (SI)."""
        scale = np.ascontiguousarray(scale, np.float64)
        T = np.ascontiguousarray(T, np.float64)
        ncol, ndep = T.shape
        pg_top = np.ascontiguousarray(np.broadcast_to(np.asarray(pg_top, np.float64), (ncol,)))
        g = math.exp(2.30258509299404568402 * 4.4) * 1.0E-02 if gravity is None else gravity    # multiatmos.c:69,82
        out = [np.empty((ncol, ndep)) for _ in range(4)]
        _lib.check(self.lib.rhb200_hse_batch(self.h, ncol, ndep, int(atm_scale), _dp(scale), _dp(T), _dp(pg_top),
                                             float(wght_per_H), float(total_abund), float(g), *[_dp(o) for o in out]))
        return tuple(out)

    def set_gravity(self, total_abund, gravity=None):
        """atmos.totalAbund / atmos.gravity for the column-mass row of ``get_scales=True`` on a height grid."""
        g = math.exp(2.30258509299404568402 * 4.4) * 1.0E-02 if gravity is None else gravity    # multiatmos.c:69,82
        _lib.check(self.lib.rhb200_set_gravity(self.h, float(total_abund), float(g)))

    def get_scales_batch(self, atmosphere, atm_scale=0, lambda_ref=500.0, wght_per_H=0.0, total_abund=0.0,
                         gravity=None, vmacro_tresh=0.0):
        """``pyrh.get_scales`` for a batch: ``[ncol, 3, ndep]`` = height [m], tau_ref, column mass [kg m^-2]."""
        a = np.ascontiguousarray(atmosphere, np.float64)
        ncol, nrow, ndep = a.shape
        g = math.exp(2.30258509299404568402 * 4.4) * 1.0E-02 if gravity is None else gravity    # multiatmos.c:69,82
        sc = np.empty((ncol, 3, ndep))
        _lib.check(self.lib.rhb200_get_scales_batch(self.h, ncol, ndep, nrow, int(atm_scale), _vp(a), self._iref(lambda_ref),
                                                    float(wght_per_H), float(total_abund), float(g),
                                                    float(vmacro_tresh) * KM_TO_M, _vp(sc)))
        return sc

    def _iref(self, lambda_ref):
        hit = np.nonzero(np.asarray(self.lam) == lambda_ref)[0]
        if len(hit) != 1:
            raise ValueError("the wavelength grid must contain lambda_ref once (sortlambda.c adds it to spectrum.lambda)")
        return int(hit[0])

    def rf_fd_batch(self, atmosphere, par_rows, par_delta, mu=1.0, atm_scale=0, lambda_ref=500.0, wght_per_H=0.0,
                    vmacro_tresh=0.0, bc_top=_lib.BC_ZERO, bc_bottom=_lib.BC_THERMALIZED, out=None,
                    keep_lambda_ref=False, depths=None):
        """Centred finite-difference response functions ``[ncol, npar, ndep, 4, nlambda]`` of the spectrum
        ``compute1d_batch`` returns, to ``atmosphere`` row ``par_rows[p]`` at every depth (step ``par_delta[p]``).
        The perturbed columns are generated and differenced on the device.  ``depths``: only these depth indices (the
        nodes of an inversion) -> ``[ncol, npar, len(depths), 4, nlambda]``."""
        a = np.ascontiguousarray(atmosphere, np.float64)
        ncol, nrow, ndep = a.shape
        rows = np.ascontiguousarray(par_rows, np.int32)
        delta = np.ascontiguousarray(par_delta, np.float64)
        iref = self._iref(lambda_ref)
        sel = None if depths is None else np.ascontiguousarray(depths, np.int32)
        nsel = ndep if sel is None else len(sel)
        rf = np.empty((ncol, len(rows), nsel, 4, self.nlambda)) if out is None else out
        _lib.check(self.lib.rhb200_rf_fd_depths_batch(self.h, ncol, ndep, nrow, float(mu), int(atm_scale), _vp(a), iref,
                                                      float(wght_per_H), float(vmacro_tresh) * KM_TO_M, int(bc_top),
                                                      int(bc_bottom), len(rows), rows.ctypes.data_as(C.POINTER(C.c_int)),
                                                      _dp(delta), nsel,
                                                      None if sel is None else sel.ctypes.data_as(C.POINTER(C.c_int)), _vp(rf)))
        return rf if keep_lambda_ref else np.delete(rf, iref, axis=4)

    def lte_stokes_batch_dev(self, ncol, ndep, d_atmos, d_chi_ai, d_eta_ai, d_stokes, mu=1.0,
                             moving=True, bc_top=_lib.BC_ZERO, bc_bottom=_lib.BC_THERMALIZED):
        _lib.check(self.lib.rhb200_lte_stokes_batch_dev(self.h, int(ncol), int(ndep), float(mu),
                                                        int(moving), int(bc_top), int(bc_bottom),
                                                        d_atmos, d_chi_ai, d_eta_ai, d_stokes))

    # -- function-level entry points -------------------------------------------
    def ltepops_elem(self, atmos_rows):
        at = np.ascontiguousarray(atmos_rows, np.float64)
        ncol, _, ndep = at.shape
        n = np.zeros((ncol, self.lt.nelem, _lib.RE_MAXSTAGE, ndep))
        _lib.check(self.lib.rhb200_ltepops_elem_batch(self.h, ncol, ndep, _dp(at), _dp(n)))
        return n

    def rlk_opacity(self, atmos_rows, mu=1.0, moving=True, to_obs=True):
        at = np.ascontiguousarray(atmos_rows, np.float64)
        ncol, _, ndep = at.shape
        chi = np.zeros((ncol, self.nlambda, 4, ndep))
        eta = np.zeros_like(chi)
        flags = np.zeros(self.nlambda, np.int32)
        _lib.check(self.lib.rhb200_rlk_opacity_batch(self.h, ncol, ndep, float(mu), int(moving),
                                                     int(to_obs), _dp(at), _dp(chi), _dp(eta),
                                                     flags.ctypes.data_as(_lib.ip)))
        return chi, eta, flags

    def set_solvers(self, s_interpolation="S_BEZIER3", s_interpolation_stokes="DELO_BEZIER3"):
        """keyword.input S_INTERPOLATION / S_INTERPOLATION_STOKES (readvalue.c:366-404)."""
        _lib.check(self.lib.rhb200_set_solvers(self.h, S_INTERPOLATION[s_interpolation],
                                               S_INTERPOLATION_STOKES[s_interpolation_stokes]))

    def molecular_opacity(self, atmos_rows, mol, mlines, zq, zshift, zstrength, lam, vmicro_char, mu=1.0,
                          moving=True, to_obs=True):
        """MolecularOpacity() for every column and wavelength: mol [ncol, nmol, 3, ndep] (n, pf, vbroad),
        mlines [nmline, ML_NFIELD].  Returns (chi, eta [ncol, nlambda, 4, ndep], flags [nlambda])."""
        at = np.ascontiguousarray(atmos_rows, np.float64)
        mol = np.ascontiguousarray(mol, np.float64)
        ml = np.ascontiguousarray(mlines, np.float64)
        zq = np.ascontiguousarray(zq, np.int32)
        zs, zt = np.ascontiguousarray(zshift, np.float64), np.ascontiguousarray(zstrength, np.float64)
        lam = np.ascontiguousarray(lam, np.float64)
        ncol, _, ndep = at.shape
        chi = np.zeros((ncol, len(lam), 4, ndep))
        eta = np.zeros_like(chi)
        flags = np.zeros(len(lam), np.int32)
        _lib.check(self.lib.rhb200_molecular_opacity_batch(
            self.h, ncol, ndep, float(mu), int(moving), int(to_obs), mol.shape[1], ml.shape[0], _dp(ml), len(zq),
            zq.ctypes.data_as(_lib.ip), _dp(zs), _dp(zt), float(vmicro_char), len(lam), _dp(lam), _dp(at), _dp(mol),
            _dp(chi), _dp(eta), flags.ctypes.data_as(_lib.ip)))
        return chi, eta, flags

    def passive_bb(self, atmos_rows, pcol, plines, c_shift, c_fraction, lam, vmicro_char, mu=1.0, moving=True,
                   to_obs=True):
        """passive_bb() for every column and wavelength: pcol [ncol, nline, 4, ndep] (n_i, n_j, vbroad, adamp),
        plines [nline, PB_NFIELD].  Returns (chi, eta [ncol, nlambda, ndep], flags [nlambda])."""
        at = np.ascontiguousarray(atmos_rows, np.float64)
        pc = np.ascontiguousarray(pcol, np.float64)
        pl = np.ascontiguousarray(plines, np.float64)
        cs, cf = np.ascontiguousarray(c_shift, np.float64), np.ascontiguousarray(c_fraction, np.float64)
        lam = np.ascontiguousarray(lam, np.float64)
        ncol, _, ndep = at.shape
        chi = np.zeros((ncol, len(lam), ndep))
        eta = np.zeros_like(chi)
        flags = np.zeros(len(lam), np.int32)
        _lib.check(self.lib.rhb200_passive_bb_batch(
            self.h, ncol, ndep, float(mu), int(moving), int(to_obs), pl.shape[0], _dp(pl), len(cs), _dp(cs), _dp(cf),
            float(vmicro_char), len(lam), _dp(lam), _dp(at), _dp(pc), _dp(chi), _dp(eta),
            flags.ctypes.data_as(_lib.ip)))
        return chi, eta, flags

    def stokes_bezier3(self, ray_col, ray_lambda, height, T, chi, S, chiQUV, mu=1.0, to_obs=True,
                       bc_top=_lib.BC_ZERO, bc_bottom=_lib.BC_THERMALIZED, want_psi=False,
                       solver="DELO_BEZIER3"):
        rc = np.ascontiguousarray(ray_col, np.int32)
        rl = np.ascontiguousarray(ray_lambda, np.float64)
        h = np.ascontiguousarray(np.atleast_2d(height), np.float64)
        t = np.ascontiguousarray(np.atleast_2d(T), np.float64)
        chi = np.ascontiguousarray(chi, np.float64)
        S = np.ascontiguousarray(S, np.float64)
        q = np.ascontiguousarray(chiQUV, np.float64)
        nray, ndep = chi.shape
        I = np.zeros((nray, 4, ndep))
        Psi = np.zeros((nray, ndep)) if want_psi else None
        _lib.check(self.lib.rhb200_stokes_ray_batch(
            self.h, S_INTERPOLATION_STOKES[solver], nray, h.shape[0], ndep, float(mu), int(to_obs), int(bc_top), int(bc_bottom),
            rc.ctypes.data_as(_lib.ip), _dp(rl), _dp(h), _dp(t), _dp(chi), _dp(S), _dp(q), _dp(I),
            _dp(Psi) if want_psi else None))
        return (I, Psi) if want_psi else I

    def bezier3(self, ray_col, ray_lambda, height, T, chi, S, mu=1.0, to_obs=True,
                bc_top=_lib.BC_ZERO, bc_bottom=_lib.BC_THERMALIZED, want_psi=False, solver="S_BEZIER3"):
        rc = np.ascontiguousarray(ray_col, np.int32)
        rl = np.ascontiguousarray(ray_lambda, np.float64)
        h = np.ascontiguousarray(np.atleast_2d(height), np.float64)
        t = np.ascontiguousarray(np.atleast_2d(T), np.float64)
        chi = np.ascontiguousarray(chi, np.float64)
        S = np.ascontiguousarray(S, np.float64)
        nray, ndep = chi.shape
        I = np.zeros((nray, ndep))
        Psi = np.zeros((nray, ndep)) if want_psi else None
        _lib.check(self.lib.rhb200_scalar_ray_batch(
            self.h, S_INTERPOLATION[solver], nray, h.shape[0], ndep, float(mu), int(to_obs), int(bc_top), int(bc_bottom),
            rc.ctypes.data_as(_lib.ip), _dp(rl), _dp(h), _dp(t), _dp(chi), _dp(S), _dp(I),
            _dp(Psi) if want_psi else None))
        return (I, Psi) if want_psi else I

    def bezier3_rf(self, ray_col, ray_lambda, height, T, chi_dn, S_dn, chi_up, S_up, dchi, deta, mu=1.0,
                   bc_top=_lib.BC_ZERO, bc_bottom=_lib.BC_THERMALIZED):
        """Down-ray + up-ray with the analytic log gf response function (get_atomic_rfs):
        dchi, deta [nray, ndep, npar].  Returns (I [nray, ndep], dI [nray, ndep, npar])."""
        rc = np.ascontiguousarray(ray_col, np.int32)
        rl = np.ascontiguousarray(ray_lambda, np.float64)
        h = np.ascontiguousarray(np.atleast_2d(height), np.float64)
        t = np.ascontiguousarray(np.atleast_2d(T), np.float64)
        a = [np.ascontiguousarray(x, np.float64) for x in (chi_dn, S_dn, chi_up, S_up, dchi, deta)]
        nray, ndep = a[0].shape
        npar = a[4].shape[2]
        I = np.zeros((nray, ndep))
        dI = np.zeros((nray, ndep, npar))
        _lib.check(self.lib.rhb200_bezier3_rf_batch(
            self.h, nray, h.shape[0], ndep, float(mu), int(bc_top), int(bc_bottom),
            rc.ctypes.data_as(_lib.ip), _dp(rl), _dp(h), _dp(t), *[_dp(x) for x in a[:4]], int(npar),
            _dp(a[4]), _dp(a[5]), _dp(I), _dp(dI)))
        return I, dI

    def feautrier(self, ray_col, ray_lambda, height, T, chi, S, mu=1.0, bc_top=_lib.BC_ZERO,
                  bc_bottom=_lib.BC_THERMALIZED, want_psi=False):
        rc = np.ascontiguousarray(ray_col, np.int32)
        rl = np.ascontiguousarray(ray_lambda, np.float64)
        h = np.ascontiguousarray(np.atleast_2d(height), np.float64)
        t = np.ascontiguousarray(np.atleast_2d(T), np.float64)
        chi = np.ascontiguousarray(chi, np.float64)
        S = np.ascontiguousarray(S, np.float64)
        nray, ndep = chi.shape
        P = np.zeros((nray, ndep))
        Iem = np.zeros(nray)
        Psi = np.zeros((nray, ndep)) if want_psi else None
        _lib.check(self.lib.rhb200_feautrier_batch(
            self.h, nray, h.shape[0], ndep, float(mu), int(bc_top), int(bc_bottom),
            rc.ctypes.data_as(_lib.ip), _dp(rl), _dp(h), _dp(t), _dp(chi), _dp(S), _dp(P),
            _dp(Psi) if want_psi else None, _dp(Iem)))
        return (P, Psi, Iem) if want_psi else (P, Iem)

    def voigt(self, a, v):
        a = np.ascontiguousarray(a, np.float64)
        v = np.ascontiguousarray(v, np.float64)
        H, F = np.zeros_like(a), np.zeros_like(a)
        reg = np.zeros(a.shape, np.int32)
        _lib.check(self.lib.rhb200_voigt_humlicek(self.h, a.size, _dp(a), _dp(v), _dp(H), _dp(F),
                                                  reg.ctypes.data_as(_lib.ip)))
        return H, F, reg

    def voigt_armstrong(self, a, v):
        a = np.ascontiguousarray(a, np.float64)
        v = np.ascontiguousarray(v, np.float64)
        H = np.zeros_like(a)
        reg = np.zeros(a.shape, np.int32)
        _lib.check(self.lib.rhb200_voigt_armstrong(self.h, a.size, _dp(a), _dp(v), _dp(H),
                                                   reg.ctypes.data_as(_lib.ip)))
        return H, reg

    def math_probe(self, func: str, x, y=None):
        code = dict(exp=0, sin=1, cos=2, pow=3, div_recip=4, div=5, atan=6, log=7, log10=8)[func]
        x = np.ascontiguousarray(x, np.float64)
        out = np.zeros_like(x)
        yy = np.ascontiguousarray(y, np.float64) if y is not None else None
        _lib.check(self.lib.rhb200_math_probe(self.h, x.size, code, _dp(x),
                                              _dp(yy) if yy is not None else None, _dp(out)))
        return out

    # -- device memory / instrumentation ------------------------------------------
    def dev_alloc(self, nbytes: int) -> C.c_void_p:
        p = C.c_void_p()
        _lib.check(self.lib.rhb200_dev_alloc(self.h, int(nbytes), C.byref(p)))
        return p

    def dev_free(self, p):
        _lib.check(self.lib.rhb200_dev_free(self.h, p))

    def h2d(self, dptr, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        _lib.check(self.lib.rhb200_memcpy_h2d(self.h, dptr, _vp(arr), arr.nbytes))

    def d2h(self, arr: np.ndarray, dptr):
        _lib.check(self.lib.rhb200_memcpy_d2h(self.h, _vp(arr), dptr, arr.nbytes))

    def synchronize(self):
        _lib.check(self.lib.rhb200_synchronize(self.h))

    def flush_l2(self):
        _lib.check(self.lib.rhb200_flush_l2(self.h))

    def timing(self, on=True):
        _lib.check(self.lib.rhb200_timing_enable(self.h, int(on)))
        _lib.check(self.lib.rhb200_timing_reset(self.h))

    def timing_get(self):
        out = {}
        for name, k in (("prep", 0), ("opacity", 1), ("delo", 2), ("bezier", 3), ("other", 4), ("nlte_gamma", 5),
                        ("nlte_J", 6), ("statequil", 7), ("ng", 8)):
            ms, n = C.c_double(), C.c_long()
            _lib.check(self.lib.rhb200_timing_get(self.h, k, C.byref(ms), C.byref(n)))
            out[name] = (ms.value, n.value)
        return out

    def timer_begin(self):
        _lib.check(self.lib.rhb200_timer_begin(self.h))

    def timer_end(self) -> float:
        ms = C.c_double()
        _lib.check(self.lib.rhb200_timer_end(self.h, C.byref(ms)))
        return ms.value

    def fp64_peak(self):
        a, b = C.c_double(), C.c_double()
        _lib.check(self.lib.rhb200_fp64_peak(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """numpy view over page-locked host memory (cudaHostAlloc) for overlapped H2D/D2H."""
    lib = _lib.load()
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    _lib.check(lib.rhb200_host_alloc_pinned(n, C.byref(p)))
    buf = (C.c_char * n).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    return arr
