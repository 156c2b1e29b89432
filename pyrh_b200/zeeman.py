"""Host-side Zeeman patterns (once per line list), thin ctypes mirror of the C ABI.

Reference: RLKdeterminate / RLKZeeman (rh/kurucz.c:925-969, 832-921), determinate / Zeeman / Lande
(rh/zeeman.c:37-85, 186-281, 139-146).  The arithmetic lives in librhb200.so (rhb200_zeeman.cu).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

GL_NOT_GIVEN = -99 * 1.0e-3          # the Kurucz list's "-99" Lande entry after the reference's MILLI scaling
MAXCOMP = 1024


def _err(rc: int):
    raise _lib.RHB200Error(f"librhb200 error {rc}: {_lib.load().rhb200_last_error().decode()}")


def lande(S: float, L: int, J: float) -> float:
    return float(_lib.load().rhb200_lande(float(S), int(L), float(J)))


def rlk_determinate(labeli: str, labelj: str):
    """(determined, Si, Li, Sj, Lj) from two Kurucz term labels."""
    S = np.zeros(2)
    L = np.zeros(2, np.int32)
    rc = _lib.load().rhb200_rlk_determinate(labeli.encode(), labelj.encode(),
                                            S[0:].ctypes.data_as(_lib.dp), L[0:].ctypes.data_as(_lib.ip),
                                            S[1:].ctypes.data_as(_lib.dp), L[1:].ctypes.data_as(_lib.ip))
    if rc < 0:
        _err(rc)
    return bool(rc), float(S[0]), int(L[0]), float(S[1]), int(L[1])


def _pattern(call):
    q = np.zeros(MAXCOMP, np.int32)
    sh = np.zeros(MAXCOMP)
    st = np.zeros(MAXCOMP)
    nc = call(MAXCOMP, q.ctypes.data_as(_lib.ip), sh.ctypes.data_as(_lib.dp), st.ctypes.data_as(_lib.dp))
    if nc < 0:
        _err(nc)
    if nc > MAXCOMP:
        raise _lib.RHB200Error(f"{nc} Zeeman components exceed MAXCOMP")
    return q[:nc].copy(), sh[:nc].copy(), st[:nc].copy()


def rlk_zeeman(gi, gj, Si, Li, Sj, Lj, gL_i=GL_NOT_GIVEN, gL_j=GL_NOT_GIVEN, LS_Lande=True):
    """(q, shift, strength) of a Kurucz line (RLKZeeman)."""
    lib = _lib.load()
    return _pattern(lambda cap, q, sh, st: lib.rhb200_rlk_zeeman(
        float(gi), float(gj), float(Si), int(Li), float(Sj), int(Lj), float(gL_i), float(gL_j),
        int(bool(LS_Lande)), cap, q, sh, st))


def determinate(label: str, g: float):
    """(determined, n, S, L, J) of a model-atom level label."""
    n = C.c_int(0); L = C.c_int(0); S = C.c_double(0); J = C.c_double(0)
    rc = _lib.load().rhb200_determinate(label.encode(), float(g), C.byref(n), C.byref(S), C.byref(L), C.byref(J))
    if rc < 0:
        _err(rc)
    return bool(rc), n.value, S.value, L.value, J.value


def zeeman(label_i: str, g_i: float, label_j: str, g_j: float, g_Lande_eff: float = 0.0):
    """(q, shift, strength) of a model-atom line (Zeeman)."""
    lib = _lib.load()
    return _pattern(lambda cap, q, sh, st: lib.rhb200_zeeman(
        label_i.encode(), float(g_i), label_j.encode(), float(g_j), float(g_Lande_eff), cap, q, sh, st))
