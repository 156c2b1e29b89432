"""Flat (SoA-friendly) tables of Kurucz background lines for the B200 path.

The reference keeps lines as an array of pointer-rich ``RLK_Line`` structs
(rh/atom.h:156-167) filled by ``readKuruczLines`` (rh/kurucz.c:121-431) and
sorted by ``qsort(..., rlk_ascend)`` (rh/background.c:292-294).  The RH host
hands the same numbers to the device library as three dense tables:

* ``lines[nline, RL_NFIELD]``   one row per line, ascending ``lambda0``
* ``zq / zshift / zstrength``   concatenated Zeeman components (``RLKZeeman``,
  rh/kurucz.c:832-921); a row's ``RL_ZOFF``/``RL_NCOMP`` index into them
* ``elems[nelem, RE_NFIELD]``, ``pf[rows, npf]``, ``Tpf[npf]``  the element data
  that ``LTEpops_elem`` (rh/ltepops.c:116-159) and the ln U(T) interpolation in
  ``rlk_opacity`` (rh/kurucz.c:666) consume.

Field order is the C ABI's (include/rhb200.h).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

# line row fields (doubles) -- include/rhb200.h RHB200_RL_*
(RL_LAMBDA0, RL_GI, RL_GJ, RL_EI, RL_EJ, RL_BJI, RL_AJI, RL_BIJ, RL_GRAD, RL_GSTARK,
 RL_GVDW, RL_HFS_FRAC, RL_ISO_FRAC, RL_CROSS, RL_ALPHA, RL_POLARIZABLE, RL_VDWAALS,
 RL_ELEM, RL_STAGE, RL_ZOFF, RL_NCOMP) = range(21)
RL_NFIELD = 24
# element row fields
RE_WEIGHT, RE_ABUND, RE_NSTAGE, RE_PFROW, RE_IONPOT0 = range(5)
RE_MAXSTAGE = 12
RE_NFIELD = RE_IONPOT0 + RE_MAXSTAGE
VDW_UNSOLD, VDW_RIDDER, VDW_BARKLEM, VDW_KURUCZ = range(4)


@dataclass
class LineTable:
    lines: np.ndarray        # [nline, RL_NFIELD] float64
    zq: np.ndarray           # [ncomp] int32
    zshift: np.ndarray       # [ncomp] float64
    zstrength: np.ndarray    # [ncomp] float64
    elems: np.ndarray        # [nelem, RE_NFIELD] float64
    pf: np.ndarray           # [rows, npf] float64 (ln U)
    Tpf: np.ndarray          # [npf] float64
    vmicro_char: float = 5.0e3   # keyword VMICRO_CHAR [m/s] (readvalue: km/s -> m/s)

    @property
    def nline(self) -> int:
        return int(self.lines.shape[0])

    @property
    def nelem(self) -> int:
        return int(self.elems.shape[0])

    def validate(self) -> None:
        assert self.lines.ndim == 2 and self.lines.shape[1] == RL_NFIELD
        assert self.elems.ndim == 2 and self.elems.shape[1] == RE_NFIELD
        lam = self.lines[:, RL_LAMBDA0]
        if np.any(np.diff(lam) < 0):
            raise ValueError("line table must be sorted by ascending lambda0 "
                             "(background.c:292-294)")
        if self.nline:
            zend = (self.lines[:, RL_ZOFF] + self.lines[:, RL_NCOMP]).max()
            assert zend <= len(self.zq)
            assert self.lines[:, RL_ELEM].max() < self.nelem

    def to_npz_dict(self, prefix: str = "lt_") -> dict:
        return {prefix + k: getattr(self, k) for k in
                ("lines", "zq", "zshift", "zstrength", "elems", "pf", "Tpf")} | {
                prefix + "vmicro_char": np.float64(self.vmicro_char)}

    @classmethod
    def from_npz(cls, z, prefix: str = "lt_") -> "LineTable":
        lt = cls(lines=np.array(z[prefix + "lines"], np.float64),
                 zq=np.array(z[prefix + "zq"], np.int32),
                 zshift=np.array(z[prefix + "zshift"], np.float64),
                 zstrength=np.array(z[prefix + "zstrength"], np.float64),
                 elems=np.array(z[prefix + "elems"], np.float64),
                 pf=np.array(z[prefix + "pf"], np.float64),
                 Tpf=np.array(z[prefix + "Tpf"], np.float64),
                 vmicro_char=float(z[prefix + "vmicro_char"]))
        lt.validate()
        return lt
