"""Synthetic perturbed-FAL-C atmospheres (BASELINE.json configs 2/3/5).

Recipe: SURVEY.md section 8(d).  Base model = the reference's FAL-C table
(tests/falc.dat, 57 depths) already converted to pyrh rows by the reference's
own ``spinor2multi`` (tests/test_compute1d.py:6-32) and stored in the golden
fixture; it is resampled to ``ndep`` points uniform in log tau_500 and
perturbed per column with ``numpy.random.default_rng(seed0 + column)``.
"""
from __future__ import annotations

import numpy as np

SEED0 = 20261017
_KERNELS: dict = {}


def _smooth_noise(rng, logtau, sigma_dex=0.5):
    """Unit-variance Gaussian noise smoothed with a sigma = 0.5 dex Gaussian kernel."""
    x = rng.standard_normal(logtau.size)
    key = (logtau.tobytes(), sigma_dex)
    if key not in _KERNELS:
        d = (logtau[:, None] - logtau[None, :]) / sigma_dex
        w = np.exp(-0.5 * d * d)
        w /= w.sum(axis=1, keepdims=True)
        _KERNELS[key] = (w, np.sqrt((w * w).sum(axis=1)))
    w, norm = _KERNELS[key]
    y = w @ x
    return y / norm


def resample_falc(base: np.ndarray, ndep: int = 70, lo: float = -6.0, hi: float = 1.4) -> np.ndarray:
    """base: pyrh rows [9, 57] with row 0 = log10 tau500.  Cubic in log tau for T,
    linear in log for ne and nH."""
    from scipy.interpolate import CubicSpline
    lt0 = base[0]
    lt = np.linspace(lo, hi, ndep)
    out = np.zeros((9, ndep))
    out[0] = lt
    out[1] = CubicSpline(lt0, base[1])(lt)
    out[2] = np.exp(np.interp(lt, lt0, np.log(base[2])))
    out[8] = np.exp(np.interp(lt, lt0, np.log(base[8])))
    out[4] = np.interp(lt, lt0, base[4])
    return out


def perturbed_column(base70: np.ndarray, column: int, seed0: int = SEED0) -> np.ndarray:
    rng = np.random.default_rng(seed0 + column)
    lt = base70[0]
    g = [_smooth_noise(rng, lt) for _ in range(5)]
    atm = base70.copy()
    T0 = base70[1]
    atm[1] = T0 * (1.0 + 0.03 * g[0])
    atm[3] = 1.5 * g[1]
    atm[4] = np.clip(1.0 + 0.5 * g[2], 0.2, 3.0)
    atm[5] = np.maximum(rng.uniform(0.0, 2500.0) * (1.0 + 0.2 * g[3]), 0.0)
    atm[6] = rng.uniform(0.0, np.pi) + 0.1 * g[4]
    atm[7] = rng.uniform(0.0, np.pi) + 0.1 * g[4]
    atm[2] = base70[2] * (T0 / atm[1])
    atm[8] = base70[8] * (T0 / atm[1])
    return atm


def perturbed_batch(base: np.ndarray, ncol: int, ndep: int = 70, seed0: int = SEED0,
                    first: int = 0) -> np.ndarray:
    """-> [ncol, 9, ndep] pyrh-unit atmospheres."""
    b = resample_falc(base, ndep)
    return np.stack([perturbed_column(b, first + c, seed0) for c in range(ncol)])
