"""ctypes binding of librhb200.so (include/rhb200.h).  Fails loudly: there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
LIBPATH = CSRC / "librhb200.so"

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)
vp = C.c_void_p

RL_NFIELD, RE_NFIELD, RE_MAXSTAGE, AT_NFIELD = 24, 16, 12, 11
BC_IRRADIATED, BC_ZERO, BC_THERMALIZED = 0, 1, 2
K_PREP, K_OPACITY, K_DELO, K_BEZIER, K_OTHER = range(5)

# every symbol include/rhb200.h declares: (name, restype, argtypes)
# int (*rhb200_allreduce_fn)(void *user, double *device_buf, size_t count, int op)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int)
REDUCE_SUM, REDUCE_MAX = 0, 1

SYMBOLS = [
    ("rhb200_version", C.c_int, []),
    ("rhb200_last_error", C.c_char_p, []),
    ("rhb200_device_count", C.c_int, []),
    ("rhb200_device_info", C.c_int, [C.c_int, C.c_char_p, C.c_int, ip, C.POINTER(C.c_size_t), ip, ip]),
    ("rhb200_open", vp, [C.c_int]),
    ("rhb200_close", None, [vp]),
    ("rhb200_set_lines", C.c_int, [vp, C.c_int, dp, C.c_int, ip, dp, dp, C.c_int, dp, C.c_int, C.c_int,
                                   dp, dp, C.c_double, C.c_int, C.c_int]),
    ("rhb200_set_wavelengths", C.c_int, [vp, C.c_int, dp]),
    ("rhb200_update_line_strengths", C.c_int, [vp, C.c_int, ip, dp, dp, dp]),
    ("rhb200_get_line_windows", C.c_int, [vp, ip, ip, ip, C.c_int, ip]),
    ("rhb200_get_wavelength_flags", C.c_int, [vp, ip]),
    ("rhb200_nlte_compute1d_batch", C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, vp, C.c_int,
                                              C.c_double, C.c_double, vp, vp, vp, vp, vp, vp]),
    ("rhb200_nlte_compute1d_stokes_batch", C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, vp, C.c_int,
                                                     C.c_double, C.c_double, vp, vp, vp, vp, vp, vp, vp]),
    ("rhb200_nlte_front_debug", C.c_int, [vp, C.c_int, dp, C.c_size_t]),
    ("rhb200_lte_stokes_batch", C.c_int, [vp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                                          vp, vp, vp, vp]),
    ("rhb200_lte_stokes_batch_dev", C.c_int, [vp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                                              vp, vp, vp, vp]),
    ("rhb200_ltepops_elem_batch", C.c_int, [vp, C.c_int, C.c_int, dp, dp]),
    ("rhb200_rlk_opacity_batch", C.c_int, [vp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, dp, dp, dp, ip]),
    ("rhb200_stokes_bezier3_batch", C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                              C.c_int, ip, dp, dp, dp, dp, dp, dp, dp, dp]),
    ("rhb200_bezier3_batch", C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                                       ip, dp, dp, dp, dp, dp, dp, dp]),
    ("rhb200_rlk_determinate", C.c_int, [C.c_char_p, C.c_char_p, dp, ip, dp, ip]),
    ("rhb200_lande", C.c_double, [C.c_double, C.c_int, C.c_double]),
    ("rhb200_rlk_zeeman", C.c_int, [C.c_double, C.c_double, C.c_double, C.c_int, C.c_double, C.c_int, C.c_double,
                                    C.c_double, C.c_int, C.c_int, ip, dp, dp]),
    ("rhb200_determinate", C.c_int, [C.c_char_p, C.c_double, ip, dp, ip, dp]),
    ("rhb200_zeeman", C.c_int, [C.c_char_p, C.c_double, C.c_char_p, C.c_double, C.c_double, C.c_int, ip, dp, dp]),
    ("rhb200_nlte_set_exact_rates", C.c_int, [vp, C.c_int]),
    ("rhb200_nlte_set_shard", C.c_int, [vp, C.c_int, C.c_int, ALLREDUCE_FN, vp]),
    ("rhb200_nlte_set_shard_nccl", C.c_int, [vp, C.c_int, C.c_int, vp]),
    ("rhb200_nccl_unique_id", C.c_int, [C.c_char_p]),
    ("rhb200_nlte_set_shard_nccl_id", C.c_int, [vp, C.c_int, C.c_int, C.c_char_p]),
    ("rhb200_nlte_shard_range", C.c_int, [vp, C.c_int, C.c_int, ip, ip]),
    ("rhb200_molecular_opacity_batch", C.c_int, [vp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                                 dp, C.c_int, ip, dp, dp, C.c_double, C.c_int, dp, dp, dp, dp, dp, ip]),
    ("rhb200_passive_bb_batch", C.c_int, [vp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, dp, C.c_int,
                                          dp, dp, C.c_double, C.c_int, dp, dp, dp, dp, dp, ip]),
    ("rhb200_continuum_batch", C.c_int, [vp, vp, C.c_int, dp, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp]),
    ("rhb200_set_continuum", C.c_int, [vp, vp, dp]),
    ("rhb200_lte_stokes_batch_pops", C.c_int, [vp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, vp, vp, vp]),
    ("rhb200_set_chemistry", C.c_int, [vp, C.c_int, ip, C.c_int, dp]),
    ("rhb200_chemistry_batch", C.c_int, [vp, C.c_int, C.c_int, dp, dp, dp]),
    ("rhb200_lte_stokes_batch_atmos", C.c_int, [vp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, vp, vp]),
    ("rhb200_compute1d_batch", C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, vp, C.c_int, C.c_double,
                                         C.c_double, C.c_int, C.c_int, vp, vp]),
    ("rhb200_compute1d_batch_multi", C.c_int, [C.c_int, C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, vp,
                                               C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, vp, vp]),
    ("rhb200_shard_columns", C.c_int, [C.c_int, C.c_int, C.c_int, ip, ip]),
    ("rhb200_compute1d_rf_batch", C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, vp, C.c_int, C.c_double,
                                            C.c_double, C.c_int, C.c_int, vp, vp, vp]),
    ("rhb200_set_loggf_rf", C.c_int, [vp, C.c_int, ip]),
    ("rhb200_rf_fd_batch", C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, vp, C.c_int, C.c_double,
                                     C.c_double, C.c_int, C.c_int, C.c_int, ip, dp, vp]),
    ("rhb200_rf_fd_depths_batch", C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, vp, C.c_int, C.c_double,
                                            C.c_double, C.c_int, C.c_int, C.c_int, ip, dp, C.c_int, ip, vp]),
    ("rhb200_set_model_lines", C.c_int, [vp, C.c_int, dp]),
    ("rhb200_set_stokes_mode", C.c_int, [vp, C.c_int]),
    ("rhb200_set_scatter", C.c_int, [vp, C.c_int, C.c_double]),
    ("rhb200_set_molecular_lines", C.c_int, [vp, C.c_int, dp, C.c_int, dp]),
    ("rhb200_set_molecular_lines_zeeman", C.c_int, [vp, C.c_int, dp, C.c_int, dp, C.c_int, ip, dp, dp]),
    ("rhb200_set_passive_lines", C.c_int, [vp, C.c_int, dp, C.c_int, dp, dp]),
    ("rhb200_hse_batch", C.c_int, [vp, C.c_int, C.c_int, C.c_int, dp, dp, dp, C.c_double, C.c_double, C.c_double,
                                   dp, dp, dp, dp]),
    ("rhb200_set_elements", C.c_int, [vp, C.c_int, dp, C.c_int, C.c_int, dp, dp]),
    ("rhb200_solve_ne_batch", C.c_int, [vp, C.c_size_t, dp, dp, dp, C.c_int]),
    ("rhb200_set_gravity", C.c_int, [vp, C.c_double, C.c_double]),
    ("rhb200_get_scales_batch", C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_double, C.c_double,
                                          C.c_double, C.c_double, vp]),
    ("rhb200_set_solvers", C.c_int, [vp, C.c_int, C.c_int]),
    ("rhb200_scalar_ray_batch", C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                          C.c_int, ip, dp, dp, dp, dp, dp, dp, dp]),
    ("rhb200_stokes_ray_batch", C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                          C.c_int, ip, dp, dp, dp, dp, dp, dp, dp, dp]),
    ("rhb200_bezier3_rf_batch", C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, ip, dp,
                                          dp, dp, dp, dp, dp, dp, C.c_int, dp, dp, dp, dp]),
    ("rhb200_feautrier_batch", C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, ip, dp,
                                         dp, dp, dp, dp, dp, dp, dp]),
    ("rhb200_voigt_humlicek", C.c_int, [vp, C.c_int, dp, dp, dp, dp, ip]),
    ("rhb200_voigt_armstrong", C.c_int, [vp, C.c_int, dp, dp, dp, ip]),
    ("rhb200_math_probe", C.c_int, [vp, C.c_int, C.c_int, dp, dp, dp]),
    ("rhb200_nlte_iterate", C.c_int, [vp, vp, C.c_int, vp, C.c_int, C.c_int, C.c_double, ip, dp, C.c_int, dp, dp, dp, dp]),
    ("rhb200_nlte_formal", C.c_int, [vp, vp, C.c_int, vp, C.c_int, C.c_int, C.c_double, dp, ip]),
    ("rhb200_solve_linear_eq_batch", C.c_int, [vp, C.c_int, C.c_int, dp, dp, C.c_int]),
    ("rhb200_dev_alloc", C.c_int, [vp, C.c_size_t, C.POINTER(vp)]),
    ("rhb200_dev_free", C.c_int, [vp, vp]),
    ("rhb200_host_alloc_pinned", C.c_int, [C.c_size_t, C.POINTER(vp)]),
    ("rhb200_host_free_pinned", C.c_int, [vp]),
    ("rhb200_memcpy_h2d", C.c_int, [vp, vp, vp, C.c_size_t]),
    ("rhb200_memcpy_d2h", C.c_int, [vp, vp, vp, C.c_size_t]),
    ("rhb200_synchronize", C.c_int, [vp]),
    ("rhb200_flush_l2", C.c_int, [vp]),
    ("rhb200_timing_enable", C.c_int, [vp, C.c_int]),
    ("rhb200_timing_reset", C.c_int, [vp]),
    ("rhb200_timing_get", C.c_int, [vp, C.c_int, dp, C.POINTER(C.c_long)]),
    ("rhb200_timer_begin", C.c_int, [vp]),
    ("rhb200_timer_end", C.c_int, [vp, dp]),
    ("rhb200_fp64_peak", C.c_int, [vp, dp, dp]),
]


class RHB200Error(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raise (never fall back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIBPATH.exists():
        raise RHB200Error(f"{LIBPATH} not built: run `python -m pyrh_b200.build` "
                          "(pyrh_b200 has no CPU fallback)")
    lib = C.CDLL(str(LIBPATH))
    for name, res, args in SYMBOLS:
        f = getattr(lib, name)          # AttributeError if the header and the .so disagree
        f.restype = res
        f.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise RHB200Error(f"librhb200 error {rc}: {load().rhb200_last_error().decode()}")
