"""BASELINE config 3: finite-difference response functions (T, v_z, B, gamma, chi per depth) of NCOL atmospheres."""
import sys, time, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from pyrh_b200 import api, continuum, synthetic
from pyrh_b200.linelist import LineTable
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 64
G = ROOT / "tests" / "golden"
g0 = dict(np.load(G / "synth70_c0.npz")); full = dict(np.load(G / "falc_full.npz"))
ctx = api.Context(0)
ctx.set_lines(LineTable.from_npz(g0)); ctx.set_wavelengths(g0["lam_spect"])
ctx.set_continuum(continuum.ContinuumModel(full), np.load(G / "synth70_chem.npz")["abundance"])
ctx.set_chemistry(full["ce_nuclei"][:, 1].astype(np.int32), full["ce_mol"])
w = float(np.load(G / "pyrh_scales.npz")["tau_abund_sums"][0])
atm = synthetic.perturbed_batch(np.load(G / "falc_base.npy"), ncol)
rows = np.array([1, 3, 5, 6, 7], np.int32); delta = np.array([1.0, 0.01, 1.0, 1e-3, 1e-3])
out = api.pinned_empty((ncol, 5, 70, 4, 302))
import os
ctx.rf_fd_batch(atm[:4], rows, delta, wght_per_H=w, out=out[:4], keep_lambda_ref=True)
ctx.synchronize(); t0 = time.perf_counter()
ctx.rf_fd_batch(atm, rows, delta, wght_per_H=w, out=out, keep_lambda_ref=True)      # first full-size call: reserves the device workspace
ctx.synchronize(); dt_first = time.perf_counter() - t0
ctx.timing(True); t0 = time.perf_counter()
ctx.rf_fd_batch(atm, rows, delta, wght_per_H=w, out=out, keep_lambda_ref=True)
ctx.synchronize(); dt = time.perf_counter() - t0
kernels = ctx.timing_get()
ctx.timing(False)
nsyn = ncol * 5 * 70 * 2
rec = {"workload": f"config 3: centred FD response functions, {ncol} atmospheres x 5 parameters x 70 depths",
       "route": "single-depth (1 + 2 npar full columns per atmosphere, per-perturbation scale walk + formal solution)",
       "atmospheres": ncol, "syntheses": nsyn, "seconds": dt, "atmospheres_per_s": ncol / dt,
       "first_call_seconds": dt_first,      # includes the cudaMalloc of the workspace this batch size needs
       "syntheses_per_s": nsyn / dt, "ray_points_per_s": nsyn * 301 * 70 / dt,
       "h2d_bytes": int(atm.nbytes), "d2h_bytes": int(out.nbytes), "finite": bool(np.isfinite(out).all()),
       "kernel_ms": kernels}
# the brute-force expansion (every perturbed column synthesised in full) on a slice, for the ratio and a bitwise check
nb = min(ncol, 128)
ref = np.empty((nb, 5, 70, 4, 302))
os.environ["RHB200_RF_FD_BRUTE"] = "1"
ctx.rf_fd_batch(atm[:4], rows, delta, wght_per_H=w, out=ref[:4], keep_lambda_ref=True)
ctx.synchronize(); t0 = time.perf_counter()
ctx.rf_fd_batch(atm[:nb], rows, delta, wght_per_H=w, out=ref, keep_lambda_ref=True)
ctx.synchronize(); db = time.perf_counter() - t0
rec["brute_force"] = {"atmospheres": nb, "seconds": db, "atmospheres_per_s": nb / db,
                      "bitwise_equal_to_single_depth_route": bool(np.array_equal(ref, out[:nb]))}
print(json.dumps(rec))
