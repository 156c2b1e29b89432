"""get_atomic_rfs for a batch: rhb200_compute1d_rf_batch on NCOL perturbed benchmark columns in NO_STOKES mode, both Fe I
lines registered (BASELINE config 3's analytic counterpart).  Writes gpurun_out/loggf_rf_timing.json."""
import json
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from pyrh_b200 import api, continuum, synthetic
from pyrh_b200.linelist import LineTable
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
G = ROOT / "tests" / "golden"
g0 = dict(np.load(G / "synth70_c0.npz"))
full = dict(np.load(G / "falc_full.npz"))
ctx = api.Context(0)
ctx.set_lines(LineTable.from_npz(g0))
ctx.set_stokes_mode("NO_STOKES")
ctx.set_wavelengths(g0["lam_spect"])
ctx.set_continuum(continuum.ContinuumModel(full), np.load(G / "synth70_chem.npz")["abundance"])
ctx.set_chemistry(full["ce_nuclei"][:, 1].astype(np.int32), full["ce_mol"])
ctx.set_loggf_rf([0, 1])
w = float(np.load(G / "pyrh_scales.npz")["tau_abund_sums"][0])
atm = api.pinned_empty((ncol, 9, 70))
atm[:] = synthetic.perturbed_batch(np.load(G / "falc_base.npy"), ncol)
res = {}
for name, fn in (("compute1d_batch_no_stokes", lambda: ctx.compute1d_batch(atm, wght_per_H=w)),
                 ("compute1d_rf_batch_no_stokes", lambda: ctx.compute1d_rf_batch(atm, wght_per_H=w))):
    fn(); ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        out = fn()
    ctx.synchronize()
    dt = (time.perf_counter() - t0) / 3
    res[name] = {"ms_per_call": dt * 1e3, "columns_per_s": ncol / dt}
st, rf = out
for mode in ("NO_STOKES", "FULL_STOKES"):                       # where the time goes (kernel classes, ms per call)
    ctx.set_stokes_mode(mode); ctx.set_wavelengths(g0["lam_spect"])
    outp = api.pinned_empty((ncol, 4, 302))
    ctx.compute1d_batch(atm, wght_per_H=w, out=outp, keep_lambda_ref=True)
    ctx.timing(True)
    t0 = time.perf_counter()
    ctx.compute1d_batch(atm, wght_per_H=w, out=outp, keep_lambda_ref=True)
    ctx.synchronize()
    res["kernels_" + mode] = {"wall_ms": (time.perf_counter() - t0) * 1e3, **{k: v[0] for k, v in ctx.timing_get().items()}}
    ctx.timing(False)
res.update(ncol=ncol, nlambda=int(st.shape[2]), npar=int(rf.shape[2]), all_finite=bool(np.isfinite(rf).all()),
           nonzero_fraction=float((rf != 0).mean()))
print(json.dumps(res))
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "loggf_rf_timing.json").write_text(json.dumps(res, indent=1) + "\n")
