"""rhb200_compute1d_batch on NCOL synthetic benchmark columns (for ncu launch lists / captures)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from pyrh_b200 import api, continuum, synthetic
from pyrh_b200.linelist import LineTable
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
nrep = int(sys.argv[2]) if len(sys.argv) > 2 else 1
G = ROOT / "tests" / "golden"
g0 = dict(np.load(G / "synth70_c0.npz"))
full = dict(np.load(G / "falc_full.npz"))
ctx = api.Context(0)
ctx.set_lines(LineTable.from_npz(g0)); ctx.set_wavelengths(g0["lam_spect"])
ctx.set_continuum(continuum.ContinuumModel(full), np.load(G / "synth70_chem.npz")["abundance"])
ctx.set_chemistry(full["ce_nuclei"][:, 1].astype(np.int32), full["ce_mol"])
w = float(np.load(G / "pyrh_scales.npz")["tau_abund_sums"][0])
atm = api.pinned_empty((ncol, 9, 70))
atm[:] = synthetic.perturbed_batch(np.load(G / "falc_base.npy"), ncol)
out = api.pinned_empty((ncol, 4, 302))
for _ in range(2):
    ctx.compute1d_batch(atm, wght_per_H=w, out=out, keep_lambda_ref=True)
import time
ctx.synchronize(); t0 = time.perf_counter()
for _ in range(nrep):
    ctx.compute1d_batch(atm, wght_per_H=w, out=out, keep_lambda_ref=True)
ctx.synchronize(); dt = (time.perf_counter() - t0) / nrep
print(f"ncol={ncol} {dt*1e3:.2f} ms/call {ncol/dt:.0f} spectra/s finite={np.isfinite(out).all()}")
