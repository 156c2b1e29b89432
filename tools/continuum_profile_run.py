"""Fused LTE path with the continuum on the device on NCOL benchmark columns (for ncu captures)."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np
import bench
from pyrh_b200 import api, continuum
from pyrh_b200.linelist import LineTable
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
ctx = api.Context(0)
g0, lam, at, chi, eta = bench.synth_inputs(ncol, 0, pinned=False)
chem, ab = bench.synth_chem(ncol, 0, pinned=False)
ctx.set_lines(LineTable.from_npz(g0)); ctx.set_wavelengths(lam)
ctx.set_continuum(continuum.ContinuumModel(dict(np.load('/root/repo/tests/golden/falc_full.npz'))), ab)
for _ in range(2):
    st = ctx.lte_stokes_batch_pops(at, chem)
print(np.isfinite(st).all())
ctx.timing(True)
st = ctx.lte_stokes_batch_pops(at, chem)
print({k: round(v[0] / max(v[1], 1), 3) for k, v in ctx.timing_get().items() if v[1]})
