import sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np
from pyrh_b200 import nlte
from pyrh_b200.api import Context
g = dict(np.load('/root/repo/tests/golden/nlte_caii.npz'))
ctx = Context(0)
res = {}
for ncol in (1, 16, 256, 1024):
    prob = nlte.NlteProblem.from_golden(g, ncol=ncol)
    nlte.iterate(ctx, prob, nmax=2, limit=0.0)
    ctx.timing(True)
    t = time.perf_counter()
    out = nlte.iterate(ctx, prob)
    dt = time.perf_counter() - t
    kt = ctx.timing_get()
    ctx.timing(False)
    t = time.perf_counter()
    out = nlte.iterate(ctx, prob)
    dt2 = time.perf_counter() - t
    nit = int(out['niter'][0])
    hdr = prob.hdr
    nray_pts = 0
    res[ncol] = dict(wall_s_timed=dt, wall_s=dt2, niter=nit, kernels_ms={k: v[0] for k, v in kt.items()},
                     exact=bool(np.array_equal(out['n'][0], g['n_final'])))
    print(ncol, res[ncol])
json.dump(res, open('/root/repo/gpurun_out/nlte_timing.json', 'w'), indent=1)
