"""NLTE MALI solves (BASELINE configs 4 / 5): the CaII fixture (6 levels + 5 continua, 405 wavelengths, 3 rays) and the
H 6-level + CaII fixture of config 4 (895 wavelengths, 25 transitions), N identical FAL-C columns per call.
ray-points = columns x wavelengths x rays x 2 directions x depths x iterations (SURVEY 8(d) counting rule)."""
import sys, time, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from pyrh_b200 import nlte
from pyrh_b200.api import Context
ctx = Context(0)
res = {}
for fixture, cols in (("nlte_caii", (1, 256, 1024, 4096)), ("nlte_h_caii", (1, 256, 1024))):
    g = dict(np.load(ROOT / "tests" / "golden" / f"{fixture}.npz"))
    for ncol in cols:
        prob = nlte.NlteProblem.from_golden(g, ncol=ncol)
        nlte.iterate(ctx, prob, nmax=2, limit=0.0)
        ctx.timing(True)
        out = nlte.iterate(ctx, prob)
        kt = ctx.timing_get()
        ctx.timing(False)
        t = time.perf_counter()
        out = nlte.iterate(ctx, prob)
        dt = time.perf_counter() - t
        nit = int(out["niter"][0])
        h = prob.hdr
        pts = ncol * h["Nspect"] * h["Nrays"] * 2 * h["Ndep"] * nit
        res[f"{fixture}_{ncol}"] = dict(columns=ncol, wall_s=dt, niter=nit, atmospheres_per_s=ncol / dt,
                                        ray_points_per_s=pts / dt, nspect=h["Nspect"], nrays=h["Nrays"], ndep=h["Ndep"],
                                        kernels_ms={k: v[0] for k, v in kt.items() if v[1]},
                                        exact=bool(np.array_equal(out["n"][0], g["n_final"])))
        print(fixture, ncol, json.dumps(res[f"{fixture}_{ncol}"]))
(ROOT / "gpurun_out").mkdir(exist_ok=True)
json.dump(res, open(ROOT / "gpurun_out" / "nlte_timing.json", "w"), indent=1)
