"""cProfile of one NLTE drop-in call (host side): python tools/nlte_host_profile.py [config4|config5_sample] [ncol]"""
import cProfile, pstats, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from pyrh_b200 import nlte_host, synthetic  # noqa: E402
case = sys.argv[1] if len(sys.argv) > 1 else "config4"
ncol = int(sys.argv[2]) if len(sys.argv) > 2 else 512
bench._pyrh_data_path()
c = bench.NLTE_CASES[case]
s = nlte_host.NlteSession(bench._nlte_workdir(case), np.linspace(*c["wave"]))
atm = synthetic.perturbed_batch(np.load(ROOT / "tests" / "golden" / "falc_base.npy"), ncol, ndep=bench.NDEP, first=10000)
s.compute(atm[:64]); s.ctx.synchronize()
for rep in range(2):
    t0 = time.perf_counter(); pr = cProfile.Profile(); pr.enable()
    res = s.compute(atm); s.ctx.synchronize()
    pr.disable(); print("call", rep, "seconds", time.perf_counter() - t0)
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
s.ctx.timing(True); t0 = time.perf_counter(); s.compute(atm); s.ctx.synchronize()
print("with per-launch events:", time.perf_counter() - t0, "kernel ms total", sum(v[0] for v in s.ctx.timing_get().values()))
s.close()
