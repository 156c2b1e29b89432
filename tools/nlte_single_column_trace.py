"""Where the time of ONE NLTE atmosphere goes (BASELINE configs[3]: FAL-C, H + Ca II): host-side phase trace
(RHB200_NLTE_TRACE) + per-kernel-family CUDA-event times."""
import os, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ["RHB200_NLTE_TRACE"] = "1"
import bench
from pyrh_b200 import nlte_host
bench._pyrh_data_path()
case = sys.argv[1] if len(sys.argv) > 1 else "config4"
c = bench.NLTE_CASES[case]
s = nlte_host.NlteSession(bench._nlte_workdir(case), np.linspace(*c["wave"]))
base = np.load(ROOT / "tests" / "golden" / "falc_base.npy")
s.compute(base)
print("---- second call", file=sys.stderr)
t0 = time.perf_counter(); r = s.compute(base); dt = time.perf_counter() - t0
print(f"wall {1e3*dt:.1f} ms, iterations {int(r['niter'])}", file=sys.stderr)
os.environ.pop("RHB200_NLTE_TRACE")
s.ctx.timing(True); s.compute(base)
print({n: (round(ms, 2), cnt) for n, (ms, cnt) in s.ctx.timing_get().items() if cnt}, file=sys.stderr)
