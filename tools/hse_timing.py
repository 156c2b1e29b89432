"""pyrh.hse for NCOL perturbed FAL-C temperature runs (70 depths, log tau500 grid) in one rhb200_hse_batch call."""
import sys, time, json, os
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from pyrh_b200 import host, synthetic
os.environ.setdefault("PYRH_PATH", str(ROOT / "oracle" / "_ref" / "pyrh_path"))
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
atm = synthetic.perturbed_batch(np.load(ROOT / "tests/golden/falc_base.npy"), ncol)
s = host.HseSession(str(ROOT / "oracle" / "_ref" / "inputs" / "benchmark"))
s.hse(0, atm[:64, 0], atm[:64, 1], 0.1)
t0 = time.perf_counter()
ne, nH, rho, pg = s.hse(0, atm[:, 0], atm[:, 1], 0.1)
dt = time.perf_counter() - t0
launches = sum(v[1] for v in s.ctx.timing_get().values())
s.ctx.timing(True)
t1 = time.perf_counter()
s.hse(0, atm[:, 0], atm[:, 1], 0.1)
dt_timed = time.perf_counter() - t1
print(json.dumps({"workload": f"pyrh.hse: {ncol} columns x 70 depths, log tau500 grid, pg_top = 0.1 Pa", "columns": ncol,
                  "seconds": dt, "columns_per_s": ncol / dt, "finite": bool(np.isfinite(pg).all()),
                  "pg_bottom_mean": float(pg[:, -1].mean()), "kernel_launch_groups": int(launches),
                  "second_call_with_per_launch_events_s": dt_timed,
                  "kernel_ms": {k: round(v[0], 2) for k, v in s.ctx.timing_get().items() if v[1]}}))
