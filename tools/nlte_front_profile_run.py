"""Profiling workload (run under ncu): NLTE through the drop-in call on a batch of perturbed columns.
    python tools/nlte_front_profile_run.py [config4|config5_sample] [ncol]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from pyrh_b200 import nlte_host, synthetic  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "config5_sample"
ncol = int(sys.argv[2]) if len(sys.argv) > 2 else 256
bench._pyrh_data_path()
c = bench.NLTE_CASES[case]
s = nlte_host.NlteSession(bench._nlte_workdir(case), np.linspace(*c["wave"]))
atm = synthetic.perturbed_batch(np.load(ROOT / "tests" / "golden" / "falc_base.npy"), ncol, ndep=bench.NDEP, first=10000)
res = s.compute(atm)
print(case, ncol, "columns:", "iterations", res["niter"].min(), res["niter"].max(), "ray-points", s.ray_points(res, bench.NDEP))
s.close()
