"""Per-kernel-family time of the NLTE drop-in call on a batch of perturbed columns (CUDA events around every launch).
    python tools/nlte_family_timing.py [config4|config5_sample] [ncol] [N_MAX_ITER]
prints {"family": [ms total, launches, ms per launch]}."""
import json
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
if len(sys.argv) > 3:                     # cap the MALI iterations: every launch then runs with all columns active
    c["kw"] = dict(c["kw"], N_MAX_ITER=int(sys.argv[3]))
s = nlte_host.NlteSession(bench._nlte_workdir(case), np.linspace(*c["wave"]))
atm = synthetic.perturbed_batch(np.load(ROOT / "tests" / "golden" / "falc_base.npy"), ncol, ndep=bench.NDEP, first=10000)
s.compute(atm[:32])
s.ctx.timing(True)
res = s.compute(atm)
out = {n: [ms, cnt, ms / cnt] for n, (ms, cnt) in s.ctx.timing_get().items() if cnt}
out["iterations_median"] = float(np.median(res["niter"]))
print(json.dumps(out))
s.close()
