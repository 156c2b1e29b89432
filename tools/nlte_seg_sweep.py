"""Sweep of the rate-accumulation segment length (RHB200_NLTE_GAMMA_SEG): kernel-family times of one config-5 batch."""
import json, os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
code = '''
import sys, json, numpy as np
sys.path.insert(0, %r)
import bench
from pyrh_b200 import nlte_host, synthetic
bench._pyrh_data_path()
c = bench.NLTE_CASES["config5_sample"]
s = nlte_host.NlteSession(bench._nlte_workdir("config5_sample"), np.linspace(*c["wave"]))
atm = synthetic.perturbed_batch(np.load(%r), 512, ndep=70, first=10000)
s.compute(atm[:64])
s.ctx.timing(True)
s.compute(atm)
print(json.dumps({n: round(ms, 1) for n, (ms, cnt) in s.ctx.timing_get().items() if cnt}))
''' % (str(ROOT), str(ROOT / "tests" / "golden" / "falc_base.npy"))
out = {}
for seg in sys.argv[1:] or ["4", "8", "16", "32", "exact"]:
    env = dict(os.environ)
    if seg == "exact":
        env["RHB200_NLTE_EXACT"] = "1"
    else:
        env["RHB200_NLTE_GAMMA_SEG"] = seg
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    out[seg] = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else r.stderr[-300:]
    print(seg, out[seg], flush=True)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "r2_nlte_seg_sweep.json").write_text(json.dumps(out, indent=1))
