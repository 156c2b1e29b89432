"""Debug: non-finite columns of the config-5 NLTE sample."""
import os, sys, subprocess
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import bench
from pyrh_b200 import nlte_host, synthetic
bench._pyrh_data_path()
case = sys.argv[1] if len(sys.argv) > 1 else "config5_sample"
ncol = int(sys.argv[2]) if len(sys.argv) > 2 else 512
c = bench.NLTE_CASES[case]
base = np.load(ROOT / "tests/golden/falc_base.npy")
s = nlte_host.NlteSession(bench._nlte_workdir(case), np.linspace(*c["wave"]), 0)
atm = synthetic.perturbed_batch(base, ncol, ndep=bench.NDEP, first=10000)
res = s.compute(atm)
I, n, ns, it = res["I"], res["n"], res["nstar"], res["niter"]
fI = np.isfinite(I).reshape(ncol, -1).all(1); fn = np.isfinite(n).reshape(ncol, -1).all(1); fs = np.isfinite(ns).reshape(ncol, -1).all(1)
print("ncol", ncol, "nonfinite I", (~fI).sum(), "n", (~fn).sum(), "nstar", (~fs).sum())
bad = np.nonzero(~(fI & fn))[0]
print("bad columns", bad[:30], "niter", it[bad[:30]])
print("niter histogram of bad:", np.bincount(it[bad])[:8], " of good:", np.percentile(it[fI & fn], [0, 50, 100]))
if len(bad):
    b = bad[0]
    print("column", b, "nonfinite I count", (~np.isfinite(I[b])).sum(), "of", I[b].size, "nonfinite n levels", (~np.isfinite(n[b])).any(axis=-1) if n[b].ndim > 1 else None)
    print("atm T min/max", atm[b, 1].min(), atm[b, 1].max(), "ne min", atm[b, 2].min(), "nH min", atm[b, 8].min())
    alone = s.compute(atm[b:b + 1])
    print("alone finite:", np.isfinite(alone["I"]).all(), np.isfinite(alone["n"]).all(), "niter", alone["niter"])
s.close()
# what the unmodified reference does on the first few bad columns
for b in bad[:4]:
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--nlte-ref-worker", case, str(10000 + int(b))], capture_output=True, text=True, timeout=300)
    print("reference on column", b, "rc", p.returncode, p.stdout.strip()[-80:], p.stderr.strip()[-200:])
