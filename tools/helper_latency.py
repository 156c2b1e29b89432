"""Per-call latency of the pyrh helper drop-ins (hse, get_scales, get_ne_from_nH) with the resident sessions."""
import json, os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from pyrh_b200 import host
os.environ.setdefault("PYRH_PATH", str(ROOT / "oracle" / "_ref" / "pyrh_path"))
cwd = str(ROOT / "oracle" / "_ref" / "inputs" / "benchmark")
a = np.load(ROOT / "tests" / "golden" / "synth70_c0.npz")["atmosphere"]
out = {}
for name, fn in (("hse", lambda: host.hse(cwd, 0, a[0], a[1], 0.1)),
                 ("get_scales", lambda: host.get_scales(cwd, 0, a[0], a, 500.0)),
                 ("get_ne_from_nH", lambda: host.get_ne_from_nH(cwd, 0, a[0], a[1], a[8]))):
    t0 = time.perf_counter(); fn(); first = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(20):
        fn()
    out[name] = {"first_call_s": first, "steady_ms_per_call": 1e3 * (time.perf_counter() - t0) / 20}
host.close_sessions()
print(json.dumps(out))
