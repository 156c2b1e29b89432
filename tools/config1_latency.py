"""BASELINE config 1: one FAL-C column (57 depths, B = 1 kG), Hinode window, through host.compute1d (pyrh.compute1d's
call): first call (parses the working directory, uploads tables) and steady-state latency per call."""
import sys, os, time, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from pyrh_b200 import host
os.environ.setdefault("PYRH_PATH", str(ROOT / "oracle" / "_ref" / "pyrh_path"))
cwd = str(ROOT / "oracle" / "_ref" / "inputs" / "benchmark")
g = dict(np.load(ROOT / "tests" / "golden" / "falc_full.npz"))
t0 = time.perf_counter()
out = host.compute1d(cwd, 1.0, 0, g["atmosphere"], g["wave"])
first = time.perf_counter() - t0
n = 200
t0 = time.perf_counter()
for _ in range(n):
    out = host.compute1d(cwd, 1.0, 0, g["atmosphere"], g["wave"])
steady = (time.perf_counter() - t0) / n
print(json.dumps({"workload": "config 1: FAL-C 57 depths, B = 1 kG, 301 wavelengths, one column per call (host.compute1d)",
                  "first_call_s": first, "steady_ms_per_call": steady * 1e3, "spectra_per_s": 1 / steady,
                  "bitwise_equal_to_reference": bool(np.array_equal(np.array(out[:4]), g["stokes"]))}))
