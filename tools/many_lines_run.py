"""Robustness / throughput with a long Kurucz list: benchmark/lines_4016 replicated 20 times with 0.08 nm shifts (360
lines, ~5700 Zeeman components -> the global-memory Zeeman path), 641 wavelengths, NCOL benchmark columns."""
import sys, os, time, json, shutil, tempfile
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from pyrh_b200 import host, synthetic
os.environ.setdefault("PYRH_PATH", str(ROOT / "oracle" / "_ref" / "pyrh_path"))
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 256
src = ROOT / "oracle" / "_ref" / "inputs" / "benchmark"
tmp = Path(tempfile.mkdtemp())
for f in src.iterdir():
    if f.is_file() and f.suffix not in (".fits", ".spec", ".py"):
        shutil.copy(f, tmp / f.name)
base = [ln for ln in (src / "lines_4016").read_text().splitlines() if ln.strip()]
out = []
for rep in range(20):
    for ln in base:
        lam = float(ln[:10]) + 0.08 * rep
        out.append(f"{lam:10.4f}" + ln[10:])
(tmp / "many").write_text("\n".join(out) + "\n")
(tmp / "kurucz.input").write_text("many\n")
wave = np.linspace(401.4, 403.4, 641)
t0 = time.perf_counter()
s = host.Session(str(tmp), wave)
t_open = time.perf_counter() - t0
atm = synthetic.perturbed_batch(np.load(ROOT / "tests/golden/falc_base.npy"), ncol)
st = s.compute(atm[:8])
t0 = time.perf_counter()
st = s.compute(atm)
dt = time.perf_counter() - t0
first, count, idx = s.ctx.line_windows()
print(json.dumps({"lines": s.lt.nline, "zeeman_components": int(len(s.lt.zq)), "wavelengths": len(wave), "columns": ncol,
                  "lines_per_wavelength_mean": float(count.mean()), "lines_per_wavelength_max": int(count.max()),
                  "open_s": t_open, "seconds": dt, "spectra_per_s": ncol / dt,
                  "ray_points_per_s": ncol * len(wave) * 70 / dt, "finite": bool(np.isfinite(st).all()),
                  "line_depth": float(1 - st[0, 0].min() / st[0, 0].max())}))
