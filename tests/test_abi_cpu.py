"""CPU tests of the boundary: the C-ABI library loads, exports every symbol that
include/rhb200.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from pyrh_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_symbols_all_exported(lib):
    hdr = (ROOT / "include" / "rhb200.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(rhb200_[a-z0-9_]+)\s*\(", hdr))
    from pyrh_b200 import _lib
    bound = {s[0] for s in _lib.SYMBOLS}
    assert declared == bound, (declared - bound, bound - declared)
    for name in declared:
        assert hasattr(lib, name)


def test_version_and_no_device_fails_loudly(lib):
    assert lib.rhb200_version() == 100
    if lib.rhb200_device_count() == 0:
        from pyrh_b200 import _lib
        from pyrh_b200.api import Context
        with pytest.raises(_lib.RHB200Error):
            Context()
        assert lib.rhb200_open(0) is None
        assert b"no CUDA device" in lib.rhb200_last_error()


def test_product_never_imports_oracle():
    for p in (ROOT / "pyrh_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h"):
            txt = p.read_text()
            assert "oracle" not in txt.replace("oracle/", "").lower() or "import" not in txt or \
                not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), p
            assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), p
            assert "rhport" not in txt, p


def test_unit_conversion_and_bproject_match_reference(golden):
    """pyrh-unit rows -> SI rows, bit-exact against what the reference held in Formal()."""
    name, g = golden
    from pyrh_b200 import api
    rows = api.atmos_rows_from_pyrh(g["atmosphere"], g["col_np"], g["col_height"])
    for f, i in api.AT.items():
        assert np.array_equal(rows[i], g["col_" + f]), f


def test_synthetic_generator_is_deterministic():
    from pyrh_b200 import synthetic
    base = np.load(ROOT / "tests" / "golden" / "falc_base.npy")
    a = synthetic.perturbed_batch(base, 3)
    b = synthetic.perturbed_batch(base, 2, first=1)
    assert a.shape == (3, 9, 70)
    assert np.array_equal(a[1:], b)
    g = np.load(ROOT / "tests" / "golden" / "synth70_c1.npz")
    assert np.array_equal(g["atmosphere"], a[1])


def test_bridged_reference_library_exports_the_pyrh_symbols():
    """oracle/_build/libpyrh_bridged.so (integration/build_bridged.sh): the reference's pyrh C library with the bridge
    compiled in must export exactly what rh.pxd:140-185 binds -- rhf1d, get_RLK_lines, hse, get_scales, get_ne_from_nH --
    plus the non-breaking additions rhf1d_batch / pyrh_b200_close, and be linked against librhb200."""
    import ctypes
    import subprocess
    from pathlib import Path
    import pytest
    so = Path(__file__).resolve().parent.parent / "oracle" / "_build" / "libpyrh_bridged.so"
    if not so.exists():
        pytest.skip("bridged library not built (needs /root/reference: integration/build_bridged.sh)")
    lib = ctypes.CDLL(str(so))
    for name in ("rhf1d", "get_RLK_lines", "hse", "get_scales", "get_ne_from_nH", "rhf1d_batch", "pyrh_b200_close",
                 "pyrh_b200_solve", "pyrh_b200_save_inputs"):
        assert hasattr(lib, name), name
    needed = subprocess.run(["readelf", "-d", str(so)], capture_output=True, text=True).stdout
    assert "librhb200.so" in needed


def test_committed_bench_lines_carry_every_contract_key():
    """The bench lines committed under profiles/ (GPU arm and reference arm of the same round) carry every key the
    measurement contract names -- a guard against silently dropping one when bench.py is edited."""
    import json
    root = Path(__file__).resolve().parent.parent
    line = json.loads((root / "profiles" / "r2_bench_n1.json").read_text().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in line, k
    assert "workload" in line["config"] and "model" not in line["config"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(line["clocks"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(line["roofline"])
    assert {"value", "unit", "cores", "kind", "sample"} <= set(line["cpu_baseline"])
    assert line["gpu_launches"] > 0 and line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["value"] != line["value"]
    assert abs(line["roofline"]["frac"] - line["roofline"]["achieved"] / line["roofline"]["peak"]) < 1e-12
    ref = json.loads((root / "profiles" / "r2_bench_reference_arm.json").read_text().strip().splitlines()[-1])
    assert ref["impl"] == "reference" and ref["metric"] == line["metric"] and ref["unit"] == line["unit"]
    assert ref["config"]["workload"] == line["config"]["workload"]
    assert ref["cpu_baseline"]["kind"] == "reference" and ref["e2e"]["h2d_bytes_per_step"] == 0
