"""CPU test: the device math header (exp/pow/sin/cos re-statements of glibc 2.39's FMA variants),
compiled for the host, must agree BIT FOR BIT with this machine's glibc on random arguments.
Skipped when the host libm is not glibc 2.39 or the CPU has no FMA (glibc then selects other
ifunc variants)."""
import platform
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _glibc_fma():
    try:
        flags = Path("/proc/cpuinfo").read_text()
    except OSError:
        return False
    return platform.libc_ver()[0] == "glibc" and platform.libc_ver()[1] == "2.39" and " fma " in flags \
        and " avx2 " in flags


@pytest.mark.skipif(not _glibc_fma(), reason="needs glibc 2.39 on an FMA+AVX2 x86-64 host")
def test_device_math_matches_glibc_bitwise(tmp_path):
    exe = tmp_path / "mathcheck"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-mfma", "-ffp-contract=off",
                           "-I", str(ROOT / "pyrh_b200" / "csrc"),
                           str(ROOT / "tests" / "helpers" / "mathcheck.cpp"), "-o", str(exe), "-lm"])
    out = subprocess.run([str(exe), "400000"], capture_output=True, text=True)
    rows = [ln.split() for ln in out.stdout.strip().splitlines()]
    assert len(rows) >= 20
    for r in rows:
        name, n, bad = " ".join(r[:-3]), int(r[-3]), int(r[-2])
        assert bad == 0, f"{name}: {bad}/{n} results differ from glibc\n{out.stderr}"
    assert out.returncode == 0


def test_math_tables_regenerate_identically(tmp_path):
    """tools/gen_math_tables.py (first-principles construction) reproduces the committed tables."""
    committed = (ROOT / "pyrh_b200" / "csrc" / "rhb200_math_tables.inc").read_text()
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen", ROOT / "tools" / "gen_math_tables.py")
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    e = gen.exp_table()
    assert gen.fmt_u64(e).replace("\n", " \\\n") in committed
    p = gen.powlog_table()
    assert gen.fmt_f64(p).replace("\n", " \\\n") in committed
