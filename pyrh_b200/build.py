"""Build librhb200.so (sm_100a) in-tree with nvcc.

    python -m pyrh_b200.build [--force]

Flags: -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false
(-fmad=false: the reference x86-64 build has no FMA contraction).
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
LIB = CSRC / "librhb200.so"
SOURCES = ["rhb200_abi.cu", "rhb200_lines.cu", "rhb200_delo.cu", "rhb200_peak.cu", "rhb200_nlte.cu", "rhb200_zeeman.cu", "rhb200_continuum.cu", "rhb200_scales.cu", "rhb200_hse.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
         "-Xcompiler", "-fPIC,-O2,-ffp-contract=off", "-Xptxas", "-v"] + os.environ.get("RHB200_NVCC_EXTRA", "").split()


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [CSRC.parent.parent / "include" / "rhb200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    objs = []
    procs = []
    for s in SOURCES:
        o = CSRC / (Path(s).stem + ".o")
        cmd = [NVCC, *FLAGS, "-c", str(CSRC / s), "-o", str(o)]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(o))
    log = []
    for s, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {s}\n{out}")
        if p.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {s}")
    (CSRC / "ptxas.log").write_text("\n".join(log))
    subprocess.check_call([NVCC, "-shared", "-o", str(LIB), *objs, "-lcudart"])
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", LIB)
