"""cuobjdump -sass of the hot kernels of librhb200.so with static opcode histograms -> profiles/<name>."""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
WANT = ("opacity_fused_kernelILi10ENS_11ZeemanParamELb0ELb0", "delo_raypts_kernelILi4", "continuum_tile_kernelILi8",
        "chemeq_coop_kernelILi16")


WANT_R2 = ("nlte_gamma_kernelILb1ELb0", "nlte_gamma_atom_kernel", "nlte_ray_kernelILi2ELi8", "nlte_opacity_kernel", "nlte_prd_scatter_kernel",
           "nlte_ray_stokes_kernel", "continuum_tile_kernelILi8", "delo_vcols_kernel")


def main():
    global WANT
    out_name = sys.argv[1] if len(sys.argv) > 1 else "r1_sass_hot_kernels.txt"
    if len(sys.argv) > 2 and sys.argv[2] == "r2":
        WANT = WANT_R2
    txt = subprocess.run(["cuobjdump", "-sass", str(ROOT / "pyrh_b200" / "csrc" / "librhb200.so")],
                         capture_output=True, text=True).stdout
    blocks = re.split(r"\n\s*Function : ", txt)
    out = []
    for b in blocks[1:]:
        name = b.split("\n", 1)[0].strip()
        if not any(w in name for w in WANT):
            continue
        ops = collections.Counter()
        body = []
        for ln in b.splitlines()[1:]:
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                ins = m.group(2).strip()
                body.append(f"        /*{m.group(1)}*/                   {ins} ;")
                op = re.sub(r"^@!?U?P\d+\s+", "", ins).split()[0].split(".")[0]
                ops[op] += 1
        hist = ", ".join(f"{k} {v}" for k, v in ops.most_common(24))
        tensor = [k for k in ops if k.startswith(("HMMA", "UTC", "TCGEN", "WGMMA"))]
        out.append(f"==== {name}\n static opcode histogram ({sum(ops.values())} instructions): {hist}\n"
                   f" FP64 arithmetic: DMUL {ops['DMUL']}, DADD {ops['DADD']}, DFMA {ops['DFMA']} (DFMA only inside the "
                   f"re-stated libm and the IEEE division sequences; -fmad=false); tensor-core opcodes: {tensor or 'none (by design)'}\n\n"
                   + "\n".join(body) + "\n")
    (ROOT / "profiles" / out_name).write_text("\n".join(out))
    print(f"{len(out)} kernels -> profiles/{out_name}")


if __name__ == "__main__":
    main()
