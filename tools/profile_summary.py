"""Summarise an .ncu-rep (read with `ncu -i`, no GPU needed) into profiles/<name>.md + .csv:
per kernel duration, DRAM bytes, FP64 pipe, occupancy, divergence, top stall reasons, and the
dynamic opcode mix from the source page."""
import collections
import csv
import io
import subprocess
import sys
from pathlib import Path

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe % of peak (active)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), blocks/SM"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle / SMSP"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__cycles_elapsed.max", "SM cycles"),
]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", str(rep), "--page", page, "--csv", *extra],
                         capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, name = Path(sys.argv[1]), sys.argv[2]
    out_dir = Path(__file__).resolve().parent.parent / "profiles"
    out_dir.mkdir(exist_ok=True)
    rows = ncu_csv(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    md = [f"# ncu summary: {name}", "", f"source: `{rep.name}` (`ncu --set full --clock-control none --import-source on`)", ""]
    with open(out_dir / f"{name}.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [k for k, _ in KEYS])
        for r in data:
            w.writerow([r[idx["Kernel Name"]].split("(")[0]] + [r[idx[k]] if k in idx else "" for k, _ in KEYS])
    for r in data:
        kname = r[idx["Kernel Name"]].split("(")[0].replace("<unnamed>::", "")
        md += [f"## {kname}", "", "| metric | value |", "|---|---|"]
        for k, label in KEYS:
            if k in idx:
                md.append(f"| {label} (`{k}`) | {r[idx[k]]} {units[idx[k]]} |")
        stalls = sorted(((float(r[i] or 0), h.split("issue_stalled_")[1].split("_per_")[0]) for h, i in idx.items()
                         if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")), reverse=True)
        if stalls:
            md += ["", "top stall reasons (warps stalled per issue): " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:6])]
        # dynamic opcode mix
        src = ncu_csv(rep, "source", ["--kernel-name", "regex:" + kname.split("<")[0]])
        if len(src) > 3:
            h2 = src[1]
            i2 = {h: i for i, h in enumerate(h2)}
            agg, thr = collections.Counter(), collections.Counter()
            for s in src[2:]:
                try:
                    txt = s[i2["Source"]].strip()
                    op = (txt.split()[1] if txt.startswith("@") else txt.split()[0]).split(".")[0]
                    agg[op] += int(s[i2["Instructions Executed"]])
                    thr[op] += int(s[i2["Thread Instructions Executed"]])
                except (IndexError, ValueError, KeyError):
                    continue
            tot = sum(agg.values()) or 1
            md += ["", "dynamic opcode mix (share of executed warp instructions, avg active threads):", "",
                   "| op | share | threads |", "|---|---|---|"]
            for op, c in agg.most_common(14):
                md.append(f"| {op} | {100*c/tot:.1f} % | {thr[op]/max(c,1):.1f} |")
        md.append("")
    (out_dir / f"{name}.md").write_text("\n".join(md))
    print("\n".join(md[:60]))


if __name__ == "__main__":
    main()
