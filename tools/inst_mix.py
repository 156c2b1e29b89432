"""Dynamic SASS instruction mix per kernel from `ncu -i X.ncu-rep --page source --csv --print-source sass`."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
kern, data, hdr = None, collections.OrderedDict(), None
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        kern = r[1][:70]
        data[kern] = []
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if kern and len(r) > 6 and r[0].startswith("0x"):
        data[kern].append(r)
iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
for k, rs in data.items():
    tot = sum(int(r[iI]) for r in rs)
    cnt, smp = collections.Counter(), collections.Counter()
    for r in rs:
        ins = re.sub(r"^@!?U?P\d+\s+", "", r[1].strip())
        op = ins.split()[0].split(".")[0]
        if ins.startswith("IMAD.MOV"):
            op = "IMAD.MOV"
        cnt[op] += int(r[iI])
        smp[op] += int(r[iS])
    print(k, "total warp inst", tot)
    for op, c in cnt.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 18):
        print(f"   {op:10s} {c/tot*100:5.1f}%  samples {smp[op]/max(1, sum(smp.values()))*100:5.1f}%")
