#!/usr/bin/env python
"""Hot spots of one kernel of an .ncu-rep (read on the CPU box): opcode table and the SASS instructions that collected
the most stall samples.  usage: ncu_hot.py report.ncu-rep [kernel_index=0] [top=30]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
kernels, h, cur = [], None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1] if len(r) > 1 else "?", "rows": []}
        kernels.append(cur)
        continue
    if r and r[0] == "Address":
        h = r
        continue
    if h and cur is not None and len(r) >= len(h) - 2:
        cur["rows"].append(r)
k = kernels[which]
data = k["rows"]
iS, iE, iN = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
tot = sum(int(r[iE]) for r in data)
tsamp = sum(int(r[iN]) for r in data)
print("kernel", which, k["name"][:80], "warp instructions", tot, "samples", tsamp)
by, sm = collections.Counter(), collections.Counter()
for r in data:
    toks = r[iS].strip().split()
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    by[op] += int(r[iE]); sm[op] += int(r[iN])
for op, c in by.most_common(16):
    print(f"  {op:10s} {c:12d} {100 * c / tot:5.1f}%  samples {100 * sm[op] / max(1, tsamp):5.1f}%")
print("hottest instructions (index, samples %, executed, SASS):")
order = sorted(range(len(data)), key=lambda i: -int(data[i][iN]))[:top]
for i in sorted(order):
    r = data[i]
    print(f"  {i:5d} {100 * int(r[iN]) / max(1, tsamp):5.1f}%  {int(r[iE]):9d}  {r[iS].strip()[:110]}")
