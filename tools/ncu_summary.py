#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): key raw metrics + executed instructions by opcode."""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
npix_warps = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg.per_second",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct"]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w} [{units[i]}]: {[r[i][:70] for r in rows[2:]]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h, data, k = None, [], 0
for r in rows:
    if r and r[0] == "Kernel Name":
        k += 1
        if k > 1:
            break
        continue
    if r and r[0] == "Address":
        h = r
        continue
    if h and len(r) >= len(h) - 2:
        data.append(r)
iS, iE, iN = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
tot = sum(int(r[iE]) for r in data)
tsamp = sum(int(r[iN]) for r in data)
by, sm = collections.Counter(), collections.Counter()
for r in data:
    toks = r[iS].strip().split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    by[op] += int(r[iE])
    sm[op] += int(r[iN])
print(f"total warp instructions {tot}" + (f"  per pixel-warp {tot / npix_warps:.0f}" if npix_warps else ""), " samples", tsamp)
for op, c in by.most_common(24):
    pp = f"{c / npix_warps:8.1f}" if npix_warps else ""
    print(f"  {op:10s} {c:12d} {100 * c / tot:5.1f}%  {pp}  samples {100 * sm[op] / max(1, tsamp):5.1f}%")
# stall columns
cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
tots = {c: sum(int(r[h.index(c)] or 0) for r in data) for c in cols}
ts = sum(tots.values())
print("stall samples:", ", ".join(f"{c[6:]} {100 * v / max(1, ts):.1f}%" for c, v in sorted(tots.items(), key=lambda kv: -kv[1])[:10]))
