#!/usr/bin/env python
"""Per-block phase timeline of the cell pipeline (SLU_P3_TIMES=1): globaltimer stamps written by thread 0 of every block,
read back from the head of the projection workspace.  Prints, per kernel and phase, the earliest / median / latest stamp
relative to the first stamp of the replay."""
import os, sys
os.environ["SLU_P3_TIMES"] = "1"
os.environ.setdefault("SLU_PROJECT_EXACT", "2")          # the stamps exist in the cell pipeline only
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semanticlidarunc_b200 import ops, synth
from semanticlidarunc_b200.dataset.definitions import build_id_lut
dev = torch.device("cuda", 0)
B = int(os.environ.get("SLU_B", "16")); sensor = os.environ.get("SLU_SENSOR", "hdl64")
scans = [synth.synth_scan(i, sensor) for i in range(B)]
H, W = synth.SENSORS[sensor][4:6]
offs = np.concatenate([[0], np.cumsum([s[0].shape[0] for s in scans])])
xyzi = torch.from_numpy(np.concatenate([s[0] for s in scans])).to(dev)
raw = torch.from_numpy(np.concatenate([s[1] for s in scans]).view(np.int32)).to(dev)
lut = torch.from_numpy(build_id_lut()).to(dev)
tr = (-np.pi / 8, np.pi / 8) if os.environ.get("SLU_FIXED") == "1" else None
r = ops.project_batch(xyzi, raw, offs, H, W, lut=lut, theta_range=tr)
ws = r["workspace"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    ops.project_batch(xyzi, raw, offs, H, W, lut=lut, workspace=ws, theta_range=tr)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    ops.project_batch(xyzi, raw, offs, H, W, lut=lut, workspace=ws, theta_range=tr)
for rep in range(2):
    flush.zero_(); ws[: 3 * 4096 * 64].zero_(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    t = ws[: 3 * 4096 * 64].view(torch.int64).reshape(3, 4096, 8).cpu().numpy()
    t0 = t[:, :, :7][t[:, :, :7] > 0].min()            # slot 7 holds the SM id, not a time
    print("replay", rep, "event ms", round(a.elapsed_time(b), 4))
    if os.environ.get("SLU_DUMP"):
        np.save(os.environ["SLU_DUMP"], t)
    for k, name in enumerate(("P1 extremes", "P2 points", "P3 resolve")):
        for s in range(7):
            v = t[k, :, s]; v = v[v > 0]
            if v.size:
                v = (v - t0) / 1000.0
                print(f"  {name} stamp {s}: blocks {v.size:5d}  first {v.min():7.2f}  median {np.median(v):7.2f}  last {v.max():7.2f} us")
