#!/usr/bin/env python
"""Minimal driver for ncu: a few batched projections + back-projections (HDL-64, B from SLU_B)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from semanticlidarunc_b200 import _lib, ops, synth  # noqa: E402
from semanticlidarunc_b200.dataset.definitions import build_id_lut  # noqa: E402

dev = torch.device("cuda", 0)
B = int(os.environ.get("SLU_B", "16"))
sensor = os.environ.get("SLU_SENSOR", "hdl64")
scans = [synth.synth_scan(i, sensor) for i in range(B)]
H, W = synth.SENSORS[sensor][4:6]
offs = np.concatenate([[0], np.cumsum([s[0].shape[0] for s in scans])])
xyzi = torch.from_numpy(np.concatenate([s[0] for s in scans])).to(dev)
raw = torch.from_numpy(np.concatenate([s[1] for s in scans]).view(np.int32)).to(dev)
lut = torch.from_numpy(build_id_lut()).to(dev)
if os.environ.get("SLU_EXACT") == "1":
    _lib.lib().slu_debug_project_exact(1)
ws = None
for _ in range(int(os.environ.get("SLU_N", "4"))):
    r = ops.project_batch(xyzi, raw, offs, H, W, lut=lut, workspace=ws)
    ws = r["workspace"]
    ops.backproject(r["label"], r["pix"], offs)
torch.cuda.synchronize()
print("ok", int((r["winner"] >= 0).sum()))
