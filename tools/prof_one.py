#!/usr/bin/env python
"""Smallest command that launches one stage's kernels three times on a 16-scan batch (for ncu):
   python tools/prof_one.py evidential | loss | single | project"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semanticlidarunc_b200 import ops, synth  # noqa: E402
from semanticlidarunc_b200.dataset.definitions import build_id_lut  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
B, C, H, W = 16, 20, 64, 2048
what = sys.argv[1]
labels = torch.randint(0, C, (B, H, W), generator=g, device=dev)
cm, bins = ops.new_confmat(C, dev), ops.new_ece_bins(15, dev)
if what in ("evidential", "loss"):
    ev = torch.randn((B, C + 1, H, W), generator=g, device=dev) * 3.0
    fn = (lambda: ops.evidential_reduce(ev, labels, from_outputs=True, ignore_index=0, confmat=cm, ece_bins=bins)) if what == "evidential" \
        else (lambda: ops.evidential_loss_fused(ev, labels, ignore=(0,)))
elif what == "single":
    one = torch.randn((B, C, H, W), generator=g, device=dev) * 3.0
    fn = lambda: ops.reduce_metrics(one, labels, kind="logits", ignore_index=0, confmat=cm, ece_bins=bins)
else:
    scans = [synth.synth_scan(i, "hdl64") for i in range(B)]
    offs = np.concatenate([[0], np.cumsum([s[0].shape[0] for s in scans])])
    xyzi = torch.from_numpy(np.concatenate([s[0] for s in scans])).to(dev)
    raw = torch.from_numpy(np.concatenate([s[1] for s in scans]).view(np.int32)).to(dev)
    lut = torch.from_numpy(build_id_lut()).to(dev)
    ws = ops.project_batch(xyzi, raw, offs, H, W, lut=lut)["workspace"]
    fn = lambda: ops.project_batch(xyzi, raw, offs, H, W, lut=lut, workspace=ws)
for _ in range(3):
    fn()
torch.cuda.synchronize()
print("ok")
