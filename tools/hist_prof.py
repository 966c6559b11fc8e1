#!/usr/bin/env python
"""Smallest command that launches the streaming confusion+reliability histogram on a 64-scan chunk (for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semanticlidarunc_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
C, n = 20, 64 * 64 * 2048
mode = sys.argv[1] if len(sys.argv) > 1 else "random"
if mode == "random":
    pred, lab, conf = (torch.randint(0, C, (n,), generator=g, device=dev), torch.randint(0, C, (n,), generator=g, device=dev),
                       torch.rand((n,), generator=g, device=dev))
else:
    lab = torch.randint(0, C, (n // 64,), generator=g, device=dev).repeat_interleave(64)
    pred = lab.clone(); pred[::7] = (pred[::7] + 1) % C
    conf = 0.9 + 0.1 * torch.rand((n,), generator=g, device=dev)
cm, bins = ops.new_confmat(C, dev), ops.new_ece_bins(15, dev)
for _ in range(3):
    ops.confusion_ece(pred, lab, conf, num_classes=C, ignore_index=0, confmat=cm, ece_bins=bins)
torch.cuda.synchronize()
print("ok", int(cm.sum()), int(bins[0].sum()))
