#!/usr/bin/env python
"""Minimal driver for ncu: a few launches of the fused reduction kernel at the bench shape."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from semanticlidarunc_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
T, B, C, H, W = 20, int(os.environ.get("SLU_B", "16")), 20, 64, 2048
g = torch.Generator(device=dev).manual_seed(1)
logits = torch.randn((T, B, C, H, W), generator=g, device=dev) * 3.0
labels = torch.randint(0, C, (B, H, W), generator=g, device=dev)
confmat, bins = ops.new_confmat(C, dev), ops.new_ece_bins(15, dev)
direct = os.environ.get("SLU_DIRECT", "0") == "1"
for _ in range(int(os.environ.get("SLU_N", "5"))):
    ops.reduce_metrics(logits, labels, kind="logits", conf_mode=ops.CONF_RENORM, ignore_index=0, confmat=confmat,
                       ece_bins=bins, direct=direct)
torch.cuda.synchronize()
print("ok", int(confmat.sum()))
