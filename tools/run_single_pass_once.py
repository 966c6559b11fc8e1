#!/usr/bin/env python
"""Minimal driver for ncu: the fused kernel on single-pass logits (T=1), config 1's shape batched."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from semanticlidarunc_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
B, C, H, W = int(os.environ.get("SLU_B", "16")), 20, 64, 2048
g = torch.Generator(device=dev).manual_seed(1)
logits = torch.randn((B, C, H, W), generator=g, device=dev) * 3.0
labels = torch.randint(0, C, (B, H, W), generator=g, device=dev)
cm, bins = ops.new_confmat(C, dev), ops.new_ece_bins(15, dev)
for _ in range(5):
    ops.reduce_metrics(logits, labels, kind="logits", ignore_index=0, confmat=cm, ece_bins=bins)
torch.cuda.synchronize()
print("ok")
