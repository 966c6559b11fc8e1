#!/usr/bin/env python
"""Minimal driver for ncu: fused evidential loss + evidential reduce kernels at B=16."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from semanticlidarunc_b200 import ops, synth  # noqa: E402

dev = torch.device("cuda", 0)
B, C, H, W = 16, 20, 64, 2048
g = torch.Generator(device=dev).manual_seed(0)
ev = torch.randn((B, C + 1, H, W), generator=g, device=dev) * 3.0
coh = synth.synth_coherent_labels(5, B, C, H, W).to(dev)
ev[:, :C].scatter_add_(1, coh.unsqueeze(1), torch.full((B, 1, H, W), 8.0, device=dev))
for _ in range(3):
    ops.evidential_loss_fused(ev, coh, ignore=(0,))
    ops.evidential_reduce(ev, coh, from_outputs=True, ignore_index=0)
    alpha = ops.evidential_reduce(ev, None, from_outputs=True, want=("alpha",))["alpha"]
    ops.dirichlet_loss(alpha, coh, ignore=(0,))
torch.cuda.synchronize()
print("ok")
