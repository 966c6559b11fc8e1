#!/usr/bin/env python
"""GPU-side report for the fused reduction kernel: error statistics against the oracle and against an
fp64 evaluation of the same formulas, and CUDA-event timings of the staged (TMA) and direct variants.
Development tool (uses the oracle as checker); not part of the product or of bench.py."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import uncertainty as ou  # noqa: E402
from semanticlidarunc_b200 import ops, synth  # noqa: E402

dev = torch.device("cuda", 0)


def errs(name, got, ref, truth):
    got, ref, truth = (np.asarray(v, dtype=np.float64) for v in (got, ref, truth))
    out = {}
    for tag, r in (("vs_oracle_fp32", ref), ("vs_fp64", truth)):
        d = np.abs(got - r)
        out[tag] = {"max_abs": float(d.max()), "max_rel": float((d / np.maximum(np.abs(r), 1e-30)).max()),
                    "max_viol_1e-5rel_1e-6abs": float((d - (1e-6 + 1e-5 * np.abs(r))).max())}
    d = np.abs(ref - truth)
    out["oracle_fp32_vs_fp64"] = {"max_abs": float(d.max()), "max_rel": float((d / np.maximum(np.abs(truth), 1e-30)).max())}
    return {name: out}


report = {}
for scale in (3.0, 1.0, 10.0):
    x, lab = synth.synth_mc_logits(42, 20, 1, 20, 16, 2048, scale=scale)
    out = ops.reduce_metrics(x.to(dev), None, kind="logits", conf_mode=ops.CONF_RENORM, want=("H_norm", "MI_norm", "conf", "p_bar"))
    ref = ou.mc_reduce(x)
    tru = ou.mc_reduce(x.double())
    r = {}
    for k in ("H_norm", "MI_norm", "p_bar"):
        r.update(errs(k, out[k].cpu().numpy(), ref[k].numpy(), tru[k].numpy()))
    report[f"scale_{scale}"] = r

# timings
T, B, C, H, W = 20, 16, 20, 64, 2048
g = torch.Generator(device=dev).manual_seed(1)
logits = torch.randn((T, B, C, H, W), generator=g, device=dev) * 3.0
labels = torch.randint(0, C, (B, H, W), generator=g, device=dev)
confmat, bins = ops.new_confmat(C, dev), ops.new_ece_bins(15, dev)
bytes_algo = (4 * T * C + 8 + 24) * B * H * W
for direct in (False, True):
    for _ in range(3):
        ops.reduce_metrics(logits, labels, kind="logits", conf_mode=ops.CONF_RENORM, ignore_index=0, confmat=confmat, ece_bins=bins, direct=direct)
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.reduce_metrics(logits, labels, kind="logits", conf_mode=ops.CONF_RENORM, ignore_index=0, confmat=confmat, ece_bins=bins, direct=direct)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    report["direct" if direct else "staged"] = {"ms_median": ms, "ms_min": float(min(ts)), "GBps": bytes_algo / ms / 1e6}
print(json.dumps(report, indent=1))
