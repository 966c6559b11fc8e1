#!/usr/bin/env python
"""Time the fused reduction kernel (CUDA events, median of 30) for whatever libslu SLU_LIB_PATH names."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from semanticlidarunc_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda", 0)
T, B, C, H, W = 20, 16, 20, 64, 2048
g = torch.Generator(device=dev).manual_seed(1)
logits = torch.randn((T, B, C, H, W), generator=g, device=dev) * 3.0
labels = torch.randint(0, C, (B, H, W), generator=g, device=dev)
confmat, bins = ops.new_confmat(C, dev), ops.new_ece_bins(15, dev)
bytes_algo = (4 * T * C + 8 + 24) * B * H * W
res = {"lib": os.path.basename(_lib.LIB_PATH)}


def timeit(fn, n=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(min(ts))


for direct in ([False, True] if os.environ.get("SLU_BOTH") else [False]):
    med, mn = timeit(lambda: ops.reduce_metrics(logits, labels, kind="logits", conf_mode=ops.CONF_RENORM, ignore_index=0,
                                                confmat=confmat, ece_bins=bins, direct=direct))
    res["direct" if direct else "staged"] = {"ms": round(med, 4), "ms_min": round(mn, 4), "GBps": round(bytes_algo / med / 1e6, 1)}
if hasattr(_lib.lib(), "slu_diag_read_stream") and os.environ.get("SLU_YARD"):
    out = torch.zeros(1, device=dev)
    flat = logits.view(-1)
    med, mn = timeit(lambda: _lib.lib().slu_diag_read_stream(_lib.ptr(flat), flat.numel(), _lib.ptr(out), _lib.stream_ptr()))
    res["read_stream_yardstick"] = {"ms": round(med, 4), "GBps": round(flat.numel() * 4 / med / 1e6, 1)}
    dst = torch.empty_like(flat[: flat.numel() // 2])
    med, mn = timeit(lambda: dst.copy_(flat[: flat.numel() // 2]))
    res["torch_copy_yardstick"] = {"ms": round(med, 4), "GBps_read_plus_write": round(dst.numel() * 8 / med / 1e6, 1)}
print(json.dumps(res))
