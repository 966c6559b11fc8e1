#!/usr/bin/env python
"""Time slu_reduce_metrics of several libslu builds in ONE process (CUDA events, median of 30)."""
import ctypes as C
import glob
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from semanticlidarunc_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda", 0)
T, B, Cc, H, W = 20, 16, 20, 64, 2048
g = torch.Generator(device=dev).manual_seed(1)
logits = torch.randn((T, B, Cc, H, W), generator=g, device=dev) * 3.0
labels = torch.randint(0, Cc, (B, H, W), generator=g, device=dev)
confmat, bins = ops.new_confmat(Cc, dev), ops.new_ece_bins(15, dev)
pred = torch.empty((B, H, W), dtype=torch.int64, device=dev)
conf, hn, mi = (torch.empty((B, H, W), dtype=torch.float32, device=dev) for _ in range(3))
edges = _lib.edges_array(ops.uniform_edges(15))
bytes_algo = (4 * T * Cc + 8 + 24) * B * H * W
sig = _lib.SIGNATURES["slu_reduce_metrics"]
ref_out = None
for path in sys.argv[1:] or sorted(glob.glob(os.path.join(ROOT, "semanticlidarunc_b200", "libslu*.so"))):
    h = C.CDLL(path)
    fn = h.slu_reduce_metrics
    fn.restype, fn.argtypes = sig

    def call():
        rc = fn(_lib.ptr(logits), _lib.ptr(labels), T, B, Cc, H * W, 0, 1, 1e-12, 1, 1, 0, 15, edges,
                None, _lib.ptr(pred), _lib.ptr(conf), _lib.ptr(hn), _lib.ptr(mi), _lib.ptr(confmat), _lib.ptr(bins),
                _lib.stream_ptr())
        assert rc == 0, rc
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    ts = []
    for _ in range(30):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); call(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    chk = (float(hn.double().sum()), float(mi.double().sum()), int(pred.sum()))
    if ref_out is None:
        ref_out = chk
    same = all(abs(a - b) <= 1e-6 * abs(b) for a, b in zip(chk, ref_out))
    med = float(np.median(ts))
    print(json.dumps({"lib": os.path.basename(path), "ms": round(med, 4), "ms_min": round(min(ts), 4),
                      "GBps": round(bytes_algo / med / 1e6, 1), "same_result": same}), flush=True)
