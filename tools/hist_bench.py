#!/usr/bin/env python
"""A/B timing of the standalone confusion + reliability histogram (config 4's inner kernel): generic warp-aggregated
kernel vs the streaming one, on random and on spatially coherent maps.  CUDA events, median of 20; the 671 MB input
exceeds L2.  Algorithmic bytes: 20 B/px (SURVEY.md 8d "M standalone")."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from semanticlidarunc_b200 import _lib, ops, synth  # noqa: E402

dev = torch.device("cuda", 0)
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
C, H, W, Bc = 20, 64, 2048, 256
g = torch.Generator(device=dev).manual_seed(0)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


lab = synth.synth_coherent_labels(3, Bc, C, H, W, device="cpu").to(dev)
cases = {
    "random pred, coherent labels, uniform conf": (torch.randint(0, C, (Bc, H, W), generator=g, device=dev), lab, torch.rand((Bc, H, W), generator=g, device=dev)),
    "random pred+labels, uniform conf": (torch.randint(0, C, (Bc, H, W), generator=g, device=dev), torch.randint(0, C, (Bc, H, W), generator=g, device=dev), torch.rand((Bc, H, W), generator=g, device=dev)),
}
pred2 = lab.clone()
pred2[:, ::7] = (pred2[:, ::7] + 1) % C
cases["coherent pred+labels, conf in top bin"] = (pred2, lab, 0.9 + 0.1 * torch.rand((Bc, H, W), generator=g, device=dev))
rows = []
for name, (p, l, c) in cases.items():
    for what, kw in (("confusion+bins", dict(conf=c, bins=True, cm=True)), ("confusion only", dict(conf=None, bins=False, cm=True)),
                     ("bins only", dict(conf=c, bins=True, cm=False))):
        res = {}
        for generic in (1, 0):
            _lib.lib().slu_debug_hist_generic(generic)
            cm, bins = ops.new_confmat(C, dev), ops.new_ece_bins(15, dev)
            fn = lambda: ops.confusion_ece(p, l, kw["conf"], num_classes=C, ignore_index=0, confmat=cm if kw["cm"] else None,
                                           ece_bins=bins if kw["bins"] else None)
            ms = timeit(fn)
            nbytes = (16 + (4 if kw["conf"] is not None else 0)) * p.numel()
            res["generic" if generic else "streaming"] = {"ms": round(ms, 4), "GBps": round(nbytes / ms / 1e6, 1), "frac_of_measured_peak": round(nbytes / ms / 1e6 / PEAK, 3),
                                                          "scans_per_s": round(Bc / ms * 1e3)}
        _lib.lib().slu_debug_hist_generic(0)
        rows.append({"inputs": name, "what": what, **res})
print(json.dumps({"peak_GBps": PEAK, "pixels": Bc * H * W, "rows": rows}, indent=1))
