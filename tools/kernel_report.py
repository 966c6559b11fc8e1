#!/usr/bin/env python
"""GPU-side time of each stage's kernels, free of Python / launch overhead: every stage is captured into a CUDA graph
once and the graph is replayed between CUDA events (median of 20, L2 flushed before each replay).  The wrapper-level
numbers of tools/stage_report.py include ~15-60 us of Python per call, which hides the kernels of the small configs.
Algorithmic bytes per SURVEY.md 8d.  JSON to stdout."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from semanticlidarunc_b200 import ops, synth  # noqa: E402
from semanticlidarunc_b200.dataset.definitions import build_id_lut  # noqa: E402

dev = torch.device("cuda", 0)
if os.environ.get("SLU_NO_PACKED", "0") == "1":          # A/B: the one-pixel-per-thread kernels instead of the packed f32x2 ones
    from semanticlidarunc_b200 import _lib
    _lib.lib().slu_debug_no_packed_loss(1)
    _lib.lib().slu_debug_no_packed_evidential(1)
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
flush_rd = torch.zeros(64 << 20, dtype=torch.float32, device=dev)       # 256 MB, only ever read
# Two flushes are timed.  "ms": 256 MB written before every replay (the prescribed flush).  That leaves the L2 full of DIRTY lines of
# the flush buffer, and the first ~126 MB a timed kernel allocates each push one of them out to HBM: write-back traffic of the flush
# itself, charged to the kernel (up to 19 us at the copy peak).  "ms_clean_flush": the same write followed by a 256 MB READ of
# another buffer, so the cache is just as cold but holds clean lines when the timed graph starts.
only = set(sys.argv[1:])


def graph_time(fn, n=20):
    """(median ms after the prescribed 256 MB WRITE flush, median ms after write + read flush)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    out = []
    for clean in (False, True):
        ts = []
        for _ in range(n):
            flush.zero_()
            if clean:
                flush_rd.sum()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.replay(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        out.append(float(np.median(ts)))
    return out[0], out[1]


rows = []


def add(name, fn, nbytes, units, note=""):
    if only and not any(o in name for o in only):
        return
    ms, ms_clean = graph_time(fn)
    gbs = nbytes / ms / 1e6
    rows.append({"stage": name, "ms": round(ms, 4), "algorithmic_MB": round(nbytes / 1e6, 2), "GBps": round(gbs, 1),
                 "frac_of_measured_peak": round(gbs / PEAK, 3), "scans_per_s": round(units / ms * 1e3, 1),
                 "ms_clean_flush": round(ms_clean, 4), "frac_of_measured_peak_clean_flush": round(nbytes / ms_clean / 1e6 / PEAK, 3), "note": note})
    print(rows[-1], file=sys.stderr)


lut = torch.from_numpy(build_id_lut()).to(dev)
for sensor, B in (("hdl64", 1), ("hdl64", 16), ("os1-128", 1), ("os1-128", 16)):
    scans = [synth.synth_scan(i, sensor) for i in range(B)]
    H, W = synth.SENSORS[sensor][4:6]
    offs = np.concatenate([[0], np.cumsum([s[0].shape[0] for s in scans])])
    xyzi = torch.from_numpy(np.concatenate([s[0] for s in scans])).to(dev)
    raw = torch.from_numpy(np.concatenate([s[1] for s in scans]).view(np.int32)).to(dev)
    res = ops.project_batch(xyzi, raw, offs, H, W, lut=lut)
    ws = res["workspace"]
    n = int(offs[-1])
    add(f"projection {sensor} B={B}", lambda: ops.project_batch(xyzi, raw, offs, H, W, lut=lut, workspace=ws), 20 * n + 24 * B * H * W, B, "4 launches")
    if sensor == "os1-128":      # config 3 as the reference exercises it: SemanticCUDAL, 128 rows, fixed +-pi/8 elevation range
        tr = (-np.pi / 8, np.pi / 8)
        add(f"projection {sensor} B={B}, fixed theta range (CUDAL)", lambda: ops.project_batch(xyzi, raw, offs, H, W, lut=lut, workspace=ws, theta_range=tr),
            20 * n + 24 * B * H * W, B, "init + fused point pass + ties + resolve")
    img = res["img"]
    add(f"loader normals {sensor} B={B}", lambda: ops.frame_normals(img), 24 * B * H * W, B, "Scharr gradients of x,y,z + cross product (build_normal_xyz)")
    lab_img, pix = res["label"], res["pix"]
    add(f"back-projection {sensor} B={B}", lambda: ops.backproject(lab_img, pix, offs), 8 * n + 8 * B * H * W, B)

T, C, H, W = 20, 20, 64, 2048
g = torch.Generator(device=dev).manual_seed(0)
for B in (1, 16):
    logits = torch.randn((T, B, C, H, W), generator=g, device=dev) * 3.0
    labels = torch.randint(0, C, (B, H, W), generator=g, device=dev)
    cm, bins = ops.new_confmat(C, dev), ops.new_ece_bins(15, dev)
    add(f"MC reduce+metrics T=20 B={B}", lambda: ops.reduce_metrics(logits, labels, kind="logits", conf_mode=ops.CONF_RENORM, ignore_index=0, confmat=cm, ece_bins=bins),
        (4 * T * C + 32) * B * H * W, B, "config 2")
    one = logits[0].contiguous()
    add(f"single-pass softmax entropy+ECE T=1 B={B}", lambda: ops.reduce_metrics(one, labels, kind="logits", ignore_index=0, confmat=cm, ece_bins=bins),
        (4 * C + 32) * B * H * W, B, "config 1")
    add(f"single-pass T=1 B={B}, maps only (no histograms)", lambda: ops.reduce_metrics(one, None, kind="logits"), (4 * C + 24) * B * H * W, B)
    ev = torch.randn((B, C + 1, H, W), generator=g, device=dev) * 3.0
    add(f"evidential reduce+metrics B={B}", lambda: ops.evidential_reduce(ev, labels, from_outputs=True, ignore_index=0, confmat=cm, ece_bins=bins),
        (4 * (C + 1) + 8 + 8 + 5 * 4) * B * H * W, B, "tester Dirichlet block")
    add(f"fused evidential loss fwd+bwd B={B}", lambda: ops.evidential_loss_fused(ev, labels, ignore=(0,)), 176 * B * H * W, B, "config 5: count kernel + fused kernel")
    alpha = torch.nn.functional.softplus(ev[:, :C]) + 1.0
    add(f"Dirichlet MSE+KL terms fwd+bwd B={B}", lambda: ops.dirichlet_loss(alpha, labels, ignore=(0,)), (80 + 8 + 160) * B * H * W, B, "two gradients written")
    del logits
print(json.dumps({"peak_GBps": PEAK, "rows": rows}, indent=1))
