#!/usr/bin/env python
"""Per-stage timings on one B200 for every BASELINE.json config: CUDA events, median of 20 after 3
warm-ups, algorithmic bytes per SURVEY.md 8d.  Writes JSON to stdout (kept under profiles/)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from semanticlidarunc_b200 import ops, synth  # noqa: E402
from semanticlidarunc_b200.dataset.definitions import build_id_lut  # noqa: E402
from semanticlidarunc_b200.losses.dirichlet_losses import DirichletMSELoss  # noqa: E402
from semanticlidarunc_b200.losses.regularizers import KL_offClasses_to_uniform  # noqa: E402

dev = torch.device("cuda", 0)
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=20, flush_l2=False):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        if flush_l2:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def row(name, ms, nbytes, units, unit_name, note=""):
    gbs = nbytes / ms / 1e6
    return {"stage": name, "ms": round(ms, 4), "algorithmic_MB": round(nbytes / 1e6, 2), "GBps": round(gbs, 1),
            "frac_of_measured_peak": round(gbs / PEAK, 3), unit_name + "_per_s": round(units / ms * 1e3, 1), "note": note}


rows = []
lut = torch.from_numpy(build_id_lut()).to(dev)
for sensor, B in (("hdl64", 1), ("hdl64", 16), ("os1-128", 1), ("os1-128", 16)):
    scans = [synth.synth_scan(i, sensor) for i in range(B)]
    H, W = synth.SENSORS[sensor][4:6]
    offs = np.concatenate([[0], np.cumsum([s[0].shape[0] for s in scans])])
    xyzi = torch.from_numpy(np.concatenate([s[0] for s in scans])).to(dev)
    raw = torch.from_numpy(np.concatenate([s[1] for s in scans]).view(np.int32)).to(dev)
    ws = [None]
    def proj():
        r = ops.project_batch(xyzi, raw, offs, H, W, lut=lut, workspace=ws[0])
        ws[0] = r["workspace"]
        return r
    res = proj()
    n = int(offs[-1])
    rows.append(row(f"projection {sensor} B={B}", timeit(proj, flush_l2=True), 20 * n + 24 * B * H * W, B, "scans", "4 launches, L2 flushed between runs"))
    lab_img = res["label"]
    rows.append(row(f"back-projection {sensor} B={B}", timeit(lambda: ops.backproject(lab_img, res["pix"], offs), flush_l2=True),
                    8 * n + 8 * B * H * W, B, "scans", "pix int32 in, int64 out, int64 label image"))

T, C, H, W = 20, 20, 64, 2048
g = torch.Generator(device=dev).manual_seed(0)
for B in (1, 16):
    logits = torch.randn((T, B, C, H, W), generator=g, device=dev) * 3.0
    labels = torch.randint(0, C, (B, H, W), generator=g, device=dev)
    cm, bins = ops.new_confmat(C, dev), ops.new_ece_bins(15, dev)
    f = lambda: ops.reduce_metrics(logits, labels, kind="logits", conf_mode=ops.CONF_RENORM, ignore_index=0, confmat=cm, ece_bins=bins)
    rows.append(row(f"MC reduce+metrics T=20 B={B}", timeit(f, flush_l2=(B == 1)), (4 * T * C + 32) * B * H * W, B, "scans",
                    "config 2" if B == 16 else "single scan: 213 MB"))
    one = logits[0].contiguous()
    f1 = lambda: ops.reduce_metrics(one, labels, kind="logits", ignore_index=0, confmat=cm, ece_bins=bins)
    rows.append(row(f"single-pass softmax entropy+ECE T=1 B={B}", timeit(f1, flush_l2=True), (4 * C + 32) * B * H * W, B, "scans", "config 1 shape"))
    ev = torch.randn((B, C + 1, H, W), generator=g, device=dev) * 3.0
    fe = lambda: ops.evidential_reduce(ev, labels, from_outputs=True, ignore_index=0, confmat=cm, ece_bins=bins)
    rows.append(row(f"evidential reduce+metrics B={B}", timeit(fe, flush_l2=True), (4 * (C + 1) + 8 + 8 + 5 * 4) * B * H * W, B, "scans", "MUFU / issue-bound"))
    alpha = (torch.nn.functional.softplus(ev[:, :C]) + 1.0).requires_grad_(True)
    mse, kl = DirichletMSELoss(ignore_index=0), KL_offClasses_to_uniform(ignore_index=0)
    def loss_step():
        alpha.grad = None
        (mse(alpha, labels) + 0.05 * kl(alpha, labels)).backward()
    rows.append(row(f"Dirichlet MSE+KL fwd+bwd B={B}", timeit(loss_step, flush_l2=True), 176 * B * H * W, B, "scans",
                    "two term kernels + autograd scaling passes (torch)"))
    from semanticlidarunc_b200.losses.evidential import EvidentialLoss
    fused = EvidentialLoss(1.0, 0.05, ignore_index=0)
    # (a) i.i.d. logits per pixel: every warp sees every class as somebody's arg-max (worst case for the
    #     per-class fast path); (b) spatially coherent arg-max in 32x32 patches, as a trained model on real scans
    coh = synth.synth_coherent_labels(5, B, C, H, W, device="cpu").to(dev)
    ev_coh = ev.clone()
    ev_coh[:, :C].scatter_add_(1, coh.unsqueeze(1), torch.full((B, 1, H, W), 8.0, device=dev))
    for tag, e_in, lab_in in (("i.i.d. logits", ev, labels), ("coherent arg-max", ev_coh, coh)):
        evg = e_in.clone().requires_grad_(True)
        def fused_step():
            evg.grad = None
            fused(evg, lab_in)[0].backward()
        rows.append(row(f"fused evidential loss fwd+bwd from head outputs B={B}, {tag}", timeit(fused_step, flush_l2=True),
                        176 * B * H * W, B, "scans", "count kernel + fused kernel + one autograd scale pass"))
        fe2 = lambda: ops.evidential_reduce(e_in, lab_in, from_outputs=True, ignore_index=0, confmat=cm, ece_bins=bins)
        rows.append(row(f"evidential reduce+metrics B={B}, {tag}", timeit(fe2, flush_l2=True), (4 * (C + 1) + 8 + 8 + 5 * 4) * B * H * W, B, "scans", ""))
    del logits

# config 4: 4k-scan sweep of confusion + ECE from reduced maps (chunks of 256 scans)
Bc = 256
pred = torch.randint(0, C, (Bc, H, W), generator=g, device=dev)
lab = synth.synth_coherent_labels(3, Bc, C, H, W, device="cpu").to(dev)
conf = torch.rand((Bc, H, W), generator=g, device=dev)
cm, bins = ops.new_confmat(C, dev), ops.new_ece_bins(15, dev)
fh = lambda: ops.confusion_ece(pred, lab, conf, num_classes=C, ignore_index=0, confmat=cm, ece_bins=bins)
rows.append(row("confusion+ECE standalone, 256 scans, random pred / coherent labels", timeit(fh), 20 * Bc * H * W, Bc, "scans", "config 4 inner kernel"))
pred2 = lab.clone()
pred2[:, ::7] = (pred2[:, ::7] + 1) % C
conf2 = (0.9 + 0.1 * torch.rand((Bc, H, W), generator=g, device=dev))
fh2 = lambda: ops.confusion_ece(pred2, lab, conf2, num_classes=C, ignore_index=0, confmat=cm, ece_bins=bins)
rows.append(row("confusion+ECE standalone, 256 scans, coherent pred+labels, conf in top bin", timeit(fh2), 20 * Bc * H * W, Bc, "scans", "what a trained model produces"))
print(json.dumps({"peak_GBps": PEAK, "rows": rows}, indent=1))
