#!/usr/bin/env python
"""Scratch diagnostics for round 2 (GPU box): packed-vs-scalar evidential outputs, Dirichlet AUROC scores."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from semanticlidarunc_b200 import _lib, ops, synth
from oracle import metrics as om
dev = torch.device("cuda", 0)

# (a) packed vs scalar evidential
B, C, H, W = 3, 20, 8, 130
x, lab = synth.synth_evidential_logits(3, B, C, H, W)
x, lab = x.to(dev), lab.to(dev)
want = ("alpha", "pred", "conf", "H", "AU", "EU", "MI")
res = []
for off in (0, 1):
    _lib.lib().slu_debug_no_packed_evidential(off)
    res.append(ops.evidential_reduce(x, lab, from_outputs=True, ignore_index=0, want=want))
_lib.lib().slu_debug_no_packed_evidential(0)
a, b = res
for k in want:
    d = (a[k].double() - b[k].double()).abs()
    print("packed-vs-scalar", k, "max abs diff", float(d.max()), "n diff", int((d > 0).sum()), "of", d.numel())
ref_alpha = 1.0 + torch.nn.functional.softplus(x[:, C:C + 1]) * torch.softmax(x[:, :C], dim=1) + 1e-8
for name, r in (("packed", a), ("scalar", b)):
    d = (r["alpha"] - ref_alpha).abs()
    print(name, "alpha vs torch: max", float(d.max()), "n diff", int((d > 0).sum()))
i = (a["alpha"] != b["alpha"]).nonzero()
if i.numel():
    j = tuple(i[0].tolist())
    print("first differing alpha", j, float(a["alpha"][j]), float(b["alpha"][j]), float(ref_alpha[j]))

# (b) Dirichlet AUROC scores
sys.path.insert(0, os.path.join(ROOT))
from tests import tester_cases as tc
from semanticlidarunc_b200.metrics.auroc import AUROCAggregator
from semanticlidarunc_b200.models.probability_helper import to_alpha_concentrations_from_shape_and_scale
batches, outs = tc.make_case("dirichlet")
agg = AUROCAggregator(mode="alpha", score="entropy_norm", ignore_index=0)
S, E = [], []
for (bt, o) in zip(batches, outs):
    o = o.to(dev); labels = bt[4][:, 0].to(dev)
    alpha = to_alpha_concentrations_from_shape_and_scale(o[:, :tc.C], o[:, tc.C:tc.C + 1])
    agg.update(alpha, labels)
    al = alpha.double().cpu()
    a0 = al.sum(1, keepdim=True)
    p = (al / (a0 + 1e-12)).float()
    Hh = -(p.clamp_min(1e-12) * p.clamp_min(1e-12).log()).sum(1) / np.log(tc.C)
    pred = p.argmax(1)
    valid = labels.cpu() != 0
    S.append(Hh[valid].numpy()); E.append((pred != labels.cpu())[valid].numpy())
    score_map, pr = agg._scores_and_pred(alpha)
    print("score diff", float((score_map.cpu() - Hh).abs().max()), "pred diff", int((pr.cpu() != pred).sum()))
S, E = np.concatenate(S), np.concatenate(E)
print("exact AUROC", om.auroc_error_detection(S, E), "hist AUROC", agg.compute()[0], "n", S.size, "err rate", E.mean())
print("score range", S.min(), S.max(), "unique", np.unique(S).size)
