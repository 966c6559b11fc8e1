"""Drive the reference's OWN evaluation loop, `Tester.test_epoch` (src/models/tester.py:272-720), with a stub model.
TEST INFRASTRUCTURE ONLY.

Two modes, same unmodified loop:
  swap=False  the stock reference: its IoUEvaluator / ECEAggregator / AUROCAggregator / uncertainty aggregators and
              probability_helper, exactly as shipped;
  swap=True   the INTEGRATION.md section A import swap, done without touching a line of the reference: before
              `models.tester` is imported, sys.modules is pre-seeded so that `models.evaluator`, `metrics.ece`,
              `metrics.auroc`, `models.probability_helper` and `utils.mc_dropout` resolve to the
              semanticlidarunc_b200 mirrors (needs a GPU: the mirrors have no CPU path).

The stub model replays precomputed head outputs, so both runs see identical logits; the loader is a list of
(range, reflectivity, xyz, normals, labels) batches.  Returns the loop's results read back through the public
surfaces the reference itself uses (`.confmat`, `compute()`, result_dict.json).
"""
from __future__ import annotations

import json
import os
import sys

import torch

from . import ref_arm

SWAPS = {
    "models.evaluator": "semanticlidarunc_b200.models.evaluator",
    "metrics.ece": "semanticlidarunc_b200.metrics.ece",
    "metrics.auroc": "semanticlidarunc_b200.metrics.auroc",
    "models.probability_helper": "semanticlidarunc_b200.models.probability_helper",
    "utils.mc_dropout": "semanticlidarunc_b200.utils.mc_dropout",
}


def _purge_reference_modules():
    for name in list(sys.modules):
        top = name.split(".")[0]
        if top in ref_arm.REF_PACKAGES and not name.startswith("semanticlidarunc_b200"):
            mod = sys.modules[name]
            f = getattr(mod, "__file__", None) or ""
            spec_paths = str(getattr(mod, "__path__", ""))
            if "reference" in f or "reference" in spec_paths or f == "" or name in SWAPS:
                del sys.modules[name]


class ReplayModel(torch.nn.Module):
    """forward() returns the next precomputed output; has one dropout layer so mc_forward has something to toggle."""

    def __init__(self, outputs):
        super().__init__()
        self.drop = torch.nn.Dropout(0.5)
        self.dummy = torch.nn.Parameter(torch.zeros(1))
        self.outputs = outputs            # list of [B,Cout,H,W] tensors, consumed in call order
        self.calls = 0

    def forward(self, *inputs):
        y = self.outputs[self.calls % len(self.outputs)]
        self.calls += 1
        return y.to(self.dummy.device)


class _NullEvent:
    def record(self, *a, **k):
        pass

    def elapsed_time(self, other):
        return 0.0


class ListLoader(list):
    batch_size = 1


def make_cfg(branch: str, num_classes: int, T: int):
    ref_arm.install()
    from dataset.definitions import class_names, color_map
    names = [class_names[k] for k in sorted(class_names)][:num_classes]
    cfg = {
        "model_settings": {"use_mc_sampling": 1 if branch == "mc" else 0, "mc_samples": int(T),
                           "loss_function": "Dirichlet" if branch == "dirichlet" else "CrossEntropy",
                           "baseline": "SalsaNext", "reflectivity": 0, "normals": 0},
        "extras": {"num_classes": num_classes + 1 if branch == "dirichlet" else num_classes,
                   "class_names": names, "class_colors": color_map},
        "train_params": {"batch_size": 1},
    }
    if branch == "dirichlet":
        # the default mask is built from extras.num_classes (C+1 for the Dirichlet head, tester.py:159-162) and would not
        # match the C-class evaluator: the shipped configs carry an explicit mask, so does this one
        cfg["extras"]["test_mask"] = {i: (0 if i == 0 else 1) for i in range(num_classes)}
    return cfg


def run_tester(branch: str, swap: bool, batches, model_outputs, workdir: str, num_classes: int = 20, T: int = 4):
    """branch: "mc" | "dirichlet".  batches: list of 5-tuples (CPU tensors).  model_outputs: list of head outputs in
    the order the loop will call the model (T per batch for "mc").  Returns a dict of results."""
    _purge_reference_modules()
    ref_arm.install()
    if swap:
        import importlib
        for ref_name, ours in SWAPS.items():
            sys.modules[ref_name] = importlib.import_module(ours)
    try:
        from models.tester import Tester
        cfg = make_cfg(branch, num_classes, T)
        model = ReplayModel(model_outputs)
        # logging=True: the MC branch reads self._start.elapsed_time(self._end) unconditionally (tester.py:474), which
        # only works when the timers were recorded, i.e. with logging on and CUDA present
        tester = Tester(model, cfg, visualize=False, logging=True, checkpoint=None)
        if not torch.cuda.is_available():                 # CPU-only container: stand-in events (instance attributes only)
            tester._use_cuda_events = True
            tester._start = tester._end = _NullEvent()
            torch_sync, torch.cuda.synchronize = torch.cuda.synchronize, (lambda *a, **k: None)
        os.makedirs(workdir, exist_ok=True)
        tester.checkpoint = os.path.join(workdir, "model_3.pt")       # test_epoch derives its output folders from it
        loader = ListLoader(batches)
        tester.test_epoch(loader, epoch=0)
        with open(os.path.join(workdir, "test", "result_dict.json")) as f:
            result = json.load(f)
        out = {"classes": {k: type(v).__module__ for k, v in (("iou", tester.iou_evaluator), ("ece", tester.ece_eval),
                                                               ("auroc", tester.auroc_eval), ("ua", tester.ua_agg),
                                                               ("unc", tester.unc_agg))},
               "confmat": tester.iou_evaluator.confmat.detach().cpu().clone(),
               "mIoU": result["mIoU"], "iou": result["iou"], "model_calls": model.calls}
        (ece, mce), stats, _ = tester.ece_eval.compute(save_plot_path=os.path.join(workdir, "ece.png"))
        out["ece"], out["mce"] = float(ece), float(mce)
        out["ece_bin_n"] = [int(v) for v in stats["n"]] if "n" in stats else None
        r = tester.auroc_eval.compute(save_plot_path=os.path.join(workdir, "roc.png"))
        out["auroc"] = float(r[0])
        r = tester.auroc_eval_mi.compute(save_plot_path=os.path.join(workdir, "roc_mi.png"))
        out["auroc_mi"] = float(r[0])
        ua = tester.ua_agg.binned_accuracy(bin_width=0.05)
        out["ua_n"] = [int(v) for v in ua["n"]]
        out["ua_acc"] = [float(v) for v in ua["accuracy"]]
        out["unc_seen"] = [int(v) for v in tester.unc_agg._seen_counts]
        out["summary_saved"] = os.path.isfile(os.path.join(workdir, "outputs_summary", "summary_epoch_003.pt"))
        return out
    finally:
        if not torch.cuda.is_available() and "torch_sync" in locals():
            torch.cuda.synchronize = torch_sync
        _purge_reference_modules()
