"""The reference's OWN CPU code for the hot path, timed on the host cores.  TEST INFRASTRUCTURE ONLY (tests + bench baseline).

Runs the unmodified reference modules out of `oracle/_ref/reference_src.zip` (built by oracle/make_ref.sh from
/root/reference, git-ignored, travels to the GPU box) or, inside the build container, from /root/reference/src
directly.  Only bench.py's `--impl reference` / `cpu_baseline` legs and tests/ import this module; the product
package never does.

One scan of the workload goes through the stock reference code path:

  SemanticKitti.__getitem__            src/dataset/dataloader_semantic_KITTI.py:31-99  (np.fromfile, id_map loop,
                                        spherical_projection src/dataset/utils.py:288-349, fp32 range, build_normal_xyz)
  MC block of Tester.test_epoch        src/models/tester.py:405-471  (softmax, mean, argmax and the two nested
                                        closures mc_predictive_entropy_norm / mc_mutual_information_norm, which are
                                        compiled unmodified out of the reference source with `ast`)
  IoUEvaluator.update                  src/models/evaluator.py:39-53
  ECEAggregator.update                 src/metrics/ece.py:66-111 (mode "probs", ignore_index 0, max_samples 1 000 000:
                                        the Tester's own construction, src/models/tester.py:176-181)
  label back-projection                the reference has NO such function; the definition
                                        point_label[n] = pred[row[n], col[n]] is evaluated with oracle.projection

The loader runs the way the reference runs it: a torch DataLoader(batch_size=1, num_workers=k) over the Dataset
(src/train_semantics.py:115-127; the shipped configs use 8-16 workers); `workers=0` is the single-process form.
"""
from __future__ import annotations

import ast
import math
import os
import sys
import tempfile
import time
import zipfile
from unittest.mock import MagicMock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_ZIP = os.path.join(HERE, "_ref", "reference_src.zip")
REF_LIVE = os.path.join(os.environ.get("SLU_REFERENCE_ROOT", "/root/reference"), "src")

_STUBS = ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker", "matplotlib.patheffects", "matplotlib.colors",
          "matplotlib.cm", "matplotlib.patches", "matplotlib.lines", "matplotlib.gridspec", "seaborn", "open3d")

# top-level package names of the reference (src/ is put on sys.path, as the reference's own entry points do)
REF_PACKAGES = ("dataset", "metrics", "losses", "models", "utils")


def reference_location():
    """(kind, sys.path entry) of the unmodified reference, or (None, None)."""
    if os.path.isfile(REF_ZIP):
        return "zip", REF_ZIP + "/src"
    if os.path.isdir(REF_LIVE):
        return "tree", REF_LIVE
    return None, None


def available() -> bool:
    return reference_location()[0] is not None


def read_reference_source(rel: str) -> str:
    """Text of src/<rel> of the reference (from the archive or the live tree)."""
    kind, _ = reference_location()
    if kind == "zip":
        with zipfile.ZipFile(REF_ZIP) as z:
            return z.read("src/" + rel).decode()
    with open(os.path.join(REF_LIVE, rel)) as f:
        return f.read()


def install():
    """Make `import dataset.utils`, `import models.tester` ... resolve to the unmodified reference.  The plotting
    libraries it imports at module top (absent in this image) are stubbed; nothing else is."""
    kind, entry = reference_location()
    if kind is None:
        raise RuntimeError("no reference available: run oracle/make_ref.sh where /root/reference exists")
    for m in _STUBS:
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:
                sys.modules[m] = MagicMock()
    plt = sys.modules["matplotlib.pyplot"]
    if isinstance(plt, MagicMock):
        plt.subplots.return_value = (MagicMock(), MagicMock())
        sys.modules["matplotlib"].pyplot = plt
    # the oracle/ directory itself must not shadow the reference's `metrics`, `losses` packages
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != HERE]
    if entry not in sys.path:
        sys.path.insert(0, entry)
    return kind


def extract_closures(rel: str, names):
    """Compile nested function definitions out of a reference source file, unmodified (decorators dropped:
    they are @torch.no_grad() only)."""
    tree = ast.parse(read_reference_source(rel))
    found = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names and node.name not in found:
            node.decorator_list = []
            found[node.name] = node
    mod = ast.Module(body=[found[n] for n in names], type_ignores=[])
    ast.fix_missing_locations(mod)
    ns = {"torch": torch, "math": math}
    exec(compile(mod, rel, "exec"), ns)
    return [ns[n] for n in names]


def write_kitti_files(scans, directory):
    """Synthetic scans -> SemanticKITTI .bin / .label pairs (what the reference Dataset reads)."""
    paths = []
    for i, (xyzi, raw) in enumerate(scans):
        b = os.path.join(directory, "%06d.bin" % i)
        l = os.path.join(directory, "%06d.label" % i)
        np.ascontiguousarray(xyzi, dtype=np.float32).tofile(b)
        np.ascontiguousarray(raw, dtype=np.uint32).tofile(l)
        paths.append((b, l))
    return paths


class ReferenceArm:
    """Batches of scans through the stock reference code on the CPU."""

    def __init__(self, scans, logits_per_scan, *, H, W, C, n_bins=15, workers=0, total_items=None):
        self.kind = install()
        from dataset.dataloader_semantic_KITTI import SemanticKitti          # the reference's classes
        from metrics.ece import ECEAggregator
        from models.evaluator import IoUEvaluator
        from torch.utils.data import DataLoader
        import torch.nn.functional as F

        self.F = F
        self.H, self.W, self.C = H, W, C
        self.entropy_norm, self.mi_norm = extract_closures(
            "models/tester.py", ["mc_predictive_entropy_norm", "mc_mutual_information_norm"])
        self.tmp = tempfile.TemporaryDirectory(prefix="slu_ref_")
        self.paths = write_kitti_files(scans, self.tmp.name)
        self.scans = scans
        self.logits = logits_per_scan                                         # list of [T,1,C,H,W] host tensors
        n = len(scans)
        total = n if total_items is None else int(total_items)
        data_path = [self.paths[i % n] for i in range(total)]
        self.ds = SemanticKitti(data_path, rotate=False, flip=False, projection=(H, W), resize=False)
        self.workers = int(workers)
        kw = dict(batch_size=1, shuffle=False, num_workers=self.workers)      # src/train_semantics.py:127
        if self.workers > 0:
            kw.update(persistent_workers=True, prefetch_factor=2)
        self.loader = DataLoader(self.ds, **kw)
        self.it = iter(self.loader)
        self.cursor = 0
        self.iou = IoUEvaluator(C)                                            # src/models/tester.py:157
        self.ece = ECEAggregator(n_bins=n_bins, mode="probs", ignore_index=0, max_samples=1_000_000)   # :176-181
        self.last = None

    def scan_pass(self):
        """The next scan of the list through loader -> MC block -> metrics -> back-projection."""
        from oracle import projection as oproj
        i = self.cursor % len(self.scans)
        self.cursor += 1
        range_img, reflectivity, xyz, normals, labels = next(self.it)         # reference Dataset item, batch of 1
        labels = labels.squeeze(1).long() if (labels.ndim == 4 and labels.shape[1] == 1) else labels.long()   # tester.py:397-400
        mc_outputs = self.logits[i]                                           # [T,1,C,H,W]: what mc_forward returns (:407)
        probs = self.F.softmax(mc_outputs, dim=2)                             # :413
        p_bar = probs.mean(dim=0)                                             # :416
        preds = p_bar.argmax(dim=1)                                           # :418
        H_norm = self.entropy_norm(probs)                                     # :454
        self.iou.update(preds, labels)                                        # :458
        self.ece.update(p_bar, labels)                                        # :466
        MI_norm = self.mi_norm(probs)                                         # :470
        # label back-projection (definition): every point reads the predicted label of its pixel
        xyzi = self.scans[i][0]
        row, col, _ = oproj.projection_indices(xyzi.astype(np.float64), self.H, self.W, None)
        point_labels = preds[0].numpy()[row, col]
        self.last = {"pred": preds, "H_norm": H_norm, "MI_norm": MI_norm, "p_bar": p_bar, "point_labels": point_labels,
                     "range": range_img, "normals": normals, "labels": labels}
        return self.last

    def finish(self):
        """mIoU / ECE of everything accumulated (the sweep's end: evaluator.py:62-105, ece.py:113-168)."""
        from dataset.definitions import class_names
        miou, _ = self.iou.compute(class_names=class_names, test_mask=[0] + [1] * (self.C - 1), ignore_gt=[0],
                                   reduce="mean", ignore_th=None)
        (ece, mce), _, _ = self.ece.compute(save_plot_path=os.path.join(self.tmp.name, "ece.png"))
        return {"mIoU": float(miou), "ece": float(ece), "mce": float(mce), "confmat_sum": int(self.iou.confmat.sum())}

    def close(self):
        try:
            del self.it
            del self.loader
        except Exception:
            pass
        self.tmp.cleanup()


def time_reference(scans, logits_per_scan, *, H, W, C, steps, warmup, scans_per_step, workers, budget_s=None):
    """Timed run: `warmup` + `steps` steps of `scans_per_step` scan passes.  Returns a dict with scans/s."""
    total = (steps + warmup) * scans_per_step
    arm = ReferenceArm(scans, logits_per_scan, H=H, W=W, C=C, workers=workers, total_items=total)
    try:
        for _ in range(warmup * scans_per_step):
            arm.scan_pass()
        done_steps, t0 = 0, time.perf_counter()
        for _ in range(steps):
            for _ in range(scans_per_step):
                arm.scan_pass()
            done_steps += 1
            if budget_s is not None and time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
        res = arm.finish()
        res.update({"scans_per_s": done_steps * scans_per_step / dt, "seconds": dt, "steps": done_steps,
                    "kind": arm.kind, "workers": workers})
        return res
    finally:
        arm.close()
