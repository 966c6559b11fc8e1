"""`IoUEvaluator` with the reference's interface, accumulated on the GPU.

Mirrors src/models/evaluator.py:29-105: `IoUEvaluator(num_classes, device="cpu")`, `.update(preds,
targets)`, `.reset()`, `.compute(class_names, test_mask, ignore_gt, reduce, ignore_th) -> (mIoU,
dict)` and the public `.confmat` [C,C] int64 (rows = GT, cols = prediction).

The reference moves preds/targets to the CPU every batch and bincounts there; here `update` is one
kernel launch (slu_confusion_ece) into a device-resident int64 matrix and nothing is copied to the
host until `.confmat` / `.compute()` is read.  `device` only says where `.confmat` is presented.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib, ops


class IoUEvaluator:
    def __init__(self, num_classes: int, device="cpu"):
        self.C = int(num_classes)
        self.device = device
        self._acc = None            # device accumulator, created on first use
        self.reset()

    # -- state ---------------------------------------------------------------------------------
    def _accumulator(self, dev=None) -> torch.Tensor:
        if self._acc is None:
            dev = _lib.require_cuda(dev)
            self._acc = ops.new_confmat(self.C, dev)
        return self._acc

    def reset(self):
        if self._acc is not None:
            self._acc.zero_()

    @property
    def confmat(self) -> torch.Tensor:
        if self._acc is None:
            return torch.zeros((self.C, self.C), dtype=torch.long, device=self.device)
        return self._acc.to(self.device)

    @confmat.setter
    def confmat(self, value: torch.Tensor):
        # the reference Tester restores a cached matrix by assignment (src/models/tester.py:330)
        value = torch.as_tensor(value, dtype=torch.long)
        if tuple(value.shape) != (self.C, self.C):
            raise ValueError(f"confmat must be [{self.C},{self.C}]")
        self._accumulator(value.device if value.is_cuda else None).copy_(value)

    # -- accumulation --------------------------------------------------------------------------
    @torch.no_grad()
    def update(self, preds: torch.Tensor, targets: torch.Tensor):
        """preds/targets: [B,H,W] integer class ids; pairs outside [0,C) are dropped (evaluator.py:49)."""
        dev = preds.device if preds.is_cuda else (targets.device if targets.is_cuda else _lib.require_cuda())
        preds = preds.to(dev, non_blocking=True)
        targets = targets.to(dev, non_blocking=True)
        ops.confusion_ece(preds, targets, None, num_classes=self.C, confmat=self._accumulator(dev))

    # -- result --------------------------------------------------------------------------------
    def compute(self, class_names, test_mask=None, ignore_gt=None, reduce="mean", ignore_th=None):
        """400 integers -> IoU: host arithmetic on the finished matrix, as evaluator.py:62-105."""
        cm = self.confmat.to("cpu").double()
        if ignore_gt:
            rows = torch.tensor(ignore_gt, dtype=torch.long)
            rows = rows[(rows >= 0) & (rows < self.C)]
            cm[rows, :] = 0.0
        tp = cm.diag()
        denom = cm.sum(0) + cm.sum(1) - tp              # TP + FP + FN
        iou = torch.full((self.C,), float("nan"), dtype=torch.float64)
        ok = denom > 0
        iou[ok] = tp[ok] / denom[ok]
        if test_mask is None:
            mask = torch.ones(self.C, dtype=torch.bool)
        else:
            mask = torch.as_tensor(test_mask, dtype=torch.bool)
            if mask.numel() != self.C:
                raise ValueError("test_mask length != num_classes")
        mask = mask & torch.isfinite(iou)
        if ignore_th is not None:
            mask = mask & (iou >= ignore_th)
        out = {}
        for k in range(self.C):
            name = class_names[k] if isinstance(class_names, (list, dict)) else class_names[str(k)]
            out[name] = float(iou[k]) if torch.isfinite(iou[k]) else float("nan")
        if mask.any():
            vals = iou[mask].numpy()
            miou = float(np.mean(vals)) if reduce == "mean" else float(np.median(vals))
        else:
            miou = float("nan")
        out["mIoU"] = miou
        return miou, out
