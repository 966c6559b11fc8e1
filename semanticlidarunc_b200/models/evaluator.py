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
import pandas as pd
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


class UncertaintyAccuracyAggregator:
    """Accuracy per uncertainty bin with the reference's interface (src/models/evaluator.py:640-749):
    `update(labels, preds, uncertainty, ignore_ids=())`, `binned_accuracy(num_bins, bin_width, bin_edges)`,
    `make_bins(...)`, `reset()`.  State is the device error/score histogram ([2, 60000] int64) instead of
    per-pixel host arrays; coarse bins whose edges are multiples of 1/60000 (10, 15, 20, 50, 100 bins ...)
    are exact, other edges are exact up to the pixels inside the one fine bin that straddles each edge.
    `max_samples` is accepted and ignored (every pixel is counted)."""

    def __init__(self, max_samples: int | None = None, seed: int = 0, n_score_bins: int = ops.SCORE_BINS):
        self.max_samples = max_samples
        self.n_score_bins = int(n_score_bins)
        self._hist = None

    def reset(self):
        if self._hist is not None:
            self._hist.zero_()

    @property
    def _seen(self) -> int:
        return 0 if self._hist is None else int(self._hist.sum().item())

    @torch.no_grad()
    def update(self, labels: torch.Tensor, preds: torch.Tensor, uncertainty: torch.Tensor, ignore_ids=()):
        assert labels.shape == preds.shape == uncertainty.shape, "shapes must match"
        dev = uncertainty.device if uncertainty.is_cuda else _lib.require_cuda()
        if self._hist is None:
            self._hist = ops.new_score_hist(dev, self.n_score_bins)
        ops.score_hist(uncertainty.detach().to(dev), preds.detach().to(dev), labels.detach().to(dev), self._hist,
                       ignore=tuple(ignore_ids))

    def make_bins(self, num_bins: int | None = None, bin_width: float | None = None, bin_edges=None) -> np.ndarray:
        if bin_edges is not None:
            edges = np.asarray(bin_edges, dtype=np.float32)
        elif bin_width is not None:
            edges = np.linspace(0.0, 1.0, max(1, int(round(1.0 / float(bin_width)))) + 1, dtype=np.float32)
        else:
            edges = np.linspace(0.0, 1.0, (int(num_bins) if num_bins is not None else 10) + 1, dtype=np.float32)
        edges[0] = 0.0
        edges[-1] = 1.0
        assert np.all(np.diff(edges) > 0), "bin edges must be strictly increasing"
        return edges

    def binned_accuracy(self, num_bins: int = 10, bin_width: float | None = None, bin_edges=None) -> pd.DataFrame:
        if self._hist is None or self._seen == 0:
            return pd.DataFrame(columns=["low", "high", "label", "n", "pct", "accuracy"])
        h = self._hist.cpu().numpy()
        edges = self.make_bins(num_bins=num_bins, bin_width=bin_width, bin_edges=bin_edges)
        M = h.shape[1]
        starts = np.arange(M, dtype=np.float64) / M                       # lower edge of every fine bin
        coarse = np.searchsorted(edges.astype(np.float64), starts + 0.5 / M, side="right") - 1
        coarse = np.clip(coarse, 0, len(edges) - 2)
        n = np.bincount(coarse, weights=(h[0] + h[1]).astype(np.float64), minlength=len(edges) - 1)
        c = np.bincount(coarse, weights=h[0].astype(np.float64), minlength=len(edges) - 1)
        acc = np.divide(c, n, out=np.full_like(c, np.nan, dtype=float), where=n > 0)
        lows, highs = edges[:-1], edges[1:]
        labels = [f"[{l:.2f}, {hh:.2f})" if i < len(lows) - 1 else f"[{l:.2f}, {hh:.2f}]" for i, (l, hh) in enumerate(zip(lows, highs))]
        return pd.DataFrame({"low": lows, "high": highs, "label": labels, "n": n.astype(int),
                             "pct": 100.0 * n / max(1.0, n.sum()), "accuracy": acc})


class UncertaintyPerClassAggregator:
    """Per-class uncertainty statistics with the reference's interface (src/models/evaluator.py:191-281):
    `update(labels, uncertainty)`, `reset()`, `as_dataframe(class_names, ignore_ids)`, `_seen_counts`.
    State is a device histogram [C, 2048] plus an exact per-class sum, so the per-class mean (what
    plot_iou_sorted_by_uncertainty uses, :559-563) is exact and the distribution plots see 2048-bin densities.
    `as_dataframe` expands the histogram to bin-centre samples (at most `max_rows_per_class` per class, in
    proportion); `class_stats()` gives count / mean / quartiles directly.  `max_per_class` is accepted and ignored."""

    def __init__(self, num_classes: int, max_per_class: int | None = None, seed: int = 0, n_score_bins: int = 2048,
                 max_rows_per_class: int = 200_000):
        self.num_classes = int(num_classes)
        self.max_per_class = max_per_class
        self.n_score_bins = int(n_score_bins)
        self.max_rows_per_class = int(max_rows_per_class)
        self._hist = None
        self._sum = None

    def reset(self):
        if self._hist is not None:
            self._hist.zero_()
            self._sum.zero_()

    @property
    def _seen_counts(self):
        if self._hist is None:
            return [0] * self.num_classes
        return [int(v) for v in self._hist.sum(dim=1).cpu()]

    @torch.no_grad()
    def update(self, labels: torch.Tensor, uncertainty: torch.Tensor):
        assert labels.shape == uncertainty.shape, "labels and uncertainty must have same shape"
        dev = uncertainty.device if uncertainty.is_cuda else _lib.require_cuda()
        if self._hist is None:
            self._hist = torch.zeros((self.num_classes, self.n_score_bins), dtype=torch.int64, device=dev)
            self._sum = torch.zeros((self.num_classes,), dtype=torch.int64, device=dev)
        ops.class_score_hist(uncertainty.detach().to(dev), labels.detach().to(dev), self._hist, self._sum)

    def class_stats(self) -> pd.DataFrame:
        cols = ["class_id", "n", "mean", "q25", "median", "q75"]
        if self._hist is None:
            return pd.DataFrame(columns=cols)
        h = self._hist.cpu().numpy().astype(np.float64)
        sums = self._sum.cpu().numpy().astype(np.float64) / 4294967296.0
        centres = (np.arange(self.n_score_bins) + 0.5) / self.n_score_bins
        rows = []
        for c in range(self.num_classes):
            n = h[c].sum()
            if n == 0:
                continue
            cdf = np.cumsum(h[c]) / n
            q = [float(centres[min(np.searchsorted(cdf, t), self.n_score_bins - 1)]) for t in (0.25, 0.5, 0.75)]
            rows.append({"class_id": c, "n": int(n), "mean": float(sums[c] / n), "q25": q[0], "median": q[1], "q75": q[2]})
        return pd.DataFrame(rows, columns=cols)

    def as_dataframe(self, class_names, ignore_ids=()):
        """Long DataFrame with columns class_id, class, uncertainty (bin-centre samples of the histogram)."""
        if self._hist is None:
            return pd.DataFrame(columns=["class_id", "class", "uncertainty"])
        h = self._hist.cpu().numpy()
        centres = ((np.arange(self.n_score_bins) + 0.5) / self.n_score_bins).astype(np.float32)
        rows, skip = [], set(ignore_ids)
        for c in range(self.num_classes):
            n = int(h[c].sum())
            if c in skip or n == 0:
                continue
            reps = h[c] if n <= self.max_rows_per_class else np.round(h[c] * (self.max_rows_per_class / n)).astype(np.int64)
            rows.append(pd.DataFrame({"class_id": c, "class": class_names[c], "uncertainty": np.repeat(centres, reps)}))
        if not rows:
            return pd.DataFrame(columns=["class_id", "class", "uncertainty"])
        return pd.concat(rows, ignore_index=True)
