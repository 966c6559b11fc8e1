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
        labels = labels.detach().to(dev)
        ignore = tuple(int(v) for v in ignore_ids)
        if len(ignore) > 4:                         # the kernel takes four ids: fold longer lists into the score (NaN = skipped)
            drop = torch.isin(labels, torch.tensor(ignore, device=dev, dtype=labels.dtype))
            uncertainty = torch.where(drop, torch.full_like(uncertainty, float("nan")), uncertainty.to(dev))
            ignore = ()
        ops.score_hist(uncertainty.detach().to(dev), preds.detach().to(dev), labels, self._hist, ignore=ignore)

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


    def plot_accuracy_vs_uncertainty_bins(self, num_bins: int = 10, bin_width: float | None = None, bin_edges=None,
                                          figsize=(14, 5), title="Pixel Accuracy vs Predictive-Uncertainty (binned)",
                                          x_label="Normalized predictive-entropy bin", y_label="Accuracy",
                                          show_percent_on_bars: bool = True, annotate_min_pct: float = 0.1,
                                          annotate_every: int = 1, percent_fmt: str = "{:.1f}%", save_path: str | None = None,
                                          show: bool = False, close_fig: bool = True, dpi: int = 200,
                                          cmap_name: str = "viridis", color_norm: str = "linear"):
        """Same call as src/models/evaluator.py:752-870.  The statistics come from the device histogram; the bar chart
        is drawn only when matplotlib is installed (host-side plotting is not part of the hot path).  Returns the stats."""
        stats = self.binned_accuracy(num_bins=num_bins, bin_width=bin_width, bin_edges=bin_edges)
        if stats.empty or stats["n"].sum() == 0:
            print("No data to plot.")
            return None
        plt = _pyplot()
        if plt is not None and save_path:
            fig, ax = plt.subplots(figsize=figsize, dpi=dpi)
            ax.bar(np.arange(len(stats)), stats["accuracy"].fillna(0.0).to_numpy(), width=0.9)
            ax.set_xticks(np.arange(len(stats)))
            ax.set_xticklabels(stats["label"], rotation=45, ha="right")
            if show_percent_on_bars:
                for i, (a, pc) in enumerate(zip(stats["accuracy"], stats["pct"])):
                    if pc >= annotate_min_pct and i % max(1, annotate_every) == 0 and a == a:
                        ax.text(i, a, percent_fmt.format(pc), ha="center", va="bottom", fontsize=8)
            ax.set_ylim(0, 1); ax.set_xlabel(x_label); ax.set_ylabel(y_label); ax.set_title(title)
            fig.tight_layout(); fig.savefig(save_path, dpi=dpi, bbox_inches="tight")
            if close_fig:
                plt.close(fig)
        return stats


def _pyplot():
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
        return plt
    except Exception:
        return None


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

    def density(self, bins: int = 2048, bandwidth="silverman"):
        """Per-class smoothed densities on a [0,1] grid: the histogram convolved with a Gaussian whose width follows
        Silverman's / Scott's rule (or a float in x units), reflected at both ends -- what plot_ridgeline_fast draws
        (src/models/evaluator.py:413-538).  Returns (grid [bins], {class_id: density [bins]})."""
        grid = (np.arange(bins) + 0.5) / bins
        if self._hist is None:
            return grid, {}
        h = self._hist.cpu().numpy().astype(np.float64)
        M = self.n_score_bins
        centres = (np.arange(M) + 0.5) / M
        stats = self.class_stats().set_index("class_id")
        out = {}
        for c in range(self.num_classes):
            n = h[c].sum()
            if n == 0:
                continue
            mean = float(stats.loc[c, "mean"])
            std = float(np.sqrt(max(1e-12, (h[c] * (centres - mean) ** 2).sum() / n)))
            if isinstance(bandwidth, str):
                iqr = float(stats.loc[c, "q75"] - stats.loc[c, "q25"])
                sig = min(std, iqr / 1.349) if iqr > 0 else std
                bw = (0.9 if bandwidth == "silverman" else 1.059) * sig * n ** (-0.2)
            else:
                bw = float(bandwidth)
            bw = max(bw, 1.0 / bins)
            coarse = np.bincount(np.minimum((centres * bins).astype(np.int64), bins - 1), weights=h[c], minlength=bins)
            half = int(min(bins, np.ceil(4 * bw * bins)))
            k = np.exp(-0.5 * ((np.arange(-half, half + 1) / bins) / bw) ** 2)
            k /= k.sum()
            padded = np.concatenate([coarse[:half][::-1], coarse, coarse[-half:][::-1]]) if half > 0 else coarse
            dens = np.convolve(padded, k, mode="same")[half:half + bins] if half > 0 else coarse
            out[c] = dens / max(dens.sum() / bins, 1e-30)
        return grid, out

    def plot_ridgeline_fast(self, class_names: list, color_map: dict, ignore_ids=(), figsize=(14, 9),
                            title="Normalized Uncertainty per Class (Ridgeline)", x_label="Normalized uncertainty",
                            bins: int = 2048, bandwidth="silverman", fill_alpha: float = 0.9, line_width: float = 1.0,
                            save_path: str | None = None, dpi: int = 200):
        """Same call as src/models/evaluator.py:413-538; densities from the device histogram, drawn only when matplotlib
        is installed.  Returns (grid, densities)."""
        grid, dens = self.density(bins=min(int(bins), 8192), bandwidth=bandwidth)
        ids = [c for c in dens if c not in set(ignore_ids)]
        if not ids:
            print("No data to plot.")
            return grid, {}
        plt = _pyplot()
        if plt is not None and save_path:
            fig, ax = plt.subplots(figsize=figsize, dpi=dpi)
            for row, c in enumerate(ids):
                y = dens[c] / max(dens[c].max(), 1e-30) * 0.9
                col = np.array(color_map[c]) / 255.0
                ax.fill_between(grid, row, row + y, color=col, alpha=fill_alpha, linewidth=line_width)
            ax.set_yticks(np.arange(len(ids)) + 0.2); ax.set_yticklabels([class_names[c] for c in ids])
            ax.set_xlim(0, 1); ax.set_xlabel(x_label); ax.set_title(title)
            fig.tight_layout(); fig.savefig(save_path, dpi=dpi, bbox_inches="tight"); plt.close(fig)
        return grid, {c: dens[c] for c in ids}


def plot_iou_sorted_by_uncertainty(unc_agg, result_dict: dict, class_names: list, color_map: dict, ignore_ids=(0,),
                                   figsize=(18, 6), title="mIoU per class, sorted by mean uncertainty", y_label="mIoU",
                                   save_path: str | None = None, dpi: int = 200):
    """Same call as src/models/evaluator.py:546-630: per-class IoU ordered by the class's mean uncertainty (exact means from
    the aggregator's fixed-point sums).  Returns the ordered DataFrame; draws only when matplotlib is installed."""
    st = unc_agg.class_stats()
    st = st[~st["class_id"].isin(set(ignore_ids))].copy()
    if st.empty:
        print("No data to plot.")
        return st
    st["class"] = [class_names[int(c)] for c in st["class_id"]]
    st["iou"] = [float(result_dict.get(n, float("nan"))) for n in st["class"]]
    st = st.sort_values("mean").reset_index(drop=True)
    plt = _pyplot()
    if plt is not None and save_path:
        fig, ax = plt.subplots(figsize=figsize, dpi=dpi)
        ax.bar(np.arange(len(st)), st["iou"], color=[np.array(color_map[int(c)]) / 255.0 for c in st["class_id"]])
        ax.set_xticks(np.arange(len(st))); ax.set_xticklabels(st["class"], rotation=45, ha="right")
        ax.set_ylabel(y_label); ax.set_title(title)
        fig.tight_layout(); fig.savefig(save_path, dpi=dpi, bbox_inches="tight"); plt.close(fig)
    return st
