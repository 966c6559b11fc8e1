"""`AUROCAggregator` with the reference's interface (src/metrics/auroc.py:8-164), as a device histogram.

Same constructor `(mode, score, ignore_index, max_samples, seed, eps)`, `update(preds, labels,
score_override=None)`, `compute(save_plot_path, title, dpi)`, `reset()`.  The reference keeps every
(score, is_error) pair on the host and argsorts them at compute(); here `update` is the fused
uncertainty kernel (score map) plus one histogram kernel, and the state is `[2, 2^20]` int64 counts in the
HYBRID bins of slu_score_hist_hybrid (uniform steps of 2^-20 above 1/16, 2048 bins per binary octave below:
entropy-like scores pile up at both ends of [0,1]).

AUROC from the histogram treats the samples of one bin as tied (the reference orders exact ties
arbitrarily); the difference is at most half the probability that a wrong and a correct pixel share a
bin -- below 1e-4 even when a third of the pixels sit within 1e-3 of the maximum entropy, and far below the
sampling noise of the reference's `max_samples` subsample, which is accepted and ignored here (every pixel is
counted).  Any other `n_score_bins` selects a uniform grid of that size.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import _lib, ops

_MODES = {"alpha": ("alpha", ops.CONF_RAW), "logits": ("logits", ops.CONF_RAW), "probs": ("probs", ops.CONF_RENORM)}


def roc_from_hist(hist: np.ndarray, lower_edges=None):
    """(fpr, tpr, thresholds, auroc) for 'high score => error', from [2,M] counts (row 1 = errors).  `lower_edges`: the
    bins' lower edges (default: a uniform grid)."""
    ok, err = hist[0].astype(np.float64), hist[1].astype(np.float64)
    P, N = err.sum(), ok.sum()
    if P == 0 or N == 0:
        return None, None, None, float("nan")
    M = hist.shape[1]
    tps = np.cumsum(err[::-1])
    fps = np.cumsum(ok[::-1])
    tpr = np.concatenate(([0.0], tps / P, [1.0]))
    fpr = np.concatenate(([0.0], fps / N, [1.0]))
    low = np.arange(M) / M if lower_edges is None else np.asarray(lower_edges)[:M]
    thr = np.concatenate(([np.inf], low[::-1], [-np.inf]))
    return fpr, tpr, thr, float(np.trapezoid(tpr, fpr))


class AUROCAggregator:
    def __init__(self, mode="alpha", score="entropy_norm", ignore_index=None, max_samples=None, seed=0, eps=1e-12,
                 n_score_bins: int = ops.HYBRID_BINS):
        assert mode in {"alpha", "logits", "probs"}
        assert score in {"entropy", "entropy_norm", "mi", "mi_norm", "1-maxprob"}
        self.mode, self.score = mode, score
        self.ignore_index = ignore_index
        self.max_samples = max_samples
        self.eps = float(eps)
        self.n_score_bins = int(n_score_bins)
        self.hybrid = self.n_score_bins == ops.HYBRID_BINS
        self._hist = None

    def _accumulator(self, dev=None):
        if self._hist is None:
            self._hist = ops.new_score_hist(_lib.require_cuda(dev), self.n_score_bins)
        return self._hist

    def reset(self):
        if self._hist is not None:
            self._hist.zero_()

    @property
    def _seen(self) -> int:
        return 0 if self._hist is None else int(self._hist.sum().item())

    @torch.no_grad()
    def _scores_and_pred(self, preds: torch.Tensor):
        """(score map in [0,1]-ish, argmax map) exactly as _uncertainty_score / _to_probs (auroc.py:36-63)."""
        C = preds.size(1)
        if self.mode == "alpha" and self.score in {"mi", "mi_norm"}:
            # AUROC depends on ranks only: the un-normalised scores ('mi', 'entropy') are histogrammed as their
            # /log C versions, a monotone map into [0,1]
            r = ops.evidential_reduce(preds, from_outputs=False, eps_metrics=self.eps, normalize=True, want=("MI", "pred"))
            return r["MI"], r["pred"]
        kind, conf_mode = _MODES[self.mode]
        r = ops.reduce_metrics(preds, kind=kind, conf_mode=conf_mode, eps=self.eps, want=("H_norm", "conf", "pred"),
                               normalize=True)
        if self.score == "1-maxprob":
            return 1.0 - r["conf"], r["pred"]
        return r["H_norm"], r["pred"]

    @torch.no_grad()
    def update(self, preds: torch.Tensor, labels: torch.Tensor, score_override: torch.Tensor | None = None):
        assert preds.dim() == 4 and (labels.dim() == 3 or (labels.dim() == 4 and labels.size(1) == 1)), \
            "labels must be [B,H,W] or [B,1,H,W]"
        if labels.dim() == 4:
            labels = labels[:, 0]
        dev = preds.device if preds.is_cuda else _lib.require_cuda()
        preds, labels = preds.to(dev, non_blocking=True), labels.to(dev, non_blocking=True)
        score_map, pred = self._scores_and_pred(preds)
        if score_override is not None:
            score_map = self._unit_range(score_override.to(dev), preds.size(1))
        ops.score_hist(score_map, pred, labels, self._accumulator(dev),
                       ignore=() if self.ignore_index is None else (self.ignore_index,), hybrid=self.hybrid)

    @staticmethod
    def _unit_range(score: torch.Tensor, num_classes: int) -> torch.Tensor:
        """The histogram covers [0,1].  The reference uses an override verbatim and only its RANKS matter
        (src/metrics/auroc.py:124-126), so a score that leaves [0,1] -- an un-normalised entropy or MI, at most
        ln C -- is mapped into it monotonically instead of saturating in the last bin: divide by ln C when that
        suffices, otherwise squash with x / (1 + x).  Negative scores raise (no entropy-like score is negative)."""
        lo, hi = torch.aminmax(torch.nan_to_num(score.detach().float(), nan=0.0))
        lo, hi = float(lo), float(hi)
        if lo < -1e-6:
            raise ValueError(f"score_override has negative values (min {lo}); AUROCAggregator expects a non-negative uncertainty score")
        if hi <= 1.0:
            return score
        if hi <= math.log(num_classes) * (1 + 1e-6):
            return score / math.log(num_classes)
        return score / (1.0 + score)

    @property
    def _scores(self) -> torch.Tensor:
        """Host view for code that reads the reference's per-pixel buffer (the sanity prints of Tester.test_epoch,
        src/models/tester.py:672-677): bin-centre scores, one per counted pixel, thinned proportionally to at most
        `max_samples` (default 1 000 000) entries.  `_is_error` is aligned with it."""
        return self._expanded()[0]

    @property
    def _is_error(self) -> torch.Tensor:
        return self._expanded()[1]

    def _expanded(self):
        if self._hist is None:
            return torch.empty(0), torch.empty(0, dtype=torch.uint8)
        h = self._hist.cpu().numpy()
        cap = int(self.max_samples) if self.max_samples else 1_000_000
        total = int(h.sum())
        if total > cap:
            h = np.floor(h * (cap / total) + 0.5).astype(np.int64)
        edges = ops.hybrid_bin_lower_edges() if self.hybrid else np.arange(h.shape[1] + 1) / h.shape[1]
        centres = (0.5 * (edges[:-1] + edges[1:])).astype(np.float32)
        scores = np.concatenate([np.repeat(centres, h[0]), np.repeat(centres, h[1])])
        err = np.concatenate([np.zeros(int(h[0].sum()), np.uint8), np.ones(int(h[1].sum()), np.uint8)])
        return torch.from_numpy(scores), torch.from_numpy(err)

    def add_maps(self, score_map: torch.Tensor, pred: torch.Tensor, labels: torch.Tensor):
        """Accumulate from maps the fused kernel already produced (no second pass over the class axis).  The score
        must already lie in [0,1] (the kernels' *_norm maps do); nothing is checked here, so the call never synchronises."""
        ops.score_hist(score_map, pred, labels, self._accumulator(score_map.device),
                       ignore=() if self.ignore_index is None else (self.ignore_index,), hybrid=self.hybrid)

    def compute(self, save_plot_path: str | None = None, title: str = "ROC: error detection", dpi: int = 200):
        if self._hist is None or self._seen == 0:
            return float("nan"), {}
        fpr, tpr, thr, auroc = roc_from_hist(self._hist.cpu().numpy(), ops.hybrid_bin_lower_edges() if self.hybrid else None)
        if math.isnan(auroc):
            return auroc, {}
        fig = None
        if save_plot_path is not None:
            try:
                import matplotlib
                matplotlib.use("Agg")
                import matplotlib.pyplot as plt
                fig, ax = plt.subplots(figsize=(6.0, 5.0), dpi=dpi)
                ax.plot([0, 1], [0, 1]); ax.plot(fpr, tpr)
                ax.set_xlim(0, 1); ax.set_ylim(0, 1); ax.set_xlabel("FPR"); ax.set_ylabel("TPR")
                ax.set_title(f"{title}\nAUROC = {auroc:.4f}"); ax.grid(True, alpha=0.3)
                fig.tight_layout(); fig.savefig(save_plot_path, bbox_inches="tight", dpi=dpi); plt.close(fig)
            except Exception:
                fig = None
        return auroc, {"fpr": fpr, "tpr": tpr, "thresholds": thr}, fig
