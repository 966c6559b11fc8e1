"""`ECEAggregator` with the reference's interface, as a streaming device histogram.

Mirrors src/metrics/ece.py:13-212: ctor `(n_bins=15, mode in {alpha,logits,probs}, ignore_index,
max_samples, seed, eps, binning, plot_style)`, `.update(preds[B,C,H,W], labels[B,H,W])`,
`.compute(save_plot_path, title, dpi)`, `.reset()`.

The reference appends (confidence, correct) of every valid pixel to host tensors and histograms them
at compute(); here `update` is one fused kernel (softmax / normalise -> top-label confidence ->
bin) that adds into 3*n_bins int64 counters on the GPU: n, n_correct and sum(conf) in 2^-32 fixed
point.  Differences, all deliberate:
  * every sample is counted; `max_samples` (a host-memory cap implemented by random subsampling,
    ece.py:93-111) is accepted and ignored, so results equal the reference with max_samples=None;
  * the per-bin sums are exact integers, where np.histogram accumulates float32 weights
    (ece.py:136-138); ECE agrees to ~1e-6 relative, not bit for bit;
  * binning="adaptive" (equal-mass edges) reads its quantile edges off a fine [2,60000] confidence histogram kept
    beside the counters, i.e. at a confidence resolution of 1/60000 instead of from every stored sample;
  * compute(save_plot_path=None) returns fig=None instead of raising UnboundLocalError (ece.py:212).
The private `_conf/_correct` arrays the reference Tester caches (src/models/tester.py:334-337) do
not exist; `state_dict()/load_state_dict()` carry the counters instead.
"""
from __future__ import annotations

import numpy as np
import pandas as pd
import torch

from .. import _lib, ops

_MODES = {"alpha": ("alpha", ops.CONF_RAW), "logits": ("logits", ops.CONF_RAW), "probs": ("probs", ops.CONF_RENORM)}


class ECEAggregator:
    def __init__(self, n_bins=15, mode="alpha", ignore_index=None, max_samples=None, seed=0, eps=1e-12,
                 binning: str = "uniform", plot_style: str = "classic"):
        assert binning in {"uniform", "adaptive"}
        assert plot_style in {"classic", "classic+hist", "gap"}
        assert mode in {"alpha", "logits", "probs"}
        assert n_bins >= 2
        if n_bins > _lib.MAX_BINS:
            raise ValueError(f"n_bins={n_bins} > {_lib.MAX_BINS}")
        self.n_bins = int(n_bins)
        self.mode = mode
        self.ignore_index = ignore_index
        self.max_samples = max_samples
        self.eps = float(eps)
        self.binning = binning
        self.plot_style = plot_style
        self._bins = None           # [3,n_bins] int64 on the GPU
        self._edges = ops.uniform_edges(self.n_bins)
        # binning="adaptive" (equal-mass edges, src/metrics/ece.py:118-126) needs the confidence distribution: a fine
        # [2, 60000] (correct | wrong) histogram of the confidences is kept beside the uniform bins and the quantile
        # edges are read off its CDF at compute() -- resolution 1/60000 in confidence instead of every sample
        self._fine = None

    # -- state ---------------------------------------------------------------------------------
    def _accumulator(self, dev=None) -> torch.Tensor:
        if self._bins is None:
            self._bins = ops.new_ece_bins(self.n_bins, _lib.require_cuda(dev))
        return self._bins

    def reset(self):
        if self._bins is not None:
            self._bins.zero_()
        if self._fine is not None:
            self._fine.zero_()

    @property
    def _seen(self) -> int:
        return 0 if self._bins is None else int(self._bins[0].sum().item())

    def state_dict(self):
        b = self._bins.cpu() if self._bins is not None else torch.zeros((3, self.n_bins), dtype=torch.int64)
        return {"ece_bins": b, "n_bins": self.n_bins, "mode": self.mode}

    def load_state_dict(self, sd):
        b = torch.as_tensor(sd["ece_bins"], dtype=torch.int64)
        if tuple(b.shape) != (3, self.n_bins):
            raise ValueError("ece_bins shape mismatch")
        self._accumulator().copy_(b)

    # -- accumulation --------------------------------------------------------------------------
    @torch.no_grad()
    def update(self, preds: torch.Tensor, labels: torch.Tensor):
        assert preds.dim() == 4 and labels.dim() == 3
        dev = preds.device if preds.is_cuda else (labels.device if labels.is_cuda else _lib.require_cuda())
        kind, conf_mode = _MODES[self.mode]
        labels = labels.to(dev, non_blocking=True)
        r = ops.reduce_metrics(preds.to(dev, non_blocking=True), labels, kind=kind,
                               conf_mode=conf_mode, eps=self.eps, ignore_index=self.ignore_index,
                               edges=self._edges, ece_bins=self._accumulator(dev),
                               want=("conf", "pred") if self.binning == "adaptive" else ())
        if self.binning == "adaptive":
            if self._fine is None:
                self._fine = ops.new_score_hist(dev)
            ops.score_hist(r["conf"], r["pred"], labels, self._fine,
                           ignore=() if self.ignore_index is None else (self.ignore_index,))

    def add_bins(self, ece_bins: torch.Tensor):
        """Merge counters produced elsewhere (the fused MC path, another rank)."""
        self._accumulator(ece_bins.device if ece_bins.is_cuda else None).add_(ece_bins.to(self._accumulator().device))

    # -- result --------------------------------------------------------------------------------
    def _adaptive_stats(self) -> pd.DataFrame:
        h = self._fine.cpu().numpy().astype(np.float64)
        M = h.shape[1]
        tot = h[0] + h[1]
        N = tot.sum()
        cdf = np.cumsum(tot) / N
        q = np.linspace(0.0, 1.0, self.n_bins + 1)
        edges = np.minimum((np.searchsorted(cdf, q, side="left") + 1) / M, 1.0)
        edges[0], edges[-1] = 0.0, 1.0
        edges = np.unique(edges)
        if edges.size < self.n_bins + 1:                      # many identical confidences: fall back, as ece.py:124-125
            edges = np.linspace(0.0, 1.0, self.n_bins + 1)
        centres = (np.arange(M) + 0.5) / M
        which = np.clip(np.searchsorted(edges, centres, side="right") - 1, 0, edges.size - 2)
        n = np.bincount(which, weights=tot, minlength=edges.size - 1)
        c = np.bincount(which, weights=h[0], minlength=edges.size - 1)
        sconf = np.bincount(which, weights=tot * centres, minlength=edges.size - 1)
        with np.errstate(invalid="ignore", divide="ignore"):
            acc = np.where(n > 0, c / n, np.nan)
            avg = np.where(n > 0, sconf / n, np.nan)
        lows, highs = edges[:-1].astype(np.float32), edges[1:].astype(np.float32)
        return pd.DataFrame({"low": lows, "high": highs, "center": 0.5 * (lows + highs), "width": highs - lows,
                             "n": n.astype(int), "pct": 100.0 * n / max(1, int(N)), "acc": acc, "conf": avg})

    def _stats_df(self) -> pd.DataFrame:
        if self._bins is None or self._seen == 0:
            return pd.DataFrame(columns=["low", "high", "center", "width", "n", "pct", "acc", "conf"])
        if self.binning == "adaptive" and self._fine is not None:
            return self._adaptive_stats()
        _, _, n, acc, avg = ops.ece_from_bins(self._bins)
        lows, highs = self._edges[:-1], self._edges[1:]
        return pd.DataFrame({"low": lows, "high": highs, "center": 0.5 * (lows + highs), "width": highs - lows,
                             "n": n.astype(int), "pct": 100.0 * n / max(1, int(n.sum())), "acc": acc, "conf": avg})

    def compute(self, save_plot_path: str | None = None, title: str = "Reliability Diagram", dpi: int = 200):
        stats = self._stats_df()
        if stats.empty or stats["n"].sum() == 0:
            return (float("nan"), float("nan")), stats          # 2-tuple, as ece.py:157-158
        w = stats["n"].to_numpy().astype(np.float64)
        acc = np.nan_to_num(stats["acc"].to_numpy(), nan=0.0)
        conf = np.nan_to_num(stats["conf"].to_numpy(), nan=0.0)
        gap = np.abs(acc - conf)
        ece = float(np.sum((w / max(1, w.sum())) * gap))
        mce = float(np.max(gap[w > 0]))
        fig = None
        if save_plot_path is not None:
            fig = self._plot(stats, acc, conf, ece, mce, save_plot_path, title, dpi)
        return (ece, mce), stats, fig

    def _plot(self, stats, acc, conf, ece, mce, path, title, dpi):
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
        except Exception:                       # plotting is optional; the numbers above are the result
            return None
        fig, ax = plt.subplots(figsize=(6.8, 5.0), dpi=dpi)
        x, widths = stats["center"].to_numpy(), stats["width"].to_numpy()
        if self.plot_style == "gap":
            signed = conf - acc
            ax.axhline(0.0, color="k", linewidth=1)
            ax.bar(x, signed, width=widths * 0.9, color=np.where(signed >= 0, "tab:red", "tab:green"))
            ax.set_ylabel("conf - acc  (positive = over-confident)")
        else:
            ax.plot([0, 1], [0, 1], label="perfect calibration", linewidth=2)
            ax.plot(x, acc, marker="o", label="accuracy")
            ax.plot(x, conf, marker="x", linestyle="--", label="avg. confidence")
            ax.set_ylabel("Accuracy / Avg. Confidence")
            ax.set_ylim(0, 1)
            if self.plot_style == "classic+hist":
                ax2 = ax.twinx()
                ax2.bar(x, stats["n"].to_numpy() / max(1, int(stats["n"].sum())), width=widths * 0.9, alpha=0.25)
                ax2.set_ylim(0, 1)
            ax.legend(loc="lower right")
        ax.set_xlim(0, 1)
        ax.set_xlabel("Confidence (bin center)")
        ax.set_title(f"{title}\nECE={ece:.4f}  |  MCE={mce:.4f}")
        ax.grid(True, alpha=0.3)
        fig.tight_layout()
        fig.savefig(path, bbox_inches="tight", dpi=dpi)
        return fig
