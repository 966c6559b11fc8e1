"""Risk-coverage metrics with the reference's interface (src/metrics/aurc.py), as a device histogram.

`UncertaintyAggregator(ignore_index, use_max_prob_confidence, reservoir_size, seed)`, `.add_batch(probs[B,C,H,W],
labels[B,1,H,W], ent_mc=None)`, `.finalize(...) -> {"AURC","EAURC","num_pixels",...}` (aurc.py:210-350),
`compute_batch_uncertainty_metrics` (:86-120), `entropy_from_probs` (:48-54), `aurc_from_risks_confids` (:38-45).

The reference keeps every valid pixel's (risk, confidence) on the host and argsorts them in `finalize`; here
`add_batch` is one pass of the fused uncertainty kernel (entropy / max-prob / arg-max maps) plus one histogram
kernel, and the whole state is `[2, 2^20]` int64 counts (uncertainty bin x is_error), so shards combine with one
all-reduce.  The curve is then built on the host from the 16 MB histogram with the reference's own point/weight
rule (a point at the first sample of each tie group), the samples of one bin of width 2^-20 counting as tied; the
reference orders exact ties arbitrarily, and the difference stays below 1e-5 absolute on continuous scores
(tests).  `reservoir_size` (a host-memory cap) is accepted and ignored: every pixel is counted.  E-AURC's optimal
term is evaluated in closed form from (n, n_errors) instead of sorting the risks.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib, ops

AURC_BINS = 1 << 20


def rc_from_hist(hist: np.ndarray):
    """(coverages, selective_risks, weights) from [2,M] counts over UNCERTAINTY bins (row 1 = errors), following
    aurc.py:7-35 with each non-empty bin as one tie group; within a group the risk of its first sample -- which the
    reference takes in unspecified (argsort) order -- is the group's error rate."""
    n_k = (hist[0] + hist[1])[::-1].astype(np.float64)          # ascending confidence = descending uncertainty
    e_k = hist[1][::-1].astype(np.float64)
    keep = n_k > 0
    n_k, e_k = n_k[keep], e_k[keep]
    n, total = float(n_k.sum()), float(e_k.sum())
    if n == 0:
        return np.zeros(0), np.zeros(0), np.zeros(0)
    start = np.concatenate(([0.0], np.cumsum(n_k)[:-1]))        # index of each group's first sample
    e_before = np.concatenate(([0.0], np.cumsum(e_k)[:-1]))
    rec = start <= n - 2                                        # the walk stops at i = n - 2
    start, e_before, first = start[rec], e_before[rec], (e_k / n_k)[rec]
    coverages = np.concatenate(([1.0], (n - 1.0 - start) / n))
    sel = np.concatenate(([total / n], (total - e_before - first) / (n - 1.0 - start)))
    weights = np.diff(np.concatenate(([-1.0], start))) / n
    tail = (n - 2.0 - start[-1]) if start.size else 0.0
    if tail > 0:
        coverages, sel, weights = np.append(coverages, 0.0), np.append(sel, sel[-1]), np.append(weights, tail / n)
    return coverages, sel, weights


def optimal_aurc(n: int, n_err: int) -> float:
    """mean_m max(0, m - (n - e)) / m over m = 1..n: the reference's cumsum(sort(risks))/arange (aurc.py:42-43) for 0/1 risks."""
    if n <= 0 or n_err <= 0:
        return 0.0
    from scipy.special import digamma
    k = n - n_err
    return float((n_err - k * (digamma(n + 1.0) - digamma(k + 1.0))) / n)


def aurc_from_hist(hist: np.ndarray):
    """-> (aurc, eaurc, coverages, rc_risks)"""
    cov, sel, w = rc_from_hist(hist)
    if sel.size == 0:
        return float("nan"), float("nan"), cov, sel
    aurc = float(np.sum((sel[:-1] + sel[1:]) * 0.5 * w))
    return aurc, aurc - optimal_aurc(int(hist.sum()), int(hist[1].sum())), cov, sel


def error_recall_from_hist(hist: np.ndarray, ks=(1, 2, 5, 10, 20, 30, 40, 50)) -> np.ndarray:
    """Recall of errors among the k% most uncertain pixels (aurc.py:101-108); a bin cut by the k% boundary
    contributes its errors pro rata."""
    n_k = (hist[0] + hist[1])[::-1].astype(np.float64)
    e_k = hist[1][::-1].astype(np.float64)
    n, total = n_k.sum(), e_k.sum()
    cn, ce = np.concatenate(([0.0], np.cumsum(n_k))), np.concatenate(([0.0], np.cumsum(e_k)))
    out = []
    for k in ks:
        m = max(1, int(n * k / 100))
        out.append(float(np.interp(m, cn, ce) / max(total, 1.0)))
    return np.asarray(out)


@torch.no_grad()
def entropy_from_probs(probs: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """Normalised entropy [B,H,W] of probs [B,C,H,W] (aurc.py:48-54), by the fused reduction kernel."""
    return ops.reduce_metrics(probs, kind="probs", conf_mode=ops.CONF_RAW, eps=eps, want=("H_norm",), normalize=True)["H_norm"]


@torch.no_grad()
def _accumulate(hist, outputs_semantic, semantic, ent_mc, ignore_index, use_max_prob_confidence):
    dev = outputs_semantic.device if outputs_semantic.is_cuda else _lib.require_cuda()
    probs = outputs_semantic.to(dev, non_blocking=True)
    gt = semantic.to(dev, non_blocking=True)
    if gt.dim() == 4:
        gt = gt[:, 0]
    r = ops.reduce_metrics(probs, kind="probs", conf_mode=ops.CONF_RAW, eps=1e-12, want=("H_norm", "conf", "pred"), normalize=True)
    if use_max_prob_confidence:
        score = 1.0 - r["conf"]                       # confidence = max prob  <=>  uncertainty = 1 - max prob
    else:
        score = r["H_norm"] if ent_mc is None else ent_mc.to(dev)     # confidence = 1 - clamp(entropy, 0, 1); the kernel clamps
    ops.score_hist(score, r["pred"], gt, hist, ignore=(int(ignore_index),))


def compute_batch_uncertainty_metrics(outputs_semantic, semantic, ent_mc=None, ignore_index: int = 255,
                                      use_max_prob_confidence: bool = False, ks=(1, 2, 5, 10, 20, 30, 40, 50)):
    """AURC / E-AURC / curves for one batch (aurc.py:86-120)."""
    dev = outputs_semantic.device if outputs_semantic.is_cuda else _lib.require_cuda()
    hist = ops.new_score_hist(dev, AURC_BINS)
    _accumulate(hist, outputs_semantic, semantic, ent_mc, ignore_index, use_max_prob_confidence)
    h = hist.cpu().numpy()
    aurc, eaurc, cov, sel = aurc_from_hist(h)
    return {"AURC": aurc, "EAURC": eaurc, "coverages": cov, "rc_risks": sel, "ks": np.asarray(ks),
            "recalls": error_recall_from_hist(h, ks), "num_pixels": int(h.sum()), "num_errors": int(h[1].sum()),
            "valid_mask_shape": tuple(semantic.squeeze(1).shape)}


def aurc_from_risks_confids(risks, confids):
    """(aurc, eaurc, coverages, rc_risks) for flat 0/1 risks and confidences in [0,1] (aurc.py:38-45): the arrays are
    histogrammed on the device (uncertainty = 1 - confidence)."""
    dev = _lib.require_cuda()
    risks = torch.as_tensor(np.asarray(risks)).to(dev)
    conf = torch.as_tensor(np.asarray(confids), dtype=torch.float32).to(dev)
    hist = ops.new_score_hist(dev, AURC_BINS)
    # pred != label  <=>  risk == 1
    ops.score_hist(1.0 - conf, (risks > 0.5).long(), torch.zeros_like(risks, dtype=torch.long), hist)
    return aurc_from_hist(hist.cpu().numpy())


class UncertaintyAggregator:
    def __init__(self, ignore_index: int = 255, use_max_prob_confidence: bool = False, reservoir_size=None, seed=None):
        self.ignore_index = int(ignore_index)
        self.use_max_prob_confidence = bool(use_max_prob_confidence)
        self.reservoir_size = None if reservoir_size is None else int(reservoir_size)
        self._hist = None

    def reset(self):
        if self._hist is not None:
            self._hist.zero_()

    @torch.no_grad()
    def add_batch(self, outputs_semantic: torch.Tensor, semantic: torch.Tensor, ent_mc: torch.Tensor | None = None):
        assert outputs_semantic.ndim == 4 and semantic.ndim == 4 and semantic.shape[1] == 1
        if self._hist is None:
            dev = outputs_semantic.device if outputs_semantic.is_cuda else _lib.require_cuda()
            self._hist = ops.new_score_hist(dev, AURC_BINS)
        _accumulate(self._hist, outputs_semantic, semantic, ent_mc, self.ignore_index, self.use_max_prob_confidence)

    def add_maps(self, uncertainty: torch.Tensor, pred: torch.Tensor, labels: torch.Tensor):
        """Accumulate from maps the fused kernel already produced (tester path: no second pass over the classes)."""
        if self._hist is None:
            self._hist = ops.new_score_hist(uncertainty.device, AURC_BINS)
        ops.score_hist(uncertainty, pred, labels, self._hist, ignore=(self.ignore_index,))

    def state(self) -> torch.Tensor:
        """[2, 2^20] int64 counts; sum over ranks (dist.allreduce_counts) before finalize() on a sharded sweep."""
        return self._hist

    def finalize(self, make_plots: bool = True, title_suffix: str = "", save_dir=None, epoch=None,
                 filename_prefix: str = "", dpi: int = 150, show: bool = False, close: bool = True):
        if self._hist is None or int(self._hist.sum().item()) == 0:
            raise RuntimeError("No batches added. Call add_batch(...) first.")
        h = self._hist.cpu().numpy()
        aurc, eaurc, cov, sel = aurc_from_hist(h)
        rc_path = None
        if make_plots and save_dir:
            try:
                import os
                import matplotlib
                matplotlib.use("Agg")
                import matplotlib.pyplot as plt
                os.makedirs(save_dir, exist_ok=True)
                tag = f"{filename_prefix}epoch_{int(epoch):06d}_" if epoch is not None else filename_prefix
                rc_path = os.path.join(save_dir, f"{tag}rc_curve.png")
                fig = plt.figure()
                plt.plot(cov, sel); plt.xlabel("Coverage (fraction kept)"); plt.ylabel("Risk (error rate on kept)")
                plt.title(f"Risk-Coverage (dataset){title_suffix}"); plt.grid(True, linestyle=":")
                fig.tight_layout(); fig.savefig(rc_path, bbox_inches="tight", dpi=dpi); plt.close(fig)
            except Exception:
                rc_path = None
        return {"AURC": aurc, "EAURC": eaurc, "num_pixels": int(h.sum()), "rc_curve_path": rc_path, "error_recall_path": None}
