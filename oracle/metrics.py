"""CPU restatement of the reference's evaluation histograms.  TEST INFRASTRUCTURE ONLY.

Confusion matrix / mIoU : src/models/evaluator.py:29-105 (`IoUEvaluator`).
ECE / reliability bins   : src/metrics/ece.py:54-168 (`ECEAggregator`, max_samples=None).
Pinned against the unmodified reference in tests/golden/ (oracle/gen_golden.py).
"""
from __future__ import annotations

import numpy as np
import torch


def confusion_counts(preds: torch.Tensor, targets: torch.Tensor, num_classes: int) -> torch.Tensor:
    """evaluator.py:39-53: rows = GT, cols = prediction, out-of-range pairs dropped."""
    preds = preds.reshape(-1)
    targets = targets.reshape(-1)
    C = num_classes
    ok = (targets >= 0) & (targets < C) & (preds >= 0) & (preds < C)
    idx = targets[ok] * C + preds[ok]
    return torch.bincount(idx, minlength=C * C).reshape(C, C)


def iou_from_confmat(confmat: torch.Tensor, test_mask=None, ignore_gt=None, reduce="mean", ignore_th=None):
    """evaluator.py:55-105 -> (mIoU, per-class IoU float64 tensor with NaN where undefined)."""
    cm = confmat.clone().double()
    C = cm.shape[0]
    if ignore_gt:
        rows = torch.tensor(ignore_gt, dtype=torch.long)
        rows = rows[(rows >= 0) & (rows < C)]
        cm[rows, :] = 0.0
    TP = cm.diag()
    FP = cm.sum(0) - TP
    FN = cm.sum(1) - TP
    denom = TP + FP + FN
    iou = torch.full((C,), float("nan"), dtype=torch.float64)
    valid = denom > 0
    iou[valid] = TP[valid] / denom[valid]
    if test_mask is None:
        test_mask = torch.ones(C, dtype=torch.bool)
    else:
        test_mask = torch.as_tensor(test_mask, dtype=torch.bool)
    avg_mask = test_mask & torch.isfinite(iou)
    if ignore_th is not None:
        avg_mask = avg_mask & (iou >= ignore_th)
    if avg_mask.any():
        vals = iou[avg_mask].numpy()
        miou = float(np.mean(vals)) if reduce == "mean" else float(np.median(vals))
    else:
        miou = float("nan")
    return miou, iou


def to_probs(preds: torch.Tensor, mode: str, eps: float = 1e-12) -> torch.Tensor:
    """ece.py:54-64."""
    if mode == "alpha":
        a0 = preds.sum(dim=1, keepdim=True)
        return preds / (a0 + eps)
    if mode == "logits":
        return preds.softmax(dim=1)
    p = preds.clamp_min(0)
    return p / p.sum(dim=1, keepdim=True).clamp_min(eps)


def ece_samples(preds: torch.Tensor, labels: torch.Tensor, mode: str, ignore_index=None, eps: float = 1e-12):
    """ece.py:66-90: the (conf float32, correct bool) samples one update() appends."""
    p = to_probs(preds, mode, eps)
    conf, pred = p.max(dim=1)
    lab = labels.long()
    valid = (lab != ignore_index) if ignore_index is not None else torch.ones_like(lab, dtype=torch.bool)
    conf = conf[valid].to(torch.float32).view(-1).clamp(0, 1)
    corr = pred[valid].view(-1) == lab[valid].view(-1)
    return conf, corr


def ece_edges(n_bins: int) -> np.ndarray:
    """ece.py:116,128: float32 uniform edges with exact 0 and 1 at the ends."""
    edges = np.linspace(0.0, 1.0, n_bins + 1, dtype=np.float32)
    edges[0] = 0.0
    edges[-1] = 1.0
    return edges


def ece_bin_index(conf: np.ndarray, edges: np.ndarray) -> np.ndarray:
    """np.histogram's membership rule: [lo, hi) with a closed last bin; -1 outside."""
    idx = np.searchsorted(edges, conf, side="right") - 1
    idx = np.where(conf == edges[-1], len(edges) - 2, idx)
    idx = np.where((conf < edges[0]) | (conf > edges[-1]), -1, idx)
    return idx


def ece_bin_counts(conf: np.ndarray, corr: np.ndarray, n_bins: int):
    """Exact per-bin (n, n_correct, sum_conf float64) -- what a streaming histogram accumulates."""
    edges = ece_edges(n_bins)
    idx = ece_bin_index(conf, edges)
    ok = idx >= 0
    n = np.bincount(idx[ok], minlength=n_bins).astype(np.int64)
    nc = np.bincount(idx[ok], weights=corr[ok].astype(np.float64), minlength=n_bins).astype(np.int64)
    cs = np.bincount(idx[ok], weights=conf[ok].astype(np.float64), minlength=n_bins)
    return n, nc, cs


def ece_reference_stats(conf: np.ndarray, corr: np.ndarray, n_bins: int):
    """ece.py:131-148: the three np.histogram calls, float32 weights exactly as the reference."""
    edges = ece_edges(n_bins)
    corr_f = corr.astype(np.float32)
    n = np.histogram(conf, bins=edges)[0].astype(int)
    acc_s = np.histogram(conf, bins=edges, weights=corr_f)[0]
    conf_s = np.histogram(conf, bins=edges, weights=conf)[0]
    acc = np.divide(acc_s, n, out=np.full_like(acc_s, np.nan, dtype=float), where=n > 0)
    avg_conf = np.divide(conf_s, n, out=np.full_like(conf_s, np.nan, dtype=float), where=n > 0)
    return n, acc, avg_conf


def ece_from_stats(n, acc, avg_conf):
    """ece.py:160-168 -> (ece, mce)."""
    w = np.asarray(n).astype(np.float64)
    if w.sum() == 0:
        return float("nan"), float("nan")
    acc = np.nan_to_num(np.asarray(acc, dtype=np.float64), nan=0.0)
    conf = np.nan_to_num(np.asarray(avg_conf, dtype=np.float64), nan=0.0)
    gap = np.abs(acc - conf)
    ece = float(np.sum((w / max(1, w.sum())) * gap))
    nonempty = w > 0
    mce = float(np.max(gap[nonempty])) if np.any(nonempty) else float("nan")
    return ece, mce


def ece_from_counts(n, n_correct, conf_sum):
    """ECE from streaming-histogram state (exact integer counts + float64 confidence sums)."""
    n = np.asarray(n, dtype=np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        acc = np.where(n > 0, np.asarray(n_correct, dtype=np.float64) / n, np.nan)
        avg = np.where(n > 0, np.asarray(conf_sum, dtype=np.float64) / n, np.nan)
    return ece_from_stats(n, acc, avg)


def auroc_error_detection(scores: np.ndarray, is_error: np.ndarray) -> float:
    """src/metrics/auroc.py:65-78 (`_roc_from_scores`): sort by descending score, trapezoid over (fpr, tpr)."""
    order = np.argsort(-scores)
    y = is_error[order].astype(np.float64)
    P, N = y.sum(), y.size - y.sum()
    if P == 0 or N == 0:
        return float("nan")
    tpr = np.concatenate(([0.0], np.cumsum(y) / P, [1.0]))
    fpr = np.concatenate(([0.0], np.cumsum(1.0 - y) / N, [1.0]))
    return float(np.trapezoid(tpr, fpr))


def binned_accuracy(uncert: np.ndarray, correct: np.ndarray, num_bins: int):
    """src/models/evaluator.py:726-749: np.histogram over float32 uniform edges -> (n, accuracy)."""
    edges = np.linspace(0.0, 1.0, num_bins + 1, dtype=np.float32)
    edges[0], edges[-1] = 0.0, 1.0
    n = np.histogram(uncert, bins=edges)[0].astype(int)
    csum = np.histogram(uncert, bins=edges, weights=correct.astype(np.float32))[0]
    acc = np.divide(csum, n, out=np.full_like(csum, np.nan, dtype=float), where=n > 0)
    return n, acc


# ---- risk-coverage / AURC (src/metrics/aurc.py:7-45), exact restatement without the Python loop -------------------
def rc_curve_stats(risks: np.ndarray, confids: np.ndarray):
    """aurc.py:7-35.  The reference walks the samples in ascending confidence, drops one at a time and records a
    point at i == 0 and at the first sample of every new tie group; the trailing samples form one last segment."""
    n = risks.size
    idx = np.argsort(confids)                       # same call (same tie order) as aurc.py:10
    r = risks[idx].astype(np.float64)
    c = confids[idx]
    total = float(r.sum())
    i = np.arange(0, max(n - 1, 0))
    rec = i[(i == 0) | (c[i] != c[np.maximum(i - 1, 0)])] if n > 1 else np.zeros(0, dtype=np.int64)
    cs = np.cumsum(r)
    coverages = np.concatenate(([1.0], (n - 1 - rec) / n))
    sel = np.concatenate(([total / n], (total - cs[rec]) / (n - 1 - rec)))
    weights = np.diff(np.concatenate(([-1], rec))) / n
    tail = (n - 2 - rec[-1]) if rec.size else 0
    if tail > 0:
        coverages = np.append(coverages, 0.0)
        sel = np.append(sel, sel[-1])
        weights = np.append(weights, tail / n)
    return coverages, sel, weights


def aurc_from_risks_confids(risks: np.ndarray, confids: np.ndarray):
    """aurc.py:38-45 -> (aurc, eaurc, coverages, rc_risks)."""
    coverages, rc_risks, weights = rc_curve_stats(risks, confids)
    aurc = float(np.sum((rc_risks[:-1] + rc_risks[1:]) * 0.5 * weights))
    n = risks.size
    opt = np.cumsum(np.sort(risks)) / np.arange(1, n + 1)
    return aurc, aurc - float(opt.sum() / n), coverages, rc_risks
