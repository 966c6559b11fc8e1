"""Dirichlet uncertainty measures with the reference's names (src/models/probability_helper.py:13-247),
computed by the fused evidential kernel (csrc/slu_evidential.cu).

Only the hot-path functions are here (alpha construction, predictive entropy, aleatoric / epistemic
uncertainty and their normalised variants); the reference's visualisation and label-smoothing helpers
(:41-87, :251-450) are control-plane code and stay the reference's.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import _lib, ops
from ..utils.agg import mean_aggregator

_NORM_MODE = "max"
_EPS: float = 1e-8
_T: float = 1.0


def set_norm_mode(mode: str):
    global _NORM_MODE
    if mode not in ("max", "ref"):
        raise ValueError(f"norm_mode must be 'max' or 'ref', got: {mode}")
    _NORM_MODE = mode


def get_norm_mode() -> str:
    return _NORM_MODE


def set_eps_value(eps: float):
    global _EPS
    _EPS = eps


def get_eps_value() -> float:
    return _EPS


def set_alpha_temperature(T: float):
    global _T
    _T = T


def get_alpha_temperature() -> float:
    return _T


def _cuda(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_cuda else t.to(_lib.require_cuda(), non_blocking=True)


def to_alpha_concentrations_from_shape_and_scale(shape_logits, scale_logits, T=None, eps=None):
    """alpha = 1 + softplus(scale/T) * softmax(shape) + eps  (:89-105).  No autograd: evaluation path;
    the training path builds alpha with torch ops so the graph reaches the backbone."""
    T = get_alpha_temperature() if T is None else T
    eps = get_eps_value() if eps is None else eps
    if shape_logits.requires_grad or scale_logits.requires_grad:
        scale = torch.nn.functional.softplus(scale_logits / T)
        return 1.0 + scale * torch.nn.functional.softmax(shape_logits, dim=1) + eps
    x = torch.cat([_cuda(shape_logits), _cuda(scale_logits)], dim=1)
    return ops.evidential_reduce(x, from_outputs=True, temperature=T, eps=eps, want=("alpha",))["alpha"]


def _measure(alpha, eps, key, normalize=False):
    eps = get_eps_value() if eps is None else eps
    return ops.evidential_reduce(_cuda(alpha.detach()), from_outputs=False, eps=eps, normalize=normalize, want=(key,))[key]


def get_predictive_entropy(alpha, eps=None):
    return _measure(alpha, eps, "H")                      # :116-121


def get_aleatoric_uncertainty(alpha, eps=None):
    return _measure(alpha, eps, "AU")                     # :124-130


def get_epistemic_uncertainty(alpha, eps=None):
    return _measure(alpha, eps, "EU")                     # :133-136


@mean_aggregator()
def get_predictive_entropy_norm(alpha, eps=None):
    return _measure(alpha, eps, "H", normalize=True)      # :148-153


def _au_ref(C: int) -> float:                             # psi(C+1) - psi(2) = H_C - 1  (:139-140)
    return sum(1.0 / k for k in range(2, C + 1))


def get_aleatoric_uncertainty_norm(alpha, eps=None, mode=None):
    """:156-187.  The per-pixel maps are [B,H,W]; the remaps are a handful of scalar ops on them."""
    eps = get_eps_value() if eps is None else eps
    C = alpha.shape[1]
    AU = get_aleatoric_uncertainty(alpha, eps)
    m = get_norm_mode() if mode is None else mode
    if m == "max":
        return (AU / math.log(C)).clamp(0.0, 1.0)
    if m == "ref":
        au_ref = _au_ref(C)
        span = max(math.log(C) - au_ref, eps)
        raw = (AU - au_ref) / span
        L = -au_ref / span
        return ((raw - L) / (1.0 - L)).clamp(0.0, 1.0)
    raise ValueError(f"Unknown mode: {m}")


def get_epistemic_uncertainty_norm(alpha, eps=None, mode=None):
    """:190-214."""
    eps = get_eps_value() if eps is None else eps
    m = get_norm_mode() if mode is None else mode
    if m == "max":
        return (get_epistemic_uncertainty(alpha, eps) / math.log(alpha.shape[1])).clamp(0.0, 1.0)
    if m == "ref":
        return (1.0 - get_aleatoric_uncertainty_norm(alpha, eps=eps, mode="ref")).clamp(0.0, 1.0)
    raise ValueError(f"Unknown mode: {m}")


def _fractions(alpha, eps, min_h):
    eps = get_eps_value() if eps is None else eps
    min_h = get_eps_value() if min_h is None else min_h
    r = ops.evidential_reduce(_cuda(alpha.detach()), from_outputs=False, eps=eps, normalize=False, want=("H", "AU", "EU"))
    h = torch.clamp(r["H"], min=min_h)
    return (r["AU"] / h).clamp(0.0, 1.0), (r["EU"] / h).clamp(0.0, 1.0)


def get_aleatoric_fraction(alpha, eps=None, min_h=None):
    return _fractions(alpha, eps, min_h)[0]               # :218-225


def get_epistemic_fraction(alpha, eps=None, min_h=None):
    return _fractions(alpha, eps, min_h)[1]               # :228-235


def get_eu_minus_au_fraction(alpha, eps=None, min_h=None):
    auf, euf = _fractions(alpha, eps, min_h)
    return (euf - auf).clamp(-1.0, 1.0)                   # :238-246


def _colour(x: np.ndarray, lo: float, hi: float, mask):
    """float map -> BGR uint8 image with OpenCV's TURBO colour map; `mask` rows (y, x) are blacked out."""
    import cv2
    x = np.clip((x - lo) / (hi - lo + 1e-12), 0, 1)
    img = cv2.applyColorMap((x * 255).astype(np.uint8), cv2.COLORMAP_TURBO)
    if mask is not None:
        img[mask[:, 0], mask[:, 1]] = [0, 0, 0]
    return img


def build_uncertainty_layers(alpha: torch.Tensor, names: list, idx: int = 0, h_mask_thresh: float = 0.0, eps=None, mask=None) -> dict:
    """Visualisation layers of ONE sample (:294-335): every requested map comes out of a single evidential kernel call on
    alpha[idx]; quantile clipping (2 % / 98 %) and the TURBO colour map run on the host image, as in the reference."""
    eps = get_eps_value() if eps is None else eps
    a = _cuda(alpha.detach())[idx:idx + 1]
    C = a.shape[1]
    r = ops.evidential_reduce(a, from_outputs=False, eps=eps, normalize=False, want=("H", "AU", "EU"))
    H, AU, EU = r["H"][0], r["AU"][0], r["EU"][0]
    denom = torch.clamp(H, min=1e-6)
    maps = {}
    if "H_norm" in names:
        maps["H_norm"] = H / math.log(C)
    if "AU_norm" in names:
        maps["AU_norm"] = get_aleatoric_uncertainty_norm(a, eps)[0]
    if "EU_norm" in names:
        maps["EU_norm"] = get_epistemic_uncertainty_norm(a, eps)[0]
    if "alpha0" in names:
        maps["alpha0"] = a.sum(dim=1)[0] + eps
    if "AU_frac" in names:
        maps["AU_frac"] = (AU / denom).clamp(0.0, 1.0)
    if "EU_frac" in names:
        maps["EU_frac"] = (EU / denom).clamp(0.0, 1.0)
    out = {}
    for k, m in maps.items():
        x = m.float().cpu().numpy()
        lo, hi = np.quantile(x, 0.02), np.quantile(x, 0.98)
        if hi <= lo:
            lo, hi = x.min(), x.max() + 1e-6
        out[k] = _colour(x, lo, hi, mask)
    if "EU_minus_AU_frac" in names:
        d = ((EU / denom).clamp(0.0, 1.0) - (AU / denom).clamp(0.0, 1.0)).clamp(-1.0, 1.0)
        out["EU_minus_AU_frac"] = _colour(d.float().cpu().numpy(), -1.0, 1.0, mask)
    return out


@torch.no_grad()
def evidential_reduce_from_outputs(outputs, labels=None, *, num_classes=None, iou_evaluator=None, ece_eval=None,
                                   auroc_eval=None, auroc_eval_mi=None, ua_agg=None, unc_agg=None, ua_ignore_ids=(0,),
                                   want=("pred", "conf", "H", "AU", "EU", "MI")):
    """The whole single-pass Dirichlet branch of Tester.test_epoch (src/models/tester.py:484-516) in one
    kernel: head output [B,C+1,H,W] -> pred, H_norm, AU, EU, MI_norm (+ alpha if asked) and the
    IoUEvaluator / ECEAggregator(mode='alpha') updates; the AUROC (entropy and Dirichlet-MI scores, :510-512),
    accuracy-vs-uncertainty (:502-508) and per-class uncertainty (:516) aggregators are then fed from the small
    per-pixel maps, so alpha [B,C,H,W] is never materialised."""
    x = _cuda(outputs)
    if num_classes is not None and num_classes + 1 != x.shape[1]:
        x = x[:, : num_classes + 1].contiguous()
    confmat = iou_evaluator._accumulator(x.device) if iou_evaluator is not None else None
    bins = ece_eval._accumulator(x.device) if ece_eval is not None else None
    need = set(want)
    if auroc_eval is not None or ua_agg is not None or unc_agg is not None:
        need |= {"pred", "H"}
    if auroc_eval_mi is not None:
        need |= {"pred", "MI"}
    lab = None if labels is None else _cuda(labels)
    out = ops.evidential_reduce(x, lab, from_outputs=True,
                                temperature=get_alpha_temperature(), eps=get_eps_value(),
                                eps_metrics=ece_eval.eps if ece_eval is not None else 1e-12,
                                ignore_index=ece_eval.ignore_index if ece_eval is not None else None,
                                edges=ece_eval._edges if ece_eval is not None else None,
                                confmat=confmat, ece_bins=bins, want=tuple(need))
    if lab is not None:
        lab3 = lab[:, 0] if lab.dim() == 4 else lab
        # note: the reference's AUROC 'entropy_norm' score in alpha mode clamps p at 1e-12 (auroc.py:48-53) where
        # probability_helper adds 1e-8 inside the log; both give the same ranking up to ~1e-7 in the score
        if auroc_eval is not None:
            auroc_eval.add_maps(out["H"], out["pred"], lab3)
        if auroc_eval_mi is not None:
            auroc_eval_mi.add_maps(out["MI"], out["pred"], lab3)
        if ua_agg is not None:
            ua_agg.update(labels=lab3, preds=out["pred"], uncertainty=out["H"], ignore_ids=ua_ignore_ids)
        if unc_agg is not None:
            unc_agg.update(lab3, out["H"])
    return out
