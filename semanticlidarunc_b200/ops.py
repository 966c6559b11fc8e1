"""Functional wrappers over the C ABI: torch CUDA tensors in, torch CUDA tensors out.

These are the device-resident entry points (no host copies, no synchronisation); the modules that
mirror the reference's Python interfaces (dataset/, metrics/, models/, utils/) are built on them.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import CONF_RAW, CONF_RENORM, IN_ALPHA, IN_LOGITS, IN_PROBS  # noqa: F401  (re-exported)

KINDS = {"logits": IN_LOGITS, "probs": IN_PROBS, "alpha": IN_ALPHA}


def uniform_edges(n_bins: int) -> np.ndarray:
    """float32 edges of the reference's uniform binning (src/metrics/ece.py:116,128)."""
    e = np.linspace(0.0, 1.0, n_bins + 1, dtype=np.float32)
    e[0], e[-1] = 0.0, 1.0
    return e


def new_confmat(num_classes: int, device) -> torch.Tensor:
    return torch.zeros((num_classes, num_classes), dtype=torch.int64, device=device)


def new_ece_bins(n_bins: int, device) -> torch.Tensor:
    """[3, n_bins] int64: n | n_correct | sum(conf) in units of 2^-32."""
    return torch.zeros((3, n_bins), dtype=torch.int64, device=device)


@_lib.device_guard
def reduce_metrics(x: torch.Tensor, labels: Optional[torch.Tensor] = None, *, kind: str = "logits",
                   conf_mode: int = CONF_RAW, eps: float = 1e-12, ignore_index: Optional[int] = None,
                   edges: Optional[Sequence[float]] = None,
                   confmat: Optional[torch.Tensor] = None, ece_bins: Optional[torch.Tensor] = None,
                   want: Sequence[str] = ("pred", "conf", "H_norm", "MI_norm"), normalize: bool = True,
                   direct: bool = False) -> dict:
    """Fused stage 3+4 (slu_reduce_metrics).

    x: [T,B,C,H,W] or [B,C,H,W] (T=1) float32 CUDA.  labels: [B,H,W] int64 CUDA or None.
    `want` selects the per-pixel outputs to materialise ("p_bar", "pred", "conf", "H_norm", "MI_norm").
    confmat [C,C] / ece_bins [3,n_bins] int64 are accumulated in place when given.
    """
    _lib.require_cuda()
    if x.dim() == 4:
        x = x.unsqueeze(0)
    if x.dim() != 5:
        raise ValueError("x must be [T,B,C,H,W] or [B,C,H,W]")
    x = _lib.as_buffer(x, torch.float32, "x")
    T, B, Cc, H, W = x.shape
    HW = H * W
    dev = x.device
    if labels is not None:
        if labels.dim() == 4 and labels.size(1) == 1:
            labels = labels[:, 0]
        if tuple(labels.shape) != (B, H, W):
            raise ValueError(f"labels shape {tuple(labels.shape)} != {(B, H, W)}")
        labels = _lib.as_buffer(labels, torch.int64, "labels")
    out = {}
    def mk(name, shape, dtype):
        if name in want:
            out[name] = torch.empty(shape, dtype=dtype, device=dev)
            return out[name]
        return None
    pbar = mk("p_bar", (B, Cc, H, W), torch.float32)
    pred = mk("pred", (B, H, W), torch.int64)
    conf = mk("conf", (B, H, W), torch.float32)
    hn = mk("H_norm", (B, H, W), torch.float32)
    mi = mk("MI_norm", (B, H, W), torch.float32)
    n_bins, h_edges = 0, None
    if ece_bins is not None:
        if ece_bins.dtype != torch.int64 or not ece_bins.is_contiguous() or ece_bins.dim() != 2 or ece_bins.size(0) != 3:
            raise ValueError("ece_bins must be a contiguous [3,n_bins] int64 tensor")
        n_bins = ece_bins.size(1)
        e = uniform_edges(n_bins) if edges is None else np.asarray(edges, dtype=np.float32)
        if e.shape[0] != n_bins + 1:
            raise ValueError("edges must have n_bins+1 entries")
        h_edges = _lib.edges_array(e)
    if confmat is not None and (confmat.dtype != torch.int64 or not confmat.is_contiguous() or tuple(confmat.shape) != (Cc, Cc)):
        raise ValueError(f"confmat must be a contiguous [{Cc},{Cc}] int64 tensor")
    fn = _lib.lib().slu_reduce_metrics_direct if direct else _lib.lib().slu_reduce_metrics
    rc = fn(_lib.ptr(x), _lib.ptr(labels), T, B, Cc, HW, KINDS[kind], conf_mode, float(eps), int(normalize),
            0 if ignore_index is None else 1, 0 if ignore_index is None else int(ignore_index),
            n_bins, h_edges,
            _lib.ptr(pbar), _lib.ptr(pred), _lib.ptr(conf), _lib.ptr(hn), _lib.ptr(mi),
            _lib.ptr(confmat), _lib.ptr(ece_bins), _lib.stream_ptr())
    _lib.check(rc, "slu_reduce_metrics")
    return out


@_lib.device_guard
def evidential_reduce(x: torch.Tensor, labels: Optional[torch.Tensor] = None, *, from_outputs: bool,
                      temperature: float = 1.0, eps: float = 1e-8, eps_metrics: float = 1e-12, normalize: bool = True,
                      ignore_index: Optional[int] = None, edges=None,
                      confmat: Optional[torch.Tensor] = None, ece_bins: Optional[torch.Tensor] = None,
                      want: Sequence[str] = ("pred", "conf", "H", "AU", "EU", "MI")) -> dict:
    """Evidential stage 3+4 (slu_evidential_reduce).

    x: head output [B,C+1,H,W] (from_outputs=True: C shape logits + 1 scale logit) or alpha [B,C,H,W].
    `want` may contain "alpha", "pred", "conf", "H", "AU", "EU", "MI".
    """
    _lib.require_cuda()
    x = _lib.as_buffer(x, torch.float32, "x")
    if x.dim() != 4:
        raise ValueError("x must be [B,C(+1),H,W]")
    B, Cx, H, W = x.shape
    Cc = Cx - 1 if from_outputs else Cx
    dev = x.device
    if labels is not None:
        if labels.dim() == 4 and labels.size(1) == 1:
            labels = labels[:, 0]
        if tuple(labels.shape) != (B, H, W):
            raise ValueError(f"labels shape {tuple(labels.shape)} != {(B, H, W)}")
        labels = _lib.as_buffer(labels, torch.int64, "labels")
    out = {}
    def mk(name, shape, dtype):
        if name in want:
            out[name] = torch.empty(shape, dtype=dtype, device=dev)
            return out[name]
        return None
    alpha = mk("alpha", (B, Cc, H, W), torch.float32)
    pred = mk("pred", (B, H, W), torch.int64)
    conf, h, au, eu, mi = (mk(k, (B, H, W), torch.float32) for k in ("conf", "H", "AU", "EU", "MI"))
    n_bins, h_edges = 0, None
    if ece_bins is not None:
        n_bins = ece_bins.size(1)
        h_edges = _lib.edges_array(uniform_edges(n_bins) if edges is None else np.asarray(edges, dtype=np.float32))
    rc = _lib.lib().slu_evidential_reduce(_lib.ptr(x) if from_outputs else None, None if from_outputs else _lib.ptr(x),
                                          _lib.ptr(labels), B, Cc, H * W, float(temperature), float(eps), float(eps_metrics),
                                          int(normalize), 0 if ignore_index is None else 1,
                                          0 if ignore_index is None else int(ignore_index), n_bins, h_edges,
                                          _lib.ptr(alpha), _lib.ptr(pred), _lib.ptr(conf), _lib.ptr(h), _lib.ptr(au),
                                          _lib.ptr(eu), _lib.ptr(mi), _lib.ptr(confmat), _lib.ptr(ece_bins), _lib.stream_ptr())
    _lib.check(rc, "slu_evidential_reduce")
    return out


@_lib.device_guard
def dirichlet_loss(alpha: torch.Tensor, target: torch.Tensor, *, ignore=(), keep_mask: Optional[torch.Tensor] = None,
                   eps_mse: float = 1e-8, eps_kl: float = 1e-8, want_mse: bool = True, want_kl: bool = True,
                   want_grad: bool = True, sums: Optional[torch.Tensor] = None) -> dict:
    """Evidential loss terms forward + analytic backward (slu_dirichlet_loss).

    alpha [B,C,H,W] float32 CUDA, target [B,H,W] int64.  Returns sums float64[3] (sum mse | sum kl |
    n_valid; accumulated into `sums` when given) and the per-pixel gradients grad_mse / grad_kl [B,C,H,W].
    """
    _lib.require_cuda()
    alpha = _lib.as_buffer(alpha, torch.float32, "alpha")
    if alpha.dim() != 4:
        raise ValueError("alpha must be [B,C,H,W]")
    B, Cc, H, W = alpha.shape
    if target.dim() == 4 and target.size(1) == 1:
        target = target[:, 0]
    if tuple(target.shape) != (B, H, W):
        raise ValueError(f"target shape {tuple(target.shape)} != {(B, H, W)}")
    target = _lib.as_buffer(target.to(alpha.device), torch.int64, "target")
    if keep_mask is not None:
        keep_mask = _lib.as_buffer(keep_mask.to(alpha.device), torch.bool, "keep_mask")
        if keep_mask.numel() != B * H * W:
            raise ValueError("keep_mask must have B*H*W elements")
    ign = [int(v) for v in ignore]
    h_ign = (_lib.C.c_int64 * max(1, len(ign)))(*ign) if ign else None
    if sums is None:
        sums = torch.zeros(3, dtype=torch.float64, device=alpha.device)
    g_mse = torch.empty_like(alpha) if (want_grad and want_mse) else None
    g_kl = torch.empty_like(alpha) if (want_grad and want_kl) else None
    rc = _lib.lib().slu_dirichlet_loss(_lib.ptr(alpha), _lib.ptr(target), _lib.ptr(keep_mask), B, Cc, H * W,
                                       h_ign, len(ign), float(eps_mse), float(eps_kl), int(want_mse), int(want_kl),
                                       _lib.ptr(sums), _lib.ptr(g_mse), _lib.ptr(g_kl), _lib.stream_ptr())
    _lib.check(rc, "slu_dirichlet_loss")
    return {"sums": sums, "grad_mse": g_mse, "grad_kl": g_kl}


TERM_NLL, TERM_DIGAMMA_CE, TERM_BRIER = 2, 3, 4


@_lib.device_guard
def dirichlet_term(alpha: torch.Tensor, target: torch.Tensor, term: int, *, ignore=(), keep_mask: Optional[torch.Tensor] = None,
                   eps: float = 1e-12, s_ref: Optional[float] = None, want_grad: bool = True) -> dict:
    """One alternative data-fit term (slu_dirichlet_term): sums float64[2] (sum | n_valid) and grad [B,C,H,W]."""
    _lib.require_cuda()
    alpha = _lib.as_buffer(alpha, torch.float32, "alpha")
    B, Cc, H, W = alpha.shape
    if target.dim() == 4 and target.size(1) == 1:
        target = target[:, 0]
    if tuple(target.shape) != (B, H, W):
        raise ValueError(f"target shape {tuple(target.shape)} != {(B, H, W)}")
    target = _lib.as_buffer(target.to(alpha.device), torch.int64, "target")
    if keep_mask is not None:
        keep_mask = _lib.as_buffer(keep_mask.to(alpha.device), torch.bool, "keep_mask")
    ign = [int(v) for v in ignore]
    h_ign = (_lib.C.c_int64 * max(1, len(ign)))(*ign) if ign else None
    sums = torch.zeros(2, dtype=torch.float64, device=alpha.device)
    grad = torch.empty_like(alpha) if want_grad else None
    rc = _lib.lib().slu_dirichlet_term(_lib.ptr(alpha), _lib.ptr(target), _lib.ptr(keep_mask), B, Cc, H * W, h_ign, len(ign),
                                       int(term), float(eps), -1.0 if s_ref is None else float(s_ref), _lib.ptr(sums),
                                       _lib.ptr(grad), _lib.stream_ptr())
    _lib.check(rc, "slu_dirichlet_term")
    return {"sums": sums, "grad": grad}


TERM_COMP_KL, TERM_WRONG_LOW, TERM_EVID_BAND, TERM_EVID_REG, TERM_KL_CONF = 5, 6, 7, 8, 9


def _mask_args(target, keep_mask, ignore, shape, device):
    B, H, W = shape
    if target is not None:
        if target.dim() == 4 and target.size(1) == 1:
            target = target[:, 0]
        if tuple(target.shape) != (B, H, W):
            raise ValueError(f"target shape {tuple(target.shape)} != {(B, H, W)}")
        target = _lib.as_buffer(target.to(device), torch.int64, "target")
    if keep_mask is not None:
        keep_mask = _lib.as_buffer(keep_mask.to(device), torch.bool, "keep_mask")
        if keep_mask.numel() != B * H * W:
            raise ValueError("keep_mask must have B*H*W elements")
    ign = [int(v) for v in ignore]
    h_ign = (_lib.C.c_int64 * max(1, len(ign)))(*ign) if ign else None
    return target, keep_mask, h_ign, len(ign)


@_lib.device_guard
def evidence_term(alpha: torch.Tensor, target: Optional[torch.Tensor], term: int, params, *, ignore=(),
                  keep_mask: Optional[torch.Tensor] = None, want_grad: bool = True) -> dict:
    """One of the TERM_COMP_KL / WRONG_LOW / EVID_BAND / EVID_REG / KL_CONF terms (slu_evidence_term):
    sums float64[2] (sum of per-pixel values | the term's denominator) and grad [B,C,H,W]."""
    _lib.require_cuda()
    alpha = _lib.as_buffer(alpha, torch.float32, "alpha")
    if alpha.dim() != 4:
        raise ValueError("alpha must be [B,C,H,W]")
    B, Cc, H, W = alpha.shape
    target, keep_mask, h_ign, n_ign = _mask_args(target, keep_mask, ignore, (B, H, W), alpha.device)
    prm = (_lib.C.c_float * len(params))(*[float(v) for v in params])
    sums = torch.zeros(2, dtype=torch.float64, device=alpha.device)
    grad = torch.empty_like(alpha) if want_grad else None
    rc = _lib.lib().slu_evidence_term(_lib.ptr(alpha), _lib.ptr(target), _lib.ptr(keep_mask), B, Cc, H * W, h_ign, n_ign,
                                      int(term), prm, len(params), _lib.ptr(sums), _lib.ptr(grad), _lib.stream_ptr())
    _lib.check(rc, "slu_evidence_term")
    return {"sums": sums, "grad": grad}


@_lib.device_guard
def logit_regularizer(logits: torch.Tensor, *, threshold: Optional[float] = None, target: Optional[torch.Tensor] = None,
                      ignore=(), keep_mask: Optional[torch.Tensor] = None, want_grad: bool = True) -> dict:
    """z^2 or relu(z - threshold)^2 summed over valid elements (slu_logit_regularizer): sums float64[2]
    (element sum | valid pixels) and grad [B,Cz,H,W]."""
    _lib.require_cuda()
    logits = _lib.as_buffer(logits, torch.float32, "logits")
    if logits.dim() != 4:
        raise ValueError("logits must be [B,C,H,W]")
    B, Cz, H, W = logits.shape
    target, keep_mask, h_ign, n_ign = _mask_args(target, keep_mask, ignore, (B, H, W), logits.device)
    sums = torch.zeros(2, dtype=torch.float64, device=logits.device)
    grad = torch.empty_like(logits) if want_grad else None
    rc = _lib.lib().slu_logit_regularizer(_lib.ptr(logits), _lib.ptr(target), _lib.ptr(keep_mask), B, Cz, H * W, h_ign, n_ign,
                                          int(threshold is not None), float(threshold or 0.0), _lib.ptr(sums),
                                          _lib.ptr(grad), _lib.stream_ptr())
    _lib.check(rc, "slu_logit_regularizer")
    return {"sums": sums, "grad": grad}


@_lib.device_guard
def evidential_loss_fused(outputs: torch.Tensor, target: torch.Tensor, *, w_mse: float = 1.0, w_kl: float = 0.05,
                          ignore=(), keep_mask: Optional[torch.Tensor] = None, temperature: float = 1.0,
                          eps_alpha: float = 1e-8, eps_mse: float = 1e-8, eps_kl: float = 1e-8, want_grad: bool = True,
                          count_reduce=None) -> dict:
    """Head output [B,C+1,H,W] -> sums float64[3] (sum mse | sum kl | n_valid) and d(loss)/d(outputs)
    (slu_evidential_loss_fused); loss = (w_mse*sums[0] + w_kl*sums[1]) / max(sums[2], 1).

    count_reduce: optional callable applied in place to the 1-element float64 valid-pixel count between the count kernel
    and the loss kernel (the batch-sharded step passes an all-reduce: gradient and sums then refer to the GLOBAL mean)."""
    _lib.require_cuda()
    outputs = _lib.as_buffer(outputs, torch.float32, "outputs")
    if outputs.dim() != 4:
        raise ValueError("outputs must be [B,C+1,H,W]")
    B, C1, H, W = outputs.shape
    if target.dim() == 4 and target.size(1) == 1:
        target = target[:, 0]
    if tuple(target.shape) != (B, H, W):
        raise ValueError(f"target shape {tuple(target.shape)} != {(B, H, W)}")
    target = _lib.as_buffer(target.to(outputs.device), torch.int64, "target")
    if keep_mask is not None:
        keep_mask = _lib.as_buffer(keep_mask.to(outputs.device), torch.bool, "keep_mask")
    ign = [int(v) for v in ignore]
    h_ign = (_lib.C.c_int64 * max(1, len(ign)))(*ign) if ign else None
    sums = torch.zeros(3, dtype=torch.float64, device=outputs.device)
    grad = torch.empty_like(outputs) if want_grad else None
    precounted = 0
    if count_reduce is not None:
        _lib.check(_lib.lib().slu_count_valid(_lib.ptr(target), _lib.ptr(keep_mask), B * H * W, h_ign, len(ign),
                                              _lib.ptr(sums[2:]), _lib.stream_ptr()), "slu_count_valid")
        count_reduce(sums[2:])
        precounted = 1
    rc = _lib.lib().slu_evidential_loss_fused(_lib.ptr(outputs), _lib.ptr(target), _lib.ptr(keep_mask), B, C1 - 1, H * W,
                                              h_ign, len(ign), float(temperature), float(eps_alpha), float(eps_mse),
                                              float(eps_kl), float(w_mse), float(w_kl), precounted, _lib.ptr(sums), _lib.ptr(grad),
                                              _lib.stream_ptr())
    _lib.check(rc, "slu_evidential_loss_fused")
    return {"sums": sums, "grad": grad}


@_lib.device_guard
def evidential_loss_step(outputs: torch.Tensor, target: torch.Tensor, count: torch.Tensor, state: torch.Tensor, *,
                         w_mse: float = 1.0, w_kl: float = 0.05, ignore=(), keep_mask: Optional[torch.Tensor] = None,
                         temperature: float = 1.0, eps_alpha: float = 1e-8, eps_mse: float = 1e-8, eps_kl: float = 1e-8,
                         precounted: bool = False, want_grad: bool = True, loss4: Optional[torch.Tensor] = None,
                         grad: Optional[torch.Tensor] = None) -> dict:
    """Training-step form of the fused loss (slu_evidential_loss_step): head output [B,C+1,H,W] -> loss4 float32[4]
    (loss | mse | kl | n_valid, written by the kernel) and d(loss)/d(outputs).  `count` float64[1] and `state` float64[3]
    are caller-owned, zero-initialised once; the kernel leaves them clean for the next step.  With precounted=True
    `count` already holds the (global) number of valid pixels (`count_valid` + all-reduce)."""
    _lib.require_cuda()
    outputs = _lib.as_buffer(outputs, torch.float32, "outputs")
    if outputs.dim() != 4:
        raise ValueError("outputs must be [B,C+1,H,W]")
    B, C1, H, W = outputs.shape
    if target.dim() == 4 and target.size(1) == 1:
        target = target[:, 0]
    if tuple(target.shape) != (B, H, W):
        raise ValueError(f"target shape {tuple(target.shape)} != {(B, H, W)}")
    target = _lib.as_buffer(target, torch.int64, "target")
    if keep_mask is not None:
        keep_mask = _lib.as_buffer(keep_mask, torch.bool, "keep_mask")
    if count.dtype != torch.float64 or count.numel() != 1 or state.dtype != torch.float64 or state.numel() != 3:
        raise ValueError("count must be float64[1] and state float64[3]")
    ign = [int(v) for v in ignore]
    h_ign = (_lib.C.c_int64 * max(1, len(ign)))(*ign) if ign else None
    if loss4 is None:
        loss4 = torch.empty(4, dtype=torch.float32, device=outputs.device)
    if grad is None and want_grad:
        grad = torch.empty_like(outputs)
    rc = _lib.lib().slu_evidential_loss_step(_lib.ptr(outputs), _lib.ptr(target), _lib.ptr(keep_mask), B, C1 - 1, H * W,
                                             h_ign, len(ign), float(temperature), float(eps_alpha), float(eps_mse),
                                             float(eps_kl), float(w_mse), float(w_kl), int(bool(precounted)), _lib.ptr(count),
                                             _lib.ptr(state), _lib.ptr(loss4), _lib.ptr(grad) if want_grad else None,
                                             _lib.stream_ptr())
    _lib.check(rc, "slu_evidential_loss_step")
    return {"loss4": loss4, "grad": grad if want_grad else None}


@_lib.device_guard
def count_valid(target: torch.Tensor, count: torch.Tensor, *, ignore=(), keep_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ADD the number of valid pixels of `target` (ignore list / keep mask as in the losses) to count float64[1]
    (slu_count_valid)."""
    _lib.require_cuda()
    target = _lib.as_buffer(target, torch.int64, "target")
    if keep_mask is not None:
        keep_mask = _lib.as_buffer(keep_mask, torch.bool, "keep_mask")
    ign = [int(v) for v in ignore]
    h_ign = (_lib.C.c_int64 * max(1, len(ign)))(*ign) if ign else None
    _lib.check(_lib.lib().slu_count_valid(_lib.ptr(target), _lib.ptr(keep_mask), target.numel(), h_ign, len(ign),
                                          _lib.ptr(count), _lib.stream_ptr()), "slu_count_valid")
    return count


@_lib.device_guard
def count_valid_exchange(target: torch.Tensor, count: torch.Tensor, peers, *, ignore=(), keep_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """count float64[1] <- the number of valid pixels of `target` summed over ALL ranks of `peers` (a dist.PeerCounter):
    slu_count_valid fused with the all-reduce of its result over NVLink peer memory, one kernel, capturable in a CUDA
    graph.  Every rank of the group must call it the same number of times."""
    _lib.require_cuda()
    target = _lib.as_buffer(target, torch.int64, "target")
    if keep_mask is not None:
        keep_mask = _lib.as_buffer(keep_mask, torch.bool, "keep_mask")
    ign = [int(v) for v in ignore]
    h_ign = (_lib.C.c_int64 * max(1, len(ign)))(*ign) if ign else None
    _lib.check(_lib.lib().slu_count_valid_exchange(_lib.ptr(target), _lib.ptr(keep_mask), target.numel(), h_ign, len(ign),
                                                   peers.boxes_array, peers.rank, peers.world, float(peers.timeout_s),
                                                   _lib.ptr(count), _lib.stream_ptr()), "slu_count_valid_exchange")
    return count


PEER_VEC_MAX = 512


@_lib.device_guard
def peer_allreduce_i64(a: torch.Tensor, b: Optional[torch.Tensor], peers, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[0 .. a.numel() + b.numel()) = the int64 vector (a, b) summed over ALL ranks of `peers` (a dist.PeerCounter) by one
    single-CTA kernel over NVLink peer memory (slu_peer_allreduce_i64); a and b are left untouched.  At most 512 elements."""
    _lib.require_cuda()
    a = _lib.as_buffer(a, torch.int64, "a").reshape(-1)
    n_b = 0
    if b is not None:
        b = _lib.as_buffer(b, torch.int64, "b").reshape(-1)
        n_b = b.numel()
    if out is None:
        out = torch.empty(a.numel() + n_b, dtype=torch.int64, device=a.device)
    _lib.check(_lib.lib().slu_peer_allreduce_i64(_lib.ptr(a), a.numel(), _lib.ptr(b), n_b, peers.boxes_array, peers.rank, peers.world,
                                                 float(peers.timeout_s), _lib.ptr(out), _lib.stream_ptr()), "slu_peer_allreduce_i64")
    return out


@_lib.device_guard
def special_functions(x: torch.Tensor) -> torch.Tensor:
    """[n,3] = lgamma, digamma, trigamma of x > 0 as the loss kernels evaluate them (slu_diag_special)."""
    _lib.require_cuda()
    x = _lib.as_buffer(x, torch.float32, "x").reshape(-1)
    out = torch.empty((x.numel(), 3), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().slu_diag_special(_lib.ptr(x), x.numel(), _lib.ptr(out), _lib.stream_ptr()), "slu_diag_special")
    return out


@_lib.device_guard
def confusion_ece(pred: torch.Tensor, labels: torch.Tensor, conf: Optional[torch.Tensor] = None, *,
                  num_classes: int, ignore_index: Optional[int] = None, edges=None,
                  confmat: Optional[torch.Tensor] = None, ece_bins: Optional[torch.Tensor] = None) -> None:
    """Standalone stage 4 on flat pred/labels (+ float32 conf): int64 maps (the reference's dtype, slu_confusion_ece)
    or, when BOTH are int32, the 12 B/px variant slu_confusion_ece_i32; the counters are identical."""
    _lib.require_cuda()
    i32 = pred.dtype == torch.int32 and labels.dtype == torch.int32
    it = torch.int32 if i32 else torch.int64
    pred = _lib.as_buffer(pred, it, "pred").reshape(-1)
    labels = _lib.as_buffer(labels, it, "labels").reshape(-1)
    if pred.numel() != labels.numel():
        raise ValueError("pred and labels differ in size")
    if conf is not None:
        conf = _lib.as_buffer(conf, torch.float32, "conf").reshape(-1)
        if conf.numel() != pred.numel():
            raise ValueError("conf and pred differ in size")
    n_bins, h_edges = 0, None
    if ece_bins is not None:
        n_bins = ece_bins.size(1)
        e = uniform_edges(n_bins) if edges is None else np.asarray(edges, dtype=np.float32)
        h_edges = _lib.edges_array(e)
    fn = _lib.lib().slu_confusion_ece_i32 if i32 else _lib.lib().slu_confusion_ece
    rc = fn(_lib.ptr(pred), _lib.ptr(labels), _lib.ptr(conf), pred.numel(), int(num_classes),
            0 if ignore_index is None else 1, 0 if ignore_index is None else int(ignore_index),
            n_bins, h_edges, _lib.ptr(confmat), _lib.ptr(ece_bins), _lib.stream_ptr())
    _lib.check(rc, "slu_confusion_ece")


SCORE_BINS = 60000          # divisible by 10, 15, 20, 50, 100: the usual coarse binnings fall on fine edges


def new_score_hist(device, n_score_bins: int = SCORE_BINS) -> torch.Tensor:
    """[2, n_score_bins] int64: row 0 correct pixels, row 1 wrong pixels, per score bin."""
    return torch.zeros((2, n_score_bins), dtype=torch.int64, device=device)


HYBRID_BINS = 1 << 20      # bins of the hybrid scheme (slu_score_hist_hybrid)


def hybrid_bin_lower_edges() -> np.ndarray:
    """float64 lower edges of the 2^20 hybrid bins plus the closing 1.0 (length 2^20 + 1): 2048 mantissa bins per binary
    octave from 2^-36 to 1/16 (bin 0 starts at 0), then uniform steps of 2^-20."""
    e = np.arange(32, dtype=np.float64)[:, None] - 36.0
    low = (np.exp2(e) * (1.0 + np.arange(2048, dtype=np.float64)[None, :] / 2048.0)).reshape(-1)
    low[0] = 0.0
    return np.concatenate([low, np.arange(65536, HYBRID_BINS + 1, dtype=np.float64) / HYBRID_BINS])


@_lib.device_guard
def score_hist(score: torch.Tensor, pred: torch.Tensor, labels: torch.Tensor, hist: torch.Tensor, ignore=(),
               hybrid: bool = False) -> None:
    """Accumulate (score, pred != label) pairs into `hist` [2,M] (slu_score_hist; hybrid=True: the 2^20 hybrid bins of
    slu_score_hist_hybrid, M must be HYBRID_BINS)."""
    _lib.require_cuda()
    score = _lib.as_buffer(score, torch.float32, "score").reshape(-1)
    pred = _lib.as_buffer(pred, torch.int64, "pred").reshape(-1)
    labels = _lib.as_buffer(labels, torch.int64, "labels").reshape(-1)
    if not (score.numel() == pred.numel() == labels.numel()):
        raise ValueError("score, pred and labels differ in size")
    if hist.dtype != torch.int64 or not hist.is_contiguous() or hist.dim() != 2 or hist.size(0) != 2:
        raise ValueError("hist must be a contiguous [2,M] int64 tensor")
    ign = [int(v) for v in ignore]
    h_ign = (_lib.C.c_int64 * max(1, len(ign)))(*ign) if ign else None
    if hybrid:
        if hist.size(1) != HYBRID_BINS:
            raise ValueError(f"the hybrid scheme needs hist [2,{HYBRID_BINS}]")
        rc = _lib.lib().slu_score_hist_hybrid(_lib.ptr(score), _lib.ptr(pred), _lib.ptr(labels), score.numel(),
                                              h_ign, len(ign), _lib.ptr(hist), _lib.stream_ptr())
    else:
        rc = _lib.lib().slu_score_hist(_lib.ptr(score), _lib.ptr(pred), _lib.ptr(labels), score.numel(), hist.size(1),
                                       h_ign, len(ign), _lib.ptr(hist), _lib.stream_ptr())
    _lib.check(rc, "slu_score_hist")


@_lib.device_guard
def class_score_hist(score: torch.Tensor, labels: torch.Tensor, hist: torch.Tensor, sum_fx: torch.Tensor) -> None:
    """Accumulate scores per label class (slu_class_score_hist): hist [C,M] int64, sum_fx [C] int64 (2^-32 units)."""
    _lib.require_cuda()
    score = _lib.as_buffer(score, torch.float32, "score").reshape(-1)
    labels = _lib.as_buffer(labels, torch.int64, "labels").reshape(-1)
    if score.numel() != labels.numel():
        raise ValueError("score and labels differ in size")
    rc = _lib.lib().slu_class_score_hist(_lib.ptr(score), _lib.ptr(labels), score.numel(), hist.size(0), hist.size(1),
                                         _lib.ptr(hist), _lib.ptr(sum_fx), _lib.stream_ptr())
    _lib.check(rc, "slu_class_score_hist")


def _offsets_array(offsets):
    off = np.ascontiguousarray(np.asarray(offsets, dtype=np.int64))
    return off, off.ctypes.data_as(_lib.C.c_void_p)


@_lib.device_guard
def project_batch(xyzi: torch.Tensor, raw_label: Optional[torch.Tensor], offsets, H: int, W: int, *,
                  lut: Optional[torch.Tensor] = None, theta_range=None, farthest_wins: bool = False,
                  yaw_deg=None, want_img: bool = True, want_label: bool = True,
                  workspace: Optional[torch.Tensor] = None) -> dict:
    """Batched stage 1 (slu_project_batch).

    xyzi [n_total,4] float32 CUDA (scans concatenated), raw_label [n_total] uint32-as-int32 CUDA or None,
    offsets: host sequence of B+1 point offsets; yaw_deg: optional per-scan yaw angles in degrees (the
    loaders' rotate_z augmentation).  Returns img [B,6,H,W] (x,y,z,range,intensity,label),
    label [B,H,W] int64 (train ids, the loaders' `semantics`), pix [n_total] int32, winner [B,H,W] int32,
    theta [B,2] float64, diag [B,2] int32.
    """
    _lib.require_cuda()
    xyzi = _lib.as_buffer(xyzi, torch.float32, "xyzi")
    if xyzi.dim() != 2 or xyzi.size(1) != 4:
        raise ValueError("xyzi must be [n_total,4]")
    off, off_p = _offsets_array(offsets)
    B = off.shape[0] - 1
    n_total = xyzi.size(0)
    dev = xyzi.device
    if raw_label is not None:
        if raw_label.dtype not in (torch.int32, torch.uint32):
            raise ValueError("raw_label must be int32/uint32 (the .label file's uint32 words)")
        raw_label = raw_label.contiguous()
        if raw_label.numel() != n_total:
            raise ValueError("raw_label and xyzi differ in length")
    if lut is not None:
        lut = _lib.as_buffer(lut, torch.int32, "lut")
        if lut.numel() != 65536:
            raise ValueError("lut must have 65536 entries")
    HW = H * W
    need = _lib.lib().slu_project_workspace_bytes(n_total, B, HW)
    if need < 0:
        _lib.check(int(need), "slu_project_workspace_bytes")
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(int(need), dtype=torch.uint8, device=dev)
    img = torch.empty((B, 6, H, W), dtype=torch.float32, device=dev) if want_img else None
    label = torch.empty((B, H, W), dtype=torch.int64, device=dev) if want_label else None
    pix = torch.empty((n_total,), dtype=torch.int32, device=dev)
    winner = torch.empty((B, H, W), dtype=torch.int32, device=dev)
    theta = torch.empty((B, 2), dtype=torch.float64, device=dev)
    diag = torch.empty((B, 2), dtype=torch.int32, device=dev)
    yaw = None
    if yaw_deg is not None:
        a = np.radians(np.broadcast_to(np.asarray(yaw_deg, dtype=np.float64), (B,)))
        yaw = torch.from_numpy(np.stack([np.cos(a), np.sin(a)], axis=1).copy()).to(dev)
    use_range = theta_range is not None
    lo, hi = (float(theta_range[0]), float(theta_range[1])) if use_range else (0.0, 0.0)
    rc = _lib.lib().slu_project_batch(_lib.ptr(xyzi), _lib.ptr(raw_label), _lib.ptr(lut), off_p, n_total, B, H, W,
                                      int(use_range), lo, hi, int(farthest_wins), _lib.ptr(yaw), _lib.ptr(workspace),
                                      _lib.ptr(img), _lib.ptr(label), _lib.ptr(pix), _lib.ptr(winner), _lib.ptr(theta),
                                      _lib.ptr(diag), _lib.stream_ptr())
    _lib.check(rc, "slu_project_batch")
    return {"img": img, "label": label, "pix": pix, "winner": winner, "theta": theta, "diag": diag, "workspace": workspace}


@_lib.device_guard
def project_points(pc: torch.Tensor, H: int, W: int, *, theta_range=None, farthest_wins: bool = False,
                   want_img: bool = True, workspace: Optional[torch.Tensor] = None, bins_h=None) -> dict:
    """Generic stage 1 (slu_project_points): pc [N,Cin] float64 CUDA -> img [H,W,Cin] float32.
    `workspace`: a uint8 CUDA tensor from an earlier call (returned under "workspace") is reused when large enough."""
    _lib.require_cuda()
    pc = _lib.as_buffer(pc, torch.float64, "pc")
    if pc.dim() != 2 or pc.size(1) < 3:
        raise ValueError("pc must be [N,Cin] with Cin >= 3")
    N, Cin = pc.shape
    dev = pc.device
    need = _lib.lib().slu_project_workspace_bytes(N, 1, H * W)
    if workspace is None or workspace.numel() < need or workspace.device != dev:
        workspace = torch.empty(int(need), dtype=torch.uint8, device=dev)
    img = torch.empty((H, W, Cin), dtype=torch.float32, device=dev) if want_img else None
    pix = torch.empty((N,), dtype=torch.int32, device=dev)
    winner = torch.empty((H, W), dtype=torch.int32, device=dev)
    theta = torch.empty((1, 2), dtype=torch.float64, device=dev)
    diag = torch.empty((1, 2), dtype=torch.int32, device=dev)
    use_range = theta_range is not None
    lo, hi = (float(theta_range[0]), float(theta_range[1])) if use_range else (0.0, 0.0)
    edges, increasing = None, 0
    if bins_h is not None:                        # the reference's bins_h: H monotone row edges (np.digitize's requirement)
        e = np.asarray(bins_h, dtype=np.float64).reshape(-1)
        if e.shape[0] != H:
            raise ValueError(f"bins_h must have height={H} entries")
        d = np.diff(e)
        if np.all(d >= 0):
            increasing = 1
        elif np.all(d <= 0):
            e = e[::-1]
        else:
            raise ValueError("bins must be monotonically increasing or decreasing")
        edges = torch.from_numpy(np.ascontiguousarray(e)).to(dev)
    rc = _lib.lib().slu_project_points_bins(_lib.ptr(pc), N, Cin, H, W, int(use_range), lo, hi, int(farthest_wins),
                                            _lib.ptr(edges), increasing, _lib.ptr(workspace), _lib.ptr(img), _lib.ptr(pix),
                                            _lib.ptr(winner), _lib.ptr(theta), _lib.ptr(diag), _lib.stream_ptr())
    _lib.check(rc, "slu_project_points")
    return {"img": img, "pix": pix, "winner": winner, "theta": theta, "diag": diag, "workspace": workspace}


@_lib.device_guard
def organized_planes(xyzi: torch.Tensor, raw_label: Optional[torch.Tensor], H: int, W: int, *, lut: Optional[torch.Tensor] = None,
                     flip=None, col_shift=None, yaw_deg=None) -> dict:
    """Organised clouds (slu_organized_planes): xyzi [B*H*W,4] float32 -> img [B,6,H,W] planes, missing [B]."""
    _lib.require_cuda()
    xyzi = _lib.as_buffer(xyzi, torch.float32, "xyzi").reshape(-1, 4)
    if xyzi.size(0) % (H * W) != 0:
        raise ValueError("xyzi must hold a whole number of H*W clouds")
    B = xyzi.size(0) // (H * W)
    dev = xyzi.device
    if raw_label is not None:
        raw_label = raw_label.contiguous()
    if lut is not None:
        lut = _lib.as_buffer(lut, torch.int32, "lut")
    h_flip = None
    if flip is not None:
        f = np.ascontiguousarray(np.broadcast_to(np.asarray(flip, dtype=np.uint8), (B,)))
        h_flip = f.ctypes.data_as(_lib.C.c_void_p)
    shift = None if col_shift is None else torch.as_tensor(np.broadcast_to(np.asarray(col_shift, dtype=np.int32), (B,)).copy()).to(dev)
    yaw = None
    if yaw_deg is not None:
        a = np.radians(np.broadcast_to(np.asarray(yaw_deg, dtype=np.float64), (B,)))
        yaw = torch.from_numpy(np.stack([np.cos(a), np.sin(a)], axis=1).copy()).to(dev)
    img = torch.empty((B, 6, H, W), dtype=torch.float32, device=dev)
    missing = torch.empty((B,), dtype=torch.int32, device=dev)
    rc = _lib.lib().slu_organized_planes(_lib.ptr(xyzi), _lib.ptr(raw_label), _lib.ptr(lut), B, H, W, h_flip, _lib.ptr(shift),
                                         _lib.ptr(yaw), _lib.ptr(img), _lib.ptr(missing), _lib.stream_ptr())
    _lib.check(rc, "slu_organized_planes")
    return {"img": img, "missing": missing}


@_lib.device_guard
def frame_tensors(img: torch.Tensor, out_hw=None, flip=None, norm_factor: float = 0.25, want_normals: bool = True,
                  drop_empty_rows: bool = False, label: Optional[torch.Tensor] = None) -> dict:
    """Loader glue on the device (slu_frame_tensors): img [B,6,H,W] planes -> the loaders' five tensors,
    stacked over B: range [B,1,h,w], reflectivity [B,1,h,w], xyz [B,3,h,w], normals [B,3,h,w], semantics
    [B,1,h,w] int64.  out_hw=(h,w) resizes with cv2's INTER_NEAREST rule; flip: per-scan booleans;
    drop_empty_rows removes image rows without any return first (WADS) and adds "rows_kept" [B] int32.

    label: the int64 label map [B,H,W] project_batch returned next to img.  With it, and nothing to resize, flip or
    drop, no plane is copied: range / reflectivity / xyz are VIEWS of img's planes (values, shapes and dtypes as above;
    contiguous per scan, batch stride 6*H*W), semantics a view of label, and only the normals kernel is launched."""
    _lib.require_cuda()
    img = _lib.as_buffer(img, torch.float32, "img")
    if img.dim() != 4 or img.size(1) != 6:
        raise ValueError("img must be [B,6,H,W]")
    B, _, Hs, Ws = img.shape
    Hd, Wd = (Hs, Ws) if out_hw is None else (int(out_hw[0]), int(out_hw[1]))
    dev = img.device
    if label is not None and (Hd, Wd) == (Hs, Ws) and flip is None and not drop_empty_rows:
        if label.dtype != torch.int64 or tuple(label.shape) != (B, Hs, Ws) or label.device != dev:
            raise ValueError("label must be the int64 [B,H,W] map that belongs to img")
        return {"range": img[:, 3:4], "reflectivity": img[:, 4:5], "xyz": img[:, 0:3],
                "normals": frame_normals(img, norm_factor) if want_normals else None,
                "semantics": label.unsqueeze(1)}
    h_flip = None
    if flip is not None:
        f = np.ascontiguousarray(np.broadcast_to(np.asarray(flip, dtype=np.uint8), (B,)))
        h_flip = f.ctypes.data_as(_lib.C.c_void_p)
    out = {"range": torch.empty((B, 1, Hd, Wd), dtype=torch.float32, device=dev),
           "reflectivity": torch.empty((B, 1, Hd, Wd), dtype=torch.float32, device=dev),
           "xyz": torch.empty((B, 3, Hd, Wd), dtype=torch.float32, device=dev),
           "normals": torch.empty((B, 3, Hd, Wd), dtype=torch.float32, device=dev) if want_normals else None,
           "semantics": torch.empty((B, 1, Hd, Wd), dtype=torch.int64, device=dev)}
    rowmap = torch.empty((B, Hs + 1), dtype=torch.int32, device=dev) if drop_empty_rows else None
    rc = _lib.lib().slu_frame_tensors(_lib.ptr(img), B, Hs, Ws, Hd, Wd, h_flip, float(norm_factor),
                                      _lib.ptr(out["range"]), _lib.ptr(out["reflectivity"]), _lib.ptr(out["xyz"]),
                                      _lib.ptr(out["normals"]), _lib.ptr(out["semantics"]), _lib.ptr(rowmap),
                                      int(out_hw is not None), _lib.stream_ptr())
    _lib.check(rc, "slu_frame_tensors")
    if drop_empty_rows:
        out["rows_kept"] = rowmap[:, 0]
    return out


@_lib.device_guard
def frame_normals(xyz: torch.Tensor, norm_factor: float = 0.25) -> torch.Tensor:
    """build_normal_xyz on the device (slu_frame_normals): xyz [B,3,H,W] or the projection image [B,6,H,W] (its first
    three planes are read in place) -> normals [B,3,H,W]."""
    _lib.require_cuda()
    xyz = _lib.as_buffer(xyz, torch.float32, "xyz")
    if xyz.dim() != 4 or xyz.size(1) not in (3, 6):
        raise ValueError("xyz must be [B,3,H,W] or the projection image [B,6,H,W]")
    B, Cx, H, W = xyz.shape
    out = torch.empty((B, 3, H, W), dtype=torch.float32, device=xyz.device)
    rc = _lib.lib().slu_frame_normals(_lib.ptr(xyz), B, H, W, Cx * H * W, float(norm_factor), _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "slu_frame_normals")
    return out


@_lib.device_guard
def backproject(label_img: torch.Tensor, pix: torch.Tensor, offsets) -> torch.Tensor:
    """Stage 2 (slu_backproject): label_img [B,H,W] int64, pix [n_total] int32 -> [n_total] int64."""
    _lib.require_cuda()
    if label_img.dim() == 2:
        label_img = label_img.unsqueeze(0)
    label_img = _lib.as_buffer(label_img, torch.int64, "label_img")
    pix = _lib.as_buffer(pix, torch.int32, "pix")
    off, off_p = _offsets_array(offsets)
    B = off.shape[0] - 1
    if label_img.size(0) != B:
        raise ValueError("label_img batch size does not match offsets")
    out = torch.empty((pix.numel(),), dtype=torch.int64, device=pix.device)
    rc = _lib.lib().slu_backproject(_lib.ptr(label_img), _lib.ptr(pix), off_p, pix.numel(), B,
                                    label_img.size(1) * label_img.size(2), _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "slu_backproject")
    return out


def ece_from_bins(ece_bins: torch.Tensor):
    """(ece, mce, n, acc, avg_conf) from the [3,n_bins] int64 state, following src/metrics/ece.py:139-168."""
    b = ece_bins.detach().cpu().numpy()
    n = b[0].astype(np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        acc = np.where(n > 0, b[1].astype(np.float64) / n, np.nan)
        avg = np.where(n > 0, (b[2].astype(np.float64) / 4294967296.0) / n, np.nan)
    if n.sum() == 0:
        return float("nan"), float("nan"), b[0], acc, avg
    gap = np.abs(np.nan_to_num(acc, nan=0.0) - np.nan_to_num(avg, nan=0.0))
    ece = float(np.sum((n / max(1.0, n.sum())) * gap))
    mce = float(np.max(gap[n > 0]))
    return ece, mce, b[0], acc, avg


LOG = math.log
