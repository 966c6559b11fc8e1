"""autograd bridge: one fused forward+backward kernel launch per loss term."""
from __future__ import annotations

import torch

from .. import ops
from ._mask import split_ignore

MAX_IDS = 8


class _DirichletTerm(torch.autograd.Function):
    """term 0 = expected-squared-error (DirichletMSELoss), term 1 = KL of off classes to uniform."""

    @staticmethod
    def forward(ctx, alpha, target, term: int, ignore_index, eps: float):
        ids, keep = split_ignore(ignore_index)
        if len(ids) > MAX_IDS:                       # rare: fold a long id list into a keep mask
            keep = ~torch.isin(target, torch.as_tensor(ids, device=target.device, dtype=target.dtype))
            ids = ()
        need_grad = ctx.needs_input_grad[0]
        r = ops.dirichlet_loss(alpha.detach(), target, ignore=ids, keep_mask=keep, eps_mse=eps, eps_kl=eps,
                               want_mse=(term == 0), want_kl=(term == 1), want_grad=need_grad)
        n = r["sums"][2].clamp_min(1.0)
        loss = (r["sums"][term] / n).to(alpha.dtype)
        if need_grad:
            ctx.save_for_backward(r["grad_mse"] if term == 0 else r["grad_kl"], n)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        g, n = ctx.saved_tensors
        scale = (grad_out.double() / n).to(g.dtype)
        return g * scale, None, None, None, None


class _DirichletSingleTerm(torch.autograd.Function):
    """NLL / digamma-CE / Brier (ops.TERM_*): one kernel launch, analytic gradient saved for backward."""

    @staticmethod
    def forward(ctx, alpha, target, term: int, ignore_index, eps: float, s_ref):
        ids, keep = split_ignore(ignore_index)
        if len(ids) > MAX_IDS:
            keep = ~torch.isin(target, torch.as_tensor(ids, device=target.device, dtype=target.dtype))
            ids = ()
        r = ops.dirichlet_term(alpha.detach(), target, term, ignore=ids, keep_mask=keep, eps=eps, s_ref=s_ref,
                               want_grad=ctx.needs_input_grad[0])
        n = r["sums"][1].clamp_min(1.0)
        if ctx.needs_input_grad[0]:
            ctx.save_for_backward(r["grad"], n)
        return (r["sums"][0] / n).to(alpha.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        g, n = ctx.saved_tensors
        return g * (grad_out.double() / n).to(g.dtype), None, None, None, None, None


def _ids_and_keep(target, ignore_index, mask=None):
    """(ignored ids, keep mask) from the reference's two ways of masking: `ignore_index` on target, or `mask=`."""
    if mask is not None:
        return (), mask
    if target is None:
        return (), None
    ids, keep = split_ignore(ignore_index)
    if len(ids) > MAX_IDS:
        keep = ~torch.isin(target, torch.as_tensor(ids, device=target.device, dtype=target.dtype))
        ids = ()
    return ids, keep


class _EvidenceTerm(torch.autograd.Function):
    """ops.TERM_COMP_KL / WRONG_LOW / EVID_BAND / EVID_REG / KL_CONF: value = sums[0] / max(sums[1], denom_min)."""

    @staticmethod
    def forward(ctx, alpha, target, term: int, params, ids, keep, denom_min: float):
        r = ops.evidence_term(alpha.detach(), target, term, params, ignore=ids, keep_mask=keep,
                              want_grad=ctx.needs_input_grad[0])
        n = r["sums"][1].clamp_min(denom_min)
        if ctx.needs_input_grad[0]:
            ctx.save_for_backward(r["grad"], n)
        return (r["sums"][0] / n).to(alpha.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        g, n = ctx.saved_tensors
        return g * (grad_out.double() / n).to(g.dtype), None, None, None, None, None, None


class _LogitReg(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, threshold, ids, keep):
        r = ops.logit_regularizer(logits.detach(), threshold=threshold, target=target, ignore=ids, keep_mask=keep,
                                  want_grad=ctx.needs_input_grad[0])
        if target is None and keep is None:           # no mask: plain mean over every element (regularizers.py:63-64)
            n = torch.full((), float(logits.numel()), dtype=torch.float64, device=logits.device)
        else:                                         # masked: element sum over the number of valid PIXELS (:65-69)
            n = r["sums"][1].clamp_min(1e-8)
        if ctx.needs_input_grad[0]:
            ctx.save_for_backward(r["grad"], n)
        return (r["sums"][0] / n).to(logits.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        g, n = ctx.saved_tensors
        return g * (grad_out.double() / n).to(g.dtype), None, None, None, None
