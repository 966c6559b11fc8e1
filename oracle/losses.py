"""CPU restatement of the two evidential loss terms the shipped configs enable.  TEST INFRASTRUCTURE ONLY.

`DirichletMSELoss`          : src/losses/dirichlet_losses.py:317-385
`KL_offClasses_to_uniform`  : src/losses/regularizers.py:291-389
mask helper `_valid_mask`   : src/losses/dirichlet_losses.py:15-70
(weights w_mse=1.0, w_kl=0.05: src/configs/SemanticKitti_default.yaml:50-62).
Plain differentiable torch so autograd supplies the reference gradients.
"""
from __future__ import annotations

import torch


def valid_mask(target: torch.Tensor, ignore_index):
    if target.dim() == 4 and target.size(1) == 1:
        target = target[:, 0]
    if ignore_index is None:
        return torch.ones_like(target, dtype=torch.bool)
    if torch.is_tensor(ignore_index):
        if ignore_index.dtype == torch.bool:
            return ignore_index
        return ~torch.isin(target, ignore_index.to(target.dtype))
    if isinstance(ignore_index, int):
        return target != ignore_index
    ids = list(ignore_index)
    if not ids:
        return torch.ones_like(target, dtype=torch.bool)
    return ~torch.isin(target, torch.as_tensor(ids, dtype=target.dtype))


def _one_hot_like(alpha, target):
    y = torch.zeros_like(alpha)
    y.scatter_(1, target.unsqueeze(1), 1.0)
    return y


def dirichlet_mse(alpha: torch.Tensor, target: torch.Tensor, ignore_index=None, eps: float = 1e-8):
    if target.dim() == 4 and target.size(1) == 1:
        target = target[:, 0]
    target = target.long()
    valid = valid_mask(target, ignore_index)
    if valid.sum() == 0 or alpha.shape[1] <= 2:
        return alpha.sum() * 0.0
    alpha0 = alpha.sum(dim=1, keepdim=True)
    p_hat = alpha / (alpha0 + eps)
    y = _one_hot_like(alpha, target)
    sq_err = (y - p_hat) ** 2
    var = alpha * (alpha0 - alpha) / ((alpha0 * alpha0 + eps) * (alpha0 + 1.0))
    per_pix = (sq_err + var).sum(dim=1)
    return (per_pix * valid.float()).sum() / valid.float().sum().clamp_min(1.0)


def kl_offclasses_to_uniform(alpha: torch.Tensor, target: torch.Tensor, ignore_index=None, eps: float = 1e-8):
    """with_conf_weighting=False branch (the constructor default, regularizers.py:300)."""
    if target.dim() == 4 and target.size(1) == 1:
        target = target[:, 0]
    target = target.long()
    valid = valid_mask(target, ignore_index)
    if valid.sum() == 0:
        return alpha.sum() * 0.0
    y = _one_hot_like(alpha, target)
    a_t = y + (1.0 - y) * alpha
    C = alpha.shape[1]
    flat = a_t.permute(0, 2, 3, 1).reshape(-1, C)[valid.reshape(-1)]
    a = flat.clamp_min(eps)
    s = a.sum(dim=1, keepdim=True)
    term1 = torch.lgamma(s) - torch.lgamma(a).sum(dim=1, keepdim=True)
    term2 = ((a - 1.0) * (torch.digamma(a) - torch.digamma(s))).sum(dim=1, keepdim=True)
    return (term1 + term2).squeeze(1).mean()


def _masked_mean(per_pix, valid):
    w = valid.float().to(per_pix.dtype)
    return (per_pix * w).sum() / w.sum().clamp_min(1.0)


def nll_dirichlet_categorical(alpha, target, ignore_index=None, eps: float = 1e-12):
    """src/losses/dirichlet_losses.py:73-119."""
    if target.dim() == 4 and target.size(1) == 1:
        target = target[:, 0]
    target = target.long()
    valid = valid_mask(target, ignore_index)
    if valid.sum() == 0:
        return alpha.sum() * 0.0
    a0 = alpha.sum(dim=1)
    ay = alpha.gather(1, target.unsqueeze(1)).squeeze(1)
    return _masked_mean(-(torch.log(ay + eps) - torch.log(a0 + eps)), valid)


def digamma_dirichlet_ce(alpha, target, ignore_index=None):
    """src/losses/dirichlet_losses.py:122-167."""
    if target.dim() == 4 and target.size(1) == 1:
        target = target[:, 0]
    target = target.long()
    valid = valid_mask(target, ignore_index)
    if valid.sum() == 0:
        return alpha.sum() * 0.0
    a0 = alpha.sum(dim=1)
    ay = alpha.gather(1, target.unsqueeze(1)).squeeze(1)
    return _masked_mean(torch.digamma(a0) - torch.digamma(ay), valid)


def brier_dirichlet(alpha, target, ignore_index=None, s_ref=None, eps: float = 1e-12):
    """src/losses/dirichlet_losses.py:174-220."""
    if target.dim() == 4 and target.size(1) == 1:
        target = target[:, 0]
    target = target.long()
    valid = valid_mask(target, ignore_index)
    if valid.sum() == 0:
        return alpha.sum() * 0.0
    a0 = alpha.sum(dim=1, keepdim=True)
    p_hat = alpha / (a0 + eps)
    sum_p2 = (p_hat * p_hat).sum(dim=1, keepdim=True)
    s = a0 if s_ref is None else torch.as_tensor(float(s_ref), dtype=alpha.dtype)
    sum_ep2 = (s * sum_p2 + 1.0) / (s + 1.0)
    ep_y = p_hat.gather(1, target.unsqueeze(1))
    return _masked_mean((sum_ep2 - 2.0 * ep_y + 1.0).squeeze(1), valid)
