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


# ---- the remaining terms; plain differentiable torch, any dtype -----------------------------
def _prep(target):
    if target.dim() == 4 and target.size(1) == 1:
        target = target[:, 0]
    return target.long()


def complement_kl_uniform(alpha, target, ignore_index=0, gamma=2.0, tau=0.55, sigma=0.12, s_target=None,
                          normalize=True, eps=1e-8, detach_uncert=True):
    """src/losses/dirichlet_losses.py:228-314."""
    import math
    target = _prep(target)
    valid = valid_mask(target, ignore_index)
    C = alpha.shape[1]
    if valid.sum() == 0 or C <= 2:
        return alpha.sum() * 0.0
    a0 = alpha.sum(dim=1, keepdim=True) + eps
    p = alpha / a0
    tgt = torch.where(valid, target, torch.zeros_like(target)).unsqueeze(1)
    py = p.gather(1, tgt).clamp_min(eps)
    p_off = p.clone()
    p_off.scatter_(1, tgt, 0.0)
    tilde = p_off / (1.0 - py).clamp_min(eps)
    kl_u = (tilde * tilde.clamp_min(eps).log()).sum(dim=1) + math.log(C - 1)
    if normalize:
        kl_u = kl_u / math.log(C - 1)
    pg = py.detach() if detach_uncert else py
    w = ((1.0 - pg).pow(gamma) * torch.sigmoid((tau - pg) / sigma)).squeeze(1)
    if s_target is not None:
        w = w * (float(s_target) / (a0.detach().squeeze(1) + float(s_target)))
    wm = valid.to(alpha.dtype)
    return (w * kl_u * wm).sum() / wm.sum().clamp_min(1.0)


def _mean_over_valid(x, mask):
    """src/losses/regularizers.py:56-69."""
    if mask is None:
        return x.mean()
    m = mask.to(x.dtype)
    if x.dim() == 4:
        m = m.unsqueeze(1)
    return (x * m).sum() / m.sum().clamp_min(1e-8)


def logit_regularizer(logits, threshold=None, ignore_index=None, mask=None, target=None):
    """src/losses/regularizers.py:75-110."""
    if mask is None and target is not None:
        mask = valid_mask(target, ignore_index)
    per = logits.pow(2) if threshold is None else torch.relu(logits - float(threshold)).pow(2)
    return _mean_over_valid(per, mask)


def evidence_reg_band(alpha, s_target, band=0.10, ignore_index=None, mask=None, target=None):
    """src/losses/regularizers.py:116-147."""
    if mask is None and target is not None:
        mask = valid_mask(target, ignore_index)
    a0 = alpha.sum(dim=1) + 1e-8
    over = torch.relu(torch.log(a0 / (s_target * (1.0 + band))))
    under = torch.relu(torch.log((s_target * (1.0 - band)) / a0))
    return _mean_over_valid(over.pow(2) + under.pow(2), mask)


def evidence_reg(alpha, s_target, mode="log_squared", margin=0.1, scale_correct=False, ignore_index=None,
                 mask=None, target=None):
    """src/losses/regularizers.py:149-212."""
    if mask is None and target is not None:
        mask = valid_mask(target, ignore_index)
    a0 = alpha.sum(dim=1) + 1e-8
    if mode == "log_squared":
        per = torch.log(a0 / s_target).pow(2)
        if scale_correct:
            per = (a0 / s_target) * per
    elif mode == "one_sided":
        per = torch.relu(a0 - s_target * (1.0 + margin)).pow(2)
    else:
        per = (a0 - s_target).pow(2)
    return _mean_over_valid(per, mask)


def wrong_low_evidence(alpha, target, ignore_index=None, s_low=0.0, margin=0.05, soft_margin_k=0.08, eps=1e-8):
    """src/losses/regularizers.py:218-289."""
    import math
    target = _prep(target)
    valid = valid_mask(target, ignore_index)
    if valid.sum() == 0:
        return alpha.sum() * 0.0
    C = alpha.shape[1]
    a0 = alpha.sum(dim=1, keepdim=True).clamp_min(eps)
    p = alpha / a0
    with torch.no_grad():
        pd = p.detach()
        wrong = pd.argmax(dim=1) != target
        py = pd.gather(1, target.unsqueeze(1)).clamp_min(eps)
        pmax = pd.max(dim=1, keepdim=True).values.clamp_min(eps)
        m = (pmax - py).squeeze(1)
        if margin > 0.0:
            gate = torch.sigmoid((m - margin) / soft_margin_k) if soft_margin_k > 0.0 else (m > margin).to(p.dtype)
        else:
            gate = torch.ones_like(m)
        gw = wrong.to(p.dtype) * gate * valid.to(p.dtype)
    per = torch.relu(a0.log().squeeze(1) - math.log(C + s_low + eps)).pow(2) * gw
    return per.sum() / gw.sum().clamp_min(1.0)


def kl_offclasses_conf_weighted(alpha, target, ignore_index=None, gamma=1.0, eps=1e-8):
    """KL_offClasses_to_uniform(with_conf_weighting=True), src/losses/regularizers.py:342-385."""
    target = _prep(target)
    valid = valid_mask(target, ignore_index)
    if valid.sum() == 0:
        return alpha.sum() * 0.0
    y = _one_hot_like(alpha, target)
    C = alpha.shape[1]
    vf = valid.reshape(-1)
    a = (y + (1.0 - y) * alpha).permute(0, 2, 3, 1).reshape(-1, C)[vf].clamp_min(eps)
    s = a.sum(dim=1, keepdim=True)
    kl = (torch.lgamma(s) - torch.lgamma(a).sum(dim=1, keepdim=True)
          + ((a - 1.0) * (torch.digamma(a) - torch.digamma(s))).sum(dim=1, keepdim=True)).squeeze(1)
    a0 = alpha.sum(dim=1, keepdim=True)
    py = (alpha / (a0 + eps)).gather(1, target.unsqueeze(1)).squeeze(1)
    w = ((1.0 - py).clamp(0.0, 1.0) ** gamma).reshape(-1)[vf].detach()
    return (kl * w).sum() / w.sum().clamp_min(1.0)
