"""CPU restatement of the reference's uncertainty reductions.  TEST INFRASTRUCTURE ONLY.

MC-dropout branch : src/models/tester.py:412-451 (closures, not importable),
                    src/utils/mc_dropout.py:121-133.
Evidential branch : src/models/probability_helper.py:89-153,
                    src/metrics/auroc.py:55-63 (Dirichlet MI).
Same torch CPU ops in the same order as the reference, so the oracle agrees
with the reference to the last bit on the same torch build (checked in
tests/test_oracle_golden.py against tests/golden/).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch.special import digamma

EPS_MC = 1e-12       # src/models/tester.py:420,428
EPS_DIR = 1e-8       # src/models/probability_helper.py:14


@torch.no_grad()
def mc_reduce(mc_logits: torch.Tensor, eps: float = EPS_MC, log_softmax_exp: bool = False):
    """[T,B,C,H,W] logits -> p_bar [B,C,H,W], pred [B,H,W] int64, H_norm, MI_norm [B,H,W].

    `log_softmax_exp=True` is the Trainer's variant (src/models/trainer.py:1143-1144).
    """
    if log_softmax_exp:
        probs = F.log_softmax(mc_logits, dim=2).exp()
    else:
        probs = F.softmax(mc_logits, dim=2)                       # tester.py:412
    p_bar = probs.mean(dim=0)                                     # :415
    preds = p_bar.argmax(dim=1)                                   # :417
    C = p_bar.size(1)
    H_bar = -(p_bar.clamp_min(eps) * p_bar.clamp_min(eps).log()).sum(dim=1)   # :423 / :436
    H_norm = H_bar / math.log(C)                                  # :425
    H_t = -(probs.clamp_min(eps) * probs.clamp_min(eps).log()).sum(dim=2)     # :439
    EH = H_t.mean(dim=0)                                          # :442
    MI_norm = ((H_bar - EH) / math.log(C)).clamp_min(0.0)         # :446-449
    return {"p_bar": p_bar, "pred": preds, "H_norm": H_norm, "MI_norm": MI_norm}


@torch.no_grad()
def predictive_entropy_mc(mc_probs: torch.Tensor, eps: float = 1e-12, normalize: bool = True):
    """src/utils/mc_dropout.py:121-133."""
    mean_p = mc_probs.mean(dim=0).clamp_min(eps)
    ent = -(mean_p * torch.log(mean_p)).sum(dim=1)
    if not normalize:
        return ent
    C = mean_p.shape[1]
    return ent / float(torch.log(torch.tensor(C)).item())


def to_alpha_concentrations_from_shape_and_scale(shape_logits, scale_logits, T: float = 1.0, eps: float = EPS_DIR):
    """src/models/probability_helper.py:89-105: alpha = 1 + softplus(s/T) * softmax(z) + eps."""
    alpha_scale = F.softplus(scale_logits / T)
    alpha_shape = F.softmax(shape_logits, dim=1)
    return 1.0 + alpha_scale * alpha_shape + eps


def get_predictive_entropy(alpha, eps: float = EPS_DIR):
    """:116-121 (eps added inside the log)."""
    alpha0 = alpha.sum(dim=1, keepdim=True) + eps
    p_hat = alpha / alpha0
    return -(p_hat * torch.log(p_hat + eps)).sum(dim=1)


def get_aleatoric_uncertainty(alpha, eps: float = EPS_DIR):
    """:124-130."""
    alpha0 = alpha.sum(dim=1, keepdim=True) + eps
    term = digamma(alpha + 1.0) - digamma(alpha0 + 1.0)
    p_hat = alpha / alpha0
    return -(p_hat * term).sum(dim=1)


def get_epistemic_uncertainty(alpha, eps: float = EPS_DIR):
    """:133-136."""
    return get_predictive_entropy(alpha, eps) - get_aleatoric_uncertainty(alpha, eps)


def get_predictive_entropy_norm(alpha, eps: float = EPS_DIR):
    """:148-153."""
    return get_predictive_entropy(alpha, eps) / math.log(alpha.shape[1])


@torch.no_grad()
def dirichlet_mi(alpha, eps: float = 1e-12, normalize: bool = True):
    """Dirichlet mutual information as scored by src/metrics/auroc.py:55-63 (clamp-style eps)."""
    a0 = alpha.sum(dim=1, keepdim=True) + eps
    p = alpha / a0
    H = -(p.clamp_min(eps) * p.clamp_min(eps).log()).sum(dim=1)
    term = digamma(alpha + 1.0) - digamma(a0 + 1.0)
    EH = -(p * term).sum(dim=1)
    MI = H - EH
    return MI / math.log(alpha.size(1)) if normalize else MI


@torch.no_grad()
def evidential_reduce(outputs: torch.Tensor, num_classes: int):
    """Single-pass Dirichlet branch of src/models/tester.py:484-512 on [B,C+1,H,W] head output."""
    shape_logits = outputs[:, :num_classes]
    scale_logits = outputs[:, num_classes:num_classes + 1]
    probs = F.softmax(shape_logits, dim=1)                        # :493
    preds = probs.argmax(dim=1)                                   # :495
    alpha = to_alpha_concentrations_from_shape_and_scale(shape_logits, scale_logits)   # :500
    return {
        "alpha": alpha, "pred": preds,
        "H_norm": get_predictive_entropy_norm(alpha),             # :501
        "AU": get_aleatoric_uncertainty(alpha),
        "EU": get_epistemic_uncertainty(alpha),
        "MI_norm": dirichlet_mi(alpha),
    }
