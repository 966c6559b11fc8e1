"""`DirichletMSELoss` with the reference's interface (src/losses/dirichlet_losses.py:317-385), computed
by libslu's fused forward+backward kernel (csrc/slu_loss.cu).  `_valid_mask` mirrors :15-70.

`NLLDirichletCategorical` (:73-119), `DigammaDirichletCE` (:122-167) and `BrierDirichlet` (:174-220) -- weight
0 in every shipped config (src/configs/SemanticKitti_default.yaml:50-62) -- run on the single-term kernel
(slu_dirichlet_term), `ComplementKLUniform` (:228-314) on slu_evidence_term.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .. import ops
from ._function import _DirichletSingleTerm, _DirichletTerm, _EvidenceTerm, _ids_and_keep
from ._mask import _valid_mask  # noqa: F401  (re-exported, the reference Trainer imports it from here)


class DirichletMSELoss(nn.Module):
    """Expected squared error under a Dirichlet (Sensoy et al. 2018, eq. 5), masked mean over valid pixels."""

    def __init__(self, ignore_index: Optional[int] = None, eps: float = 1e-8):
        super().__init__()
        self.ignore_index = ignore_index
        self.eps = eps

    def forward(self, alpha: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if target.dim() == 4 and target.size(1) == 1:
            target = target[:, 0]
        target = target.long()
        if alpha.shape[1] <= 2:                       # the reference returns an exact zero here (:352-353)
            return alpha.sum() * 0.0
        return _DirichletTerm.apply(alpha, target, 0, self.ignore_index, self.eps)


def _prep(target: torch.Tensor) -> torch.Tensor:
    if target.dim() == 4 and target.size(1) == 1:
        target = target[:, 0]
    return target.long()


class NLLDirichletCategorical(nn.Module):
    """-log E[p_y] = log(alpha0 + eps) - log(alpha_y + eps), mean over valid pixels (:73-119)."""

    def __init__(self, ignore_index: Optional[int] = None, eps: float = 1e-12):
        super().__init__()
        self.ignore_index = ignore_index
        self.eps = eps

    def forward(self, alpha: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return _DirichletSingleTerm.apply(alpha, _prep(target), ops.TERM_NLL, self.ignore_index, self.eps, None)


class DigammaDirichletCE(nn.Module):
    """E[-log p_y] = psi(alpha0) - psi(alpha_y), mean over valid pixels (:122-167)."""

    def __init__(self, ignore_index: Optional[int] = None, eps: float = 1e-8):
        super().__init__()
        self.ignore_index = ignore_index
        self.eps = eps

    def forward(self, alpha: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return _DirichletSingleTerm.apply(alpha, _prep(target), ops.TERM_DIGAMMA_CE, self.ignore_index, self.eps, None)


class BrierDirichlet(nn.Module):
    """Expected Brier score under the Dirichlet predictive distribution (:174-220)."""

    def __init__(self, ignore_index: Optional[int] = None, s_ref: Optional[float] = None, eps: float = 1e-12):
        super().__init__()
        self.ignore_index = ignore_index
        self.s_ref = s_ref
        self.eps = eps

    def forward(self, alpha: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return _DirichletSingleTerm.apply(alpha, _prep(target), ops.TERM_BRIER, self.ignore_index, self.eps, self.s_ref)


class ComplementKLUniform(nn.Module):
    """KL(off-class conditional || uniform) gated by (1-p_y)^gamma * sigmoid((tau-p_y)/sigma) (:228-314)."""

    def __init__(self, ignore_index: Optional[int] = 0, gamma: float = 2.0, tau: float = 0.55, sigma: float = 0.12,
                 s_target: Optional[float] = None, normalize: bool = True, eps: float = 1e-8, detach_uncert: bool = True):
        super().__init__()
        self.ignore_index = ignore_index
        self.gamma = float(gamma)
        self.tau = float(tau)
        self.sigma = float(sigma)
        self.s_target = s_target
        self.normalize = bool(normalize)
        self.eps = eps
        self.detach_uncert = bool(detach_uncert)

    def forward(self, alpha: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        target = _prep(target)
        if alpha.shape[1] <= 2:                       # the reference returns an exact zero here (:275-276)
            return alpha.sum() * 0.0
        ids, keep = _ids_and_keep(target, self.ignore_index)
        prm = (self.gamma, self.tau, self.sigma, -1.0 if self.s_target is None else float(self.s_target),
               float(self.normalize), self.eps, float(self.detach_uncert))
        return _EvidenceTerm.apply(alpha, target, ops.TERM_COMP_KL, prm, ids, keep, 1.0)
