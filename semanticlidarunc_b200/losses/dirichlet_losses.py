"""`DirichletMSELoss` with the reference's interface (src/losses/dirichlet_losses.py:317-385), computed
by libslu's fused forward+backward kernel (csrc/slu_loss.cu).  `_valid_mask` mirrors :15-70.

The other Dirichlet data-fit terms of that file (NLLDirichletCategorical :73, DigammaDirichletCE :122,
BrierDirichlet :174, ComplementKLUniform :228) have weight 0 in every shipped config
(src/configs/SemanticKitti_default.yaml:50-62) and are next-round work.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from ._function import _DirichletTerm
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
