"""`KL_offClasses_to_uniform` with the reference's interface (src/losses/regularizers.py:291-389),
computed by libslu's fused forward+backward kernel (csrc/slu_loss.cu).

`with_conf_weighting=True` (a detached per-pixel weight (1 - p_y)^gamma, :369-383) is not used by any
shipped config and is not implemented on the device path yet.  The remaining regularisers of that file
(LogitRegularizer :75, EvidenceRegBand :116, EvidenceReg :149, WrongLowEvidence :218) have weight 0 in
the shipped configs and are next-round work.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from ._function import _DirichletTerm


class KL_offClasses_to_uniform(nn.Module):
    """KL( Dir(alpha with the true class' evidence removed) || Dir(1,...,1) ), mean over valid pixels."""

    def __init__(self, ignore_index: Optional[int] = None, with_conf_weighting: bool = False, gamma: float = 1.0,
                 eps: float = 1e-8):
        super().__init__()
        if with_conf_weighting:
            raise NotImplementedError("with_conf_weighting=True is not implemented by the CUDA loss kernel")
        self.ignore_index = ignore_index
        self.eps = eps
        self.with_conf_weighting = with_conf_weighting
        self.gamma = gamma

    def forward(self, alpha: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if target.dim() == 4 and target.size(1) == 1:
            target = target[:, 0]
        return _DirichletTerm.apply(alpha, target.long(), 1, self.ignore_index, self.eps)
