"""Regularisers with the reference's interfaces (src/losses/regularizers.py), each one fused forward+backward
kernel launch of libslu:

`KL_offClasses_to_uniform` :291-389 (csrc/slu_loss.cu; `with_conf_weighting=True` -> slu_evidence_term),
`LogitRegularizer` :75-110 (slu_logit_regularizer), `EvidenceRegBand` :116-147, `EvidenceReg` :149-212,
`WrongLowEvidence` :218-289 (csrc/slu_loss_terms.cu).  `_valid_mask` :9-53 and `_mean_over_valid` :56-69
semantics are kept: `mask=` wins over `target=`+`ignore_index`, and without either the mean is over everything.
"""
from __future__ import annotations

from typing import Iterable, Optional, Union

import torch
import torch.nn as nn

from .. import ops
from ._function import _DirichletTerm, _EvidenceTerm, _LogitReg, _ids_and_keep
from ._mask import _valid_mask  # noqa: F401

_Ignore = Optional[Union[int, Iterable[int], torch.Tensor]]


def _prep(target):
    if target is None:
        return None
    if target.dim() == 4 and target.size(1) == 1:
        target = target[:, 0]
    return target.long()


class KL_offClasses_to_uniform(nn.Module):
    """KL( Dir(alpha with the true class' evidence removed) || Dir(1,...,1) ), mean over valid pixels; with
    `with_conf_weighting` each pixel is weighted by the detached (1 - p_y)^gamma and the sum divided by
    max(sum of weights, 1)."""

    def __init__(self, ignore_index: Optional[int] = None, with_conf_weighting: bool = False, gamma: float = 1.0,
                 eps: float = 1e-8):
        super().__init__()
        self.ignore_index = ignore_index
        self.eps = eps
        self.with_conf_weighting = with_conf_weighting
        self.gamma = gamma

    def forward(self, alpha: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        target = _prep(target)
        if self.with_conf_weighting:
            ids, keep = _ids_and_keep(target, self.ignore_index)
            return _EvidenceTerm.apply(alpha, target, ops.TERM_KL_CONF, (float(self.gamma), float(self.eps)), ids, keep, 1.0)
        return _DirichletTerm.apply(alpha, target, 1, self.ignore_index, self.eps)


class LogitRegularizer(nn.Module):
    """mean(z^2), or mean(relu(z - threshold)^2), over valid elements (:75-110)."""

    def __init__(self, threshold: Optional[float] = None, ignore_index: _Ignore = None):
        super().__init__()
        self.threshold = threshold
        self.ignore_index = ignore_index

    def forward(self, logits: torch.Tensor, *, mask: Optional[torch.Tensor] = None,
                target: Optional[torch.Tensor] = None) -> torch.Tensor:
        target = _prep(target) if mask is None else None
        ids, keep = _ids_and_keep(target, self.ignore_index, mask)
        if keep is None and target is not None and not ids:
            target = None                             # ignore_index=None: every pixel valid, but the masked mean applies
            keep = torch.ones(logits.shape[0:1] + logits.shape[2:], dtype=torch.bool, device=logits.device)
        return _LogitReg.apply(logits, target, None if self.threshold is None else float(self.threshold), ids, keep)


class _A0Term(nn.Module):
    """shared forward of the two total-evidence regularisers"""
    _term: int

    def _params(self):
        raise NotImplementedError

    def forward(self, alpha: torch.Tensor, *, mask: Optional[torch.Tensor] = None,
                target: Optional[torch.Tensor] = None) -> torch.Tensor:
        target = _prep(target) if mask is None else None
        ids, keep = _ids_and_keep(target, self.ignore_index, mask)
        if not ids:
            target = None                             # nothing to look up in the target
        return _EvidenceTerm.apply(alpha, target, self._term, self._params(), ids, keep, 1e-8)


class EvidenceRegBand(_A0Term):
    """Two-sided squared log hinge on a0 = sum(alpha) outside [s(1-band), s(1+band)] (:116-147)."""
    _term = ops.TERM_EVID_BAND

    def __init__(self, s_target: float, band: float = 0.10, ignore_index: _Ignore = None):
        super().__init__()
        self.s = float(s_target)
        self.band = float(band)
        self.ignore_index = ignore_index

    def _params(self):
        return (self.s, self.band)


class EvidenceReg(_A0Term):
    """log(a0/s)^2 | relu(a0 - s(1+margin))^2 | (a0 - s)^2, mean over valid pixels (:149-212)."""
    _term = ops.TERM_EVID_REG
    _MODES = {"log_squared": 0, "one_sided": 1}

    def __init__(self, s_target: float, mode: str = "log_squared", margin: float = 0.1, scale_correct: bool = False,
                 ignore_index: _Ignore = None):
        super().__init__()
        self.s_target = float(s_target)
        self.mode = mode
        self.margin = float(margin)
        self.scale_correct = bool(scale_correct)
        self.ignore_index = ignore_index

    def _params(self):
        return (self.s_target, float(self._MODES.get(self.mode, 2)), self.margin, float(self.scale_correct))


class WrongLowEvidence(nn.Module):
    """Squared hinge of log(a0) above log(C + s_low) on wrongly predicted pixels, gated by the confidence margin
    (:218-289); averaged over the sum of the gates."""

    def __init__(self, ignore_index=None, s_low: float = 0.0, margin: float = 0.05, soft_margin_k: float = 0.08,
                 eps: float = 1e-8):
        super().__init__()
        self.ignore_index = ignore_index
        self.s_low = float(s_low)
        self.margin = float(margin)
        self.k = float(soft_margin_k)
        self.eps = float(eps)

    def forward(self, alpha: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        target = _prep(target)
        ids, keep = _ids_and_keep(target, self.ignore_index)
        return _EvidenceTerm.apply(alpha, target, ops.TERM_WRONG_LOW, (self.s_low, self.margin, self.k, self.eps), ids, keep, 1.0)
