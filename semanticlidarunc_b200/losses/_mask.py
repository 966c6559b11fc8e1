"""`ignore_index` forms accepted by the reference's `_valid_mask` (src/losses/dirichlet_losses.py:15-70):
None, int, iterable of ints, integer tensor of ids, or a bool tensor shaped like target (True = keep)."""
from __future__ import annotations

from collections.abc import Iterable

import torch


def split_ignore(ignore_index):
    """-> (tuple of ignored ids, keep-mask tensor or None)."""
    if ignore_index is None:
        return (), None
    if torch.is_tensor(ignore_index):
        if ignore_index.dtype == torch.bool:
            return (), ignore_index
        return tuple(int(v) for v in ignore_index.reshape(-1).tolist()), None
    if isinstance(ignore_index, int):
        return (ignore_index,), None
    if isinstance(ignore_index, Iterable):
        return tuple(int(v) for v in ignore_index), None
    raise TypeError("ignore_index must be None, int, Iterable[int], or bool Tensor")


def _valid_mask(target: torch.Tensor, ignore_index) -> torch.Tensor:
    """Boolean mask of valid pixels (same contract as the reference helper); torch ops, any device."""
    if target.dim() == 4 and target.size(1) == 1:
        target = target[:, 0]
    ids, keep = split_ignore(ignore_index)
    if keep is not None:
        return keep
    if not ids:
        return torch.ones_like(target, dtype=torch.bool)
    return ~torch.isin(target, torch.as_tensor(ids, device=target.device, dtype=target.dtype))
