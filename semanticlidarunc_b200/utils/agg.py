"""Running mean attached to a function, with the reference's interface (src/utils/agg.py:6-91).

`@mean_aggregator()` turns `f` into a callable that still computes `f(...)` and additionally offers
`.add(x, mask=None)`, `.accumulate(*a, mask=None, **kw)`, `.mean(reset=False)`, `.reset()` and `.sync_ddp()`.
The running (sum, count) of tensor results lives on the tensors' device as one float64 pair, so `accumulate` never
synchronises; `.mean()` reads the pair back once.  Python scalars are summed on the host.
"""
from __future__ import annotations

import functools
from typing import Callable, Optional

import torch
import torch.distributed as tdist


class _RunningMean:
    """callable wrapper: behaves like the wrapped function, carries the (sum, count) state"""

    def __init__(self, fn: Callable):
        functools.update_wrapper(self, fn)
        self._fn = fn
        self.reset()

    def __call__(self, *args, **kwargs):
        return self._fn(*args, **kwargs)

    # ---- state
    def reset(self):
        self._pair: Optional[torch.Tensor] = None          # [sum, count] float64 on the results' device
        self._scalar_sum, self._scalar_n = 0.0, 0

    def _totals(self):
        total, n = self._scalar_sum, self._scalar_n
        if self._pair is not None:
            s, c = self._pair.cpu().tolist()
            total, n = total + s, n + int(c)
        return total, n

    # ---- accumulation
    def add(self, x, mask: Optional[torch.Tensor] = None):
        if not torch.is_tensor(x):
            self._scalar_sum += float(x)
            self._scalar_n += 1
            return
        x = x.detach().float()
        if mask is None:
            contrib = torch.stack([x.sum(dtype=torch.float64), torch.tensor(float(x.numel()), dtype=torch.float64, device=x.device)])
        else:
            if not torch.is_tensor(mask):
                raise TypeError("mask must be a torch.Tensor or None")
            keep = torch.broadcast_to(mask.to(x.device), x.shape)
            # the reference sums x[mask] (src/utils/agg.py:52): a NaN / Inf at a masked-out pixel must not reach the sum
            contrib = torch.stack([torch.where(keep.bool(), x, torch.zeros_like(x)).sum(dtype=torch.float64),
                                   keep.sum(dtype=torch.float64)])
        if self._pair is None:
            self._pair = contrib
        elif self._pair.device == contrib.device:
            self._pair += contrib
        else:                                              # results moved to another device: carry the state along
            self._pair = contrib + self._pair.to(contrib.device)

    def accumulate(self, *args, mask: Optional[torch.Tensor] = None, **kwargs):
        value = self._fn(*args, **kwargs)
        self.add(value, mask=mask)
        return value

    # ---- read-out
    def mean(self, reset: bool = False) -> float:
        total, n = self._totals()
        if reset:
            self.reset()
        return total / max(1, n)

    def sync_ddp(self):
        """Sum (sum, count) over the default process group (the reference's one collective, :75-83)."""
        if not (tdist.is_available() and tdist.is_initialized()):
            return
        total, n = self._totals()
        use_cuda = torch.cuda.is_available() and tdist.get_backend() != "gloo"
        t = torch.tensor([total, float(n)], dtype=torch.float64, device=torch.device("cuda", torch.cuda.current_device()) if use_cuda else "cpu")
        tdist.all_reduce(t, op=tdist.ReduceOp.SUM)
        self.reset()
        self._scalar_sum, self._scalar_n = float(t[0]), int(t[1])


def mean_aggregator():
    return _RunningMean
