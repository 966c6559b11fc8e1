"""Running-mean decorator with the reference's interface (src/utils/agg.py:6-91).

`@mean_aggregator()` attaches `.add(x, mask=None)`, `.accumulate(*a, mask=None, **kw)`,
`.mean(reset=False)`, `.reset()` and `.sync_ddp()` to a function.  Sum and count stay on the device as
one float64 pair, so `accumulate` does not synchronise; `.mean()` reads it back once.
"""
from __future__ import annotations

from functools import wraps
from typing import Optional

import torch


def mean_aggregator():
    def decorator(fn):
        state = {"acc": None, "host_sum": 0.0, "host_count": 0}

        @wraps(fn)
        def wrapped(*args, **kwargs):
            return fn(*args, **kwargs)

        def add(x, mask: Optional[torch.Tensor] = None):
            if not torch.is_tensor(x):
                state["host_sum"] += float(x)
                state["host_count"] += 1
                return
            x = x.detach()
            if mask is not None:
                if not torch.is_tensor(mask):
                    raise TypeError("mask must be a torch.Tensor or None")
                m = torch.broadcast_to(mask, x.shape).to(x.device)
                s, c = torch.where(m, x.float(), torch.zeros((), device=x.device)).sum(dtype=torch.float64), m.sum()
            else:
                s, c = x.float().sum(dtype=torch.float64), torch.tensor(x.numel(), device=x.device)
            pair = torch.stack([s, c.to(torch.float64)])
            if state["acc"] is None or state["acc"].device != pair.device:
                if state["acc"] is not None:
                    pair = pair + state["acc"].to(pair.device)
                state["acc"] = pair
            else:
                state["acc"] += pair

        def accumulate(*args, mask: Optional[torch.Tensor] = None, **kwargs):
            out = fn(*args, **kwargs)
            add(out, mask=mask)
            return out

        def _totals():
            s, c = state["host_sum"], state["host_count"]
            if state["acc"] is not None:
                a = state["acc"].cpu()
                s, c = s + float(a[0]), c + int(a[1])
            return s, c

        def mean(reset: bool = False) -> float:
            s, c = _totals()
            if reset:
                reset_()
            return s / max(1, c)

        def reset_():
            state["acc"], state["host_sum"], state["host_count"] = None, 0.0, 0

        def sync_ddp():
            if not torch.distributed.is_available() or not torch.distributed.is_initialized():
                return
            s, c = _totals()
            dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
            t = torch.tensor([s, float(c)], dtype=torch.float64, device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)
            reset_()
            state["host_sum"], state["host_count"] = float(t[0]), int(t[1])

        wrapped.add, wrapped.accumulate, wrapped.mean, wrapped.reset, wrapped.sync_ddp = add, accumulate, mean, reset_, sync_ddp
        return wrapped
    return decorator
