"""Fused evidential training loss: the trainer's chain `outputs -> alpha -> w_mse * DirichletMSELoss +
w_kl * KL_offClasses_to_uniform` (src/models/trainer.py:532-578 with the shipped weights of
src/configs/SemanticKitti_default.yaml:50-62) as ONE forward+backward kernel pass over the head output.

Use this when the term weights are fixed; when GradNorm needs `autograd.grad` per term
(src/utils/grad_norm.py:52) use the per-term modules in dirichlet_losses.py / regularizers.py.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ._mask import split_ignore


def _count_reducer(group):
    """all-reduce of the valid-pixel count over `group` (True = the default group), or None when not sharded"""
    import torch.distributed as dist
    if group is None or group is False or not (dist.is_available() and dist.is_initialized()):
        return None
    g = None if group is True else group
    if dist.get_world_size(g) == 1:
        return None
    return lambda count: dist.all_reduce(count, op=dist.ReduceOp.SUM, group=g)


class _FusedEvidentialLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, outputs, target, w_mse, w_kl, ignore_index, temperature, eps, group):
        ids, keep = split_ignore(ignore_index)
        if len(ids) > 8:
            keep = ~torch.isin(target, torch.as_tensor(ids, device=target.device, dtype=target.dtype))
            ids = ()
        r = ops.evidential_loss_fused(outputs.detach(), target, w_mse=w_mse, w_kl=w_kl, ignore=ids, keep_mask=keep,
                                      temperature=temperature, eps_alpha=eps, eps_mse=eps, eps_kl=eps,
                                      want_grad=ctx.needs_input_grad[0], count_reduce=_count_reducer(group))
        n = r["sums"][2].clamp_min(1.0)
        mse, kl = r["sums"][0] / n, r["sums"][1] / n
        if ctx.needs_input_grad[0]:
            ctx.save_for_backward(r["grad"])
        total = (w_mse * mse + w_kl * kl).to(outputs.dtype)
        mse, kl = mse.to(outputs.dtype), kl.to(outputs.dtype)
        ctx.mark_non_differentiable(mse, kl)
        return total, mse, kl

    @staticmethod
    def backward(ctx, g_total, _g_mse, _g_kl):
        (grad,) = ctx.saved_tensors
        return grad * g_total.to(grad.dtype), None, None, None, None, None, None, None


class EvidentialLoss(nn.Module):
    """forward(outputs [B,C+1,H,W], target [B,H,W] | [B,1,H,W]) -> (loss, mse.detach(), kl.detach()).

    Batch-sharded training (BASELINE.json configs[4]): with `group=True` (or a process group) every rank passes its
    shard of the batch; the valid-pixel counts are all-reduced (one float64 over NCCL) between the count kernel and the
    loss kernel, so each rank's `loss`, `mse`, `kl` are its SHARE of the global masked mean (their sum over ranks is the
    single-process value) and the gradient it writes is exactly the single-process gradient of its shard.  No other
    collective is needed; summing the backbone's parameter gradients stays the DDP wrapper's job (use sum, not mean)."""

    def __init__(self, w_mse: float = 1.0, w_kl: float = 0.05, ignore_index=None, temperature: float = 1.0, eps: float = 1e-8,
                 group=None):
        super().__init__()
        self.w_mse, self.w_kl = float(w_mse), float(w_kl)
        self.ignore_index, self.temperature, self.eps = ignore_index, float(temperature), float(eps)
        self.group = group

    def forward(self, outputs: torch.Tensor, target: torch.Tensor):
        if target.dim() == 4 and target.size(1) == 1:
            target = target[:, 0]
        return _FusedEvidentialLoss.apply(outputs, target.long(), self.w_mse, self.w_kl, self.ignore_index,
                                          self.temperature, self.eps, self.group)
