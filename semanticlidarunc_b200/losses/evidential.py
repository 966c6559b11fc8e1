"""Fused evidential training loss: the trainer's chain `outputs -> alpha -> w_mse * DirichletMSELoss +
w_kl * KL_offClasses_to_uniform` (src/models/trainer.py:532-578 with the shipped weights of
src/configs/SemanticKitti_default.yaml:50-62) as ONE forward+backward kernel pass over the head output.

Use this when the term weights are fixed; when GradNorm needs `autograd.grad` per term
(src/utils/grad_norm.py:52) use the per-term modules in dirichlet_losses.py / regularizers.py.

Three ways in, same kernel (slu_evidential_loss_step; the loss values are written by its last CTA, no host arithmetic):

  crit(outputs, target)                       autograd: `loss.backward()` flows into the backbone (one extra pass for
                                              autograd's upstream-gradient multiply)
  crit.forward_backward(outputs, target)      no autograd node: returns (loss4, grad); feed the backbone with
                                              `outputs.backward(grad)` -- the gradient is final, nothing re-reads it
  crit.capture(outputs, target); crit.replay()  the same as ONE CUDA graph over static buffers (count kernel, the NCCL
                                              all-reduce of the count when sharded, loss kernel)

Batch-sharded training (BASELINE.json configs[4], `group=`): the only thing ranks exchange is the valid-pixel count.
`prefetch_count(target)` launches the count kernel and its all-reduce on a SIDE stream as soon as the labels exist
(they do not depend on the network), so the collective overlaps whatever the main stream does next (loader kernels,
the backbone's forward) instead of sitting between the count and the loss kernel.  With `peer_exchange=True` (default)
and all ranks on one node, count and all-reduce are ONE kernel: its last CTA publishes the local count in every rank's
mailbox over NVLink peer stores and sums what the others published (`dist.PeerCounter`, csrc/slu_peer.cu); NCCL stays
the fallback (other topologies, `peer_exchange=False`).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ._mask import split_ignore


def _group_of(group):
    """(process group or None for the default group, world size); (None, 1) when not sharded"""
    import torch.distributed as dist
    if group is None or group is False or not (dist.is_available() and dist.is_initialized()):
        return None, 1
    g = None if group is True else group
    return g, dist.get_world_size(g)


def _count_reducer(group):
    """all-reduce of the valid-pixel count over `group` (True = the default group) as a callable, or None when not
    sharded -- the one collective of the batch-sharded loss (also what ops.evidential_loss_fused(count_reduce=) takes)"""
    import torch.distributed as dist
    g, world = _group_of(group)
    if world == 1:
        return None
    return lambda count: dist.all_reduce(count, op=dist.ReduceOp.SUM, group=g)


class _FusedEvidentialLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, outputs, target, crit):
        r = crit._step(outputs.detach(), target, want_grad=ctx.needs_input_grad[0])
        if ctx.needs_input_grad[0]:
            ctx.save_for_backward(r["grad"])
        l4 = r["loss4"]
        total, mse, kl = l4[0].to(outputs.dtype), l4[1].to(outputs.dtype), l4[2].to(outputs.dtype)
        ctx.mark_non_differentiable(mse, kl)
        return total, mse, kl

    @staticmethod
    def backward(ctx, g_total, _g_mse, _g_kl):
        (grad,) = ctx.saved_tensors
        return grad * g_total.to(grad.dtype), None, None


class EvidentialLoss(nn.Module):
    """forward(outputs [B,C+1,H,W], target [B,H,W] | [B,1,H,W]) -> (loss, mse.detach(), kl.detach()).

    Batch-sharded training: with `group=True` (or a process group) every rank passes its shard of the batch; the
    valid-pixel counts are all-reduced (one float64 over NCCL), so each rank's `loss`, `mse`, `kl` are its SHARE of the
    global masked mean (their sum over ranks is the single-process value) and the gradient it writes is bit for bit the
    single-process gradient of its shard.  No other collective is needed; summing the backbone's parameter gradients
    stays the DDP wrapper's job (use sum, not mean)."""

    def __init__(self, w_mse: float = 1.0, w_kl: float = 0.05, ignore_index=None, temperature: float = 1.0, eps: float = 1e-8,
                 group=None, peer_exchange: bool = True):
        super().__init__()
        self.w_mse, self.w_kl = float(w_mse), float(w_kl)
        self.ignore_index, self.temperature, self.eps = ignore_index, float(temperature), float(eps)
        self.group = group
        self.peer_exchange = bool(peer_exchange)
        self._peers = {}             # per device: dist.PeerCounter, or None once the set-up has failed (NCCL from then on)
        self._bufs = {}              # per device: count float64[1], state float64[3]
        self._side = {}              # per device: side stream of the count prefetch
        self._prefetched = None      # (event, count buffer) of a pending prefetch_count()
        self._graph = None

    # ---- plumbing ------------------------------------------------------------------------------------------------
    def _work_buffers(self, dev):
        if dev not in self._bufs:
            self._bufs[dev] = (torch.zeros(1, dtype=torch.float64, device=dev), torch.zeros(3, dtype=torch.float64, device=dev))
        return self._bufs[dev]

    def _mask(self, target):
        ids, keep = split_ignore(self.ignore_index)
        if len(ids) > 8:
            keep = ~torch.isin(target, torch.as_tensor(ids, device=target.device, dtype=target.dtype))
            ids = ()
        return ids, keep

    @staticmethod
    def _target3(target):
        if target.dim() == 4 and target.size(1) == 1:
            target = target[:, 0]
        return target if target.dtype == torch.int64 else target.long()

    def _peer_counter(self, dev):
        """the NVLink mailboxes of this device (collective set-up on first use), or None -> NCCL"""
        if not self.peer_exchange:
            return None
        if dev not in self._peers:
            from ..dist import peer_counter
            self._peers[dev] = peer_counter(self.group, dev)        # one set of mailboxes per (group, device), shared with dist.*
        return self._peers[dev]

    def count_transport(self, dev=None) -> str:
        """"peer-memory" or "nccl": how the valid-pixel count is summed over the ranks on `dev` ("none": not sharded)"""
        if _group_of(self.group)[1] == 1:
            return "none"
        dev = torch.device("cuda", torch.cuda.current_device()) if dev is None else torch.device(dev)
        return "peer-memory" if self._peer_counter(dev) is not None else "nccl"

    def _global_count(self, target, count, ids, keep, g):
        """count <- valid pixels of `target` over all ranks, enqueued on the current stream"""
        peers = self._peer_counter(target.device)
        if peers is not None:
            ops.count_valid_exchange(target, count, peers, ignore=ids, keep_mask=keep)     # one kernel: count + exchange
            return
        import torch.distributed as dist
        count.zero_()
        ops.count_valid(target, count, ignore=ids, keep_mask=keep)
        dist.all_reduce(count, op=dist.ReduceOp.SUM, group=g)

    # ---- the sharded count, off the critical path ------------------------------------------------------------------
    @torch.no_grad()
    def prefetch_count(self, target: torch.Tensor):
        """Count the valid pixels of `target` and all-reduce the count over the group on a side stream, now.  The next
        forward / forward_backward on the same target waits for it with a stream event instead of running the count
        and the collective in line.  No-op without a group."""
        g, world = _group_of(self.group)
        if world == 1:
            return
        target = self._target3(target)
        dev = target.device
        count, _ = self._work_buffers(dev)
        ids, keep = self._mask(target)
        self._peer_counter(dev)                      # collective set-up happens here, outside any stream capture
        main = torch.cuda.current_stream(dev)
        if dev not in self._side:
            self._side[dev] = torch.cuda.Stream(device=dev)
        side = self._side[dev]
        side.wait_stream(main)                       # the labels are produced on the main stream
        with torch.cuda.stream(side):
            self._global_count(target, count, ids, keep, g)
            ev = torch.cuda.Event()
            ev.record(side)
        target.record_stream(side)
        self._prefetched = (ev, dev)

    def _step(self, outputs, target, want_grad=True, loss4=None, grad=None):
        target = self._target3(target)
        dev = outputs.device
        count, state = self._work_buffers(dev)
        ids, keep = self._mask(target)
        g, world = _group_of(self.group)
        precounted = world > 1
        if precounted:
            if self._prefetched is not None and self._prefetched[1] == dev:
                torch.cuda.current_stream(dev).wait_event(self._prefetched[0])
                self._prefetched = None
            else:                                    # in line: count kernel (+ exchange) -> loss kernel
                self._global_count(target, count, ids, keep, g)
        return ops.evidential_loss_step(outputs, target, count, state, w_mse=self.w_mse, w_kl=self.w_kl, ignore=ids,
                                        keep_mask=keep, temperature=self.temperature, eps_alpha=self.eps, eps_mse=self.eps,
                                        eps_kl=self.eps, precounted=precounted, want_grad=want_grad, loss4=loss4, grad=grad)

    # ---- the three entry points ----------------------------------------------------------------------------------
    def forward(self, outputs: torch.Tensor, target: torch.Tensor):
        return _FusedEvidentialLoss.apply(outputs, self._target3(target), self)

    @torch.no_grad()
    def forward_backward(self, outputs: torch.Tensor, target: torch.Tensor, loss4=None, grad=None):
        """(loss4, grad) without an autograd node: loss4 float32[4] = loss | mse | kl | n_valid, grad = d(loss)/d(outputs)
        (final: already divided by the global count).  Continue into the backbone with `outputs.backward(grad)`."""
        r = self._step(outputs.detach(), target, want_grad=True, loss4=loss4, grad=grad)
        return r["loss4"], r["grad"]

    def capture(self, outputs: torch.Tensor, target: torch.Tensor):
        """Record forward_backward for THESE tensors into a CUDA graph (count kernel, the count all-reduce when sharded,
        loss kernel).  `outputs` / `target` become the graph's static inputs; returns the static (loss4, grad)."""
        dev = outputs.device
        with torch.cuda.device(dev):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    self.forward_backward(outputs, target)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self.forward_backward(outputs, target)
            self._graph = (g, out)
        return out

    def replay(self):
        if self._graph is None:
            raise RuntimeError("capture() first")
        self._graph[0].replay()
        return self._graph[1]
