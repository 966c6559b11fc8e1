"""Multi-GPU plumbing: scans shard by index, only the integer evaluation counts are combined.

The reference is single-process (its one dormant collective is an all_reduce of [sum, count],
src/utils/agg.py:75-83).  Here every rank owns `scans[rank::world]`, runs stages 1-4 on them with no
data-path collective, and the confusion matrix + reliability bins (C*C + 3*n_bins int64, ~3.6 KB)
are summed with ONE all-reduce at the end of a sweep.  Integer sums are order independent, so the
N-GPU counts are bit-identical to the 1-GPU counts.  Backend: NCCL for CUDA tensors (NVLink /
NVSwitch: pure latency at this size), gloo for the CPU tests.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist


def world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_indices(n_items: int, rank: Optional[int] = None, world_size: Optional[int] = None) -> range:
    """Indices of the scans rank `rank` owns: i = rank, rank+world, ...."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside [0,{world_size})")
    return range(rank, n_items, world_size)


def pack_counts(confmat: torch.Tensor, ece_bins: Optional[torch.Tensor] = None) -> torch.Tensor:
    parts = [confmat.reshape(-1)]
    if ece_bins is not None:
        parts.append(ece_bins.reshape(-1))
    return torch.cat(parts).to(torch.int64)


_PEERS = {}                      # (id of the group, device index) -> PeerCounter, or None once the set-up has failed there


def peer_counter(group=None, device=None):
    """The NVLink mailboxes of (group, device): created on first use (a collective: every rank of the group must get
    here, as it must for the all-reduce this replaces), None when the ranks do not share a node, the driver refuses the
    IPC mapping, or SLU_NO_PEER_EXCHANGE=1 -- the callers then use the process group's all-reduce."""
    if os.environ.get("SLU_NO_PEER_EXCHANGE", "0") == "1" or not torch.cuda.is_available():
        return None
    if not (dist.is_available() and dist.is_initialized()):
        return None
    g = None if group is True else group
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    key = (id(g) if g is not None else 0, dev.index)
    if key not in _PEERS:
        _PEERS[key] = PeerCounter.create(g, device=dev)
    return _PEERS[key]


def count_transport(group=None, device=None) -> str:
    """"peer-memory", "nccl" (the process group's all-reduce) or "none" (one rank): how the small exchanges travel"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(None if group is True else group) == 1:
        return "none"
    return "peer-memory" if peer_counter(group, device) is not None else "nccl"


def allreduce_packed(buf: torch.Tensor, group=None) -> torch.Tensor:
    """Sum a packed int64 counter buffer over all ranks, in place (no-op without a process group).  CUDA buffers of up to
    512 elements go through one single-CTA kernel over NVLink peer memory (ops.peer_allreduce_i64) when the ranks share a
    node; everything else through the process group (NCCL / gloo)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        if buf.is_cuda and buf.dtype == torch.int64 and buf.is_contiguous() and buf.numel() <= 512:
            peers = peer_counter(group, buf.device)
            if peers is not None:
                from . import ops
                buf.copy_(ops.peer_allreduce_i64(buf, None, peers))
                return buf
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf


def reduced_counts(confmat: torch.Tensor, ece_bins: Optional[torch.Tensor] = None, group=None):
    """(confmat, ece_bins) summed over all ranks with a single all-reduce of a packed COPY.  The arguments are
    left untouched, so live accumulators can keep accumulating and be reduced again later."""
    n = confmat.numel()
    if (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1 and confmat.is_cuda and
            confmat.dtype == torch.int64 and confmat.is_contiguous() and (ece_bins is None or (ece_bins.dtype == torch.int64 and ece_bins.is_contiguous())) and
            n + (0 if ece_bins is None else ece_bins.numel()) <= 512):
        peers = peer_counter(group, confmat.device)
        if peers is not None:                        # one kernel reads both accumulators in place and writes the packed sums
            from . import ops
            buf = ops.peer_allreduce_i64(confmat, ece_bins, peers)
            return buf[:n].view_as(confmat), (None if ece_bins is None else buf[n:].view_as(ece_bins))
    buf = allreduce_packed(pack_counts(confmat, ece_bins), group=group)
    return buf[:n].view_as(confmat), (None if ece_bins is None else buf[n:].view_as(ece_bins))


def allreduce_counts(confmat: torch.Tensor, ece_bins: Optional[torch.Tensor] = None, group=None) -> None:
    """Sum the evaluation counters over all ranks IN PLACE with a single all-reduce.

    The arguments afterwards hold GLOBAL sums: do not keep accumulating into them and do not reduce them a second
    time (every count would be multiplied by the world size) -- end-of-sweep use only.  For live accumulators use
    `reduced_counts`, which reduces a copy."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    buf = allreduce_packed(pack_counts(confmat, ece_bins), group=group)
    n = confmat.numel()
    confmat.copy_(buf[:n].view_as(confmat))
    if ece_bins is not None:
        ece_bins.copy_(buf[n:].view_as(ece_bins))


def bind_to_gpu_numa_node(index: int):
    """Pin this rank to the CPUs NVML reports as local to its GPU, BEFORE any pinned host memory is
    allocated: first-touch then places the staging buffers on the GPU's own NUMA node, so 8 ranks do not
    pull 3.4 GB per step each through one socket's memory controllers and the inter-socket link."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1 and 64 * w + b < n_cpu]
        allowed = set(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


class PeerCounter:
    """Mailboxes for `ops.count_valid_exchange`: the valid-pixel count of the batch-sharded training step summed over the
    ranks by the count kernel itself, through NVLink / NVSwitch peer stores instead of an NCCL all-reduce (tens of
    microseconds of collective latency between two kernels that take a few microseconds each).

    Set-up is plumbing: every rank allocates a 4 KB mailbox in libslu (cudaMalloc), the 64-byte CUDA IPC handles go round
    with ONE all_gather over the process group, and every rank maps the others' mailboxes.  Works for ranks on one node
    (one process per GPU); `PeerCounter.create` returns None when that is not the case or the driver refuses the
    mapping, and the caller keeps the NCCL path."""

    def __init__(self, boxes, rank, world, timeout_s=None):
        import ctypes as C
        self._boxes = list(boxes)                       # device pointers (ints); boxes[rank] is the own mailbox
        # how long a kernel waits for its peers before it poisons its result (NaN / INT64_MIN) instead of hanging: long enough for
        # ordinary skew between ranks (a slow data loader, a first-iteration compile), short enough to surface a dead rank
        if timeout_s is None:
            timeout_s = float(os.environ.get("SLU_PEER_TIMEOUT_S", "30"))
        self.rank, self.world, self.timeout_s = int(rank), int(world), float(timeout_s)
        self.boxes_array = (C.c_void_p * self.world)(*[C.c_void_p(b) for b in self._boxes])
        self._closed = False

    @classmethod
    def create(cls, group=None, device=None, timeout_s=None):
        import ctypes as C
        import socket
        from . import _lib
        if not (dist.is_available() and dist.is_initialized()):
            return None
        g = None if group is True else group
        world_size, rank = dist.get_world_size(g), dist.get_rank(g)
        if world_size < 2 or world_size > 16 or not torch.cuda.is_available():
            return None
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        own, handle, err = C.c_void_p(), (C.c_uint8 * 64)(), ""
        with torch.cuda.device(dev):
            try:
                _lib.check(_lib.lib().slu_peer_mailbox_create(C.byref(own), handle), "slu_peer_mailbox_create")
            except Exception as e:                      # no IPC support here: every rank must still take part below
                err = str(e)
            mine = {"host": socket.gethostname(), "handle": bytes(handle), "err": err}
            everyone = [None] * world_size
            dist.all_gather_object(everyone, mine, group=g)
            ok = all(not e["err"] for e in everyone) and len({e["host"] for e in everyone}) == 1
            boxes, opened = [None] * world_size, []
            if ok:
                for r, e in enumerate(everyone):
                    if r == rank:
                        boxes[r] = own.value
                        continue
                    p = C.c_void_p()
                    buf = (C.c_uint8 * 64).from_buffer_copy(e["handle"])
                    if _lib.lib().slu_peer_mailbox_open(buf, C.byref(p)) != 0:
                        ok = False
                        break
                    boxes[r] = p.value
                    opened.append(p.value)
            # all or nothing: a rank that could not map a peer makes every rank fall back
            flags = [None] * world_size
            dist.all_gather_object(flags, bool(ok), group=g)
            if not all(flags):
                for p in opened:
                    _lib.lib().slu_peer_mailbox_close(C.c_void_p(p))
                if own.value:
                    _lib.lib().slu_peer_mailbox_destroy(own)
                return None
        return cls(boxes, rank, world_size, timeout_s)

    def timeouts(self) -> int:
        """exchanges of THIS rank that gave up waiting for a peer (synchronises the device)"""
        import ctypes as C
        from . import _lib
        n = C.c_uint32(0)
        _lib.check(_lib.lib().slu_peer_mailbox_timeouts(C.c_void_p(self._boxes[self.rank]), C.byref(n)), "slu_peer_mailbox_timeouts")
        return int(n.value)

    def close(self, group=None):
        """Unmap the peers' mailboxes, then (after a barrier: nobody may still be mapped) free the own one."""
        import ctypes as C
        from . import _lib
        if self._closed:
            return
        self._closed = True
        torch.cuda.synchronize()
        for r, b in enumerate(self._boxes):
            if r != self.rank and b:
                _lib.lib().slu_peer_mailbox_close(C.c_void_p(b))
        if dist.is_available() and dist.is_initialized():
            dist.barrier(group=None if group is True else group)
        _lib.lib().slu_peer_mailbox_destroy(C.c_void_p(self._boxes[self.rank]))
