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


def allreduce_packed(buf: torch.Tensor, group=None) -> torch.Tensor:
    """Sum a packed int64 counter buffer over all ranks, in place (no-op without a process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf


def reduced_counts(confmat: torch.Tensor, ece_bins: Optional[torch.Tensor] = None, group=None):
    """(confmat, ece_bins) summed over all ranks with a single all-reduce of a packed COPY.  The arguments are
    left untouched, so live accumulators can keep accumulating and be reduced again later."""
    buf = allreduce_packed(pack_counts(confmat, ece_bins), group=group)
    n = confmat.numel()
    return buf[:n].view_as(confmat), (None if ece_bins is None else buf[n:].view_as(ece_bins))


def allreduce_counts(confmat: torch.Tensor, ece_bins: Optional[torch.Tensor] = None, group=None) -> None:
    """Sum the evaluation counters over all ranks IN PLACE with a single all-reduce.

    The arguments afterwards hold GLOBAL sums: do not keep accumulating into them and do not reduce them a second
    time (every count would be multiplied by the world size) -- end-of-sweep use only.  For live accumulators use
    `reduced_counts`, which reduces a copy."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    buf = pack_counts(confmat, ece_bins)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
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
