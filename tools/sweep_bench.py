#!/usr/bin/env python
"""BASELINE.json configs[3]: a validation-sized sweep (4 000 synthetic scans of 64x2048) of confusion-matrix mIoU +
15-bin ECE from reduced maps, scans sharded by index over the ranks, ONE int64 all-reduce of the counts at the end.

  python tools/sweep_bench.py                                        1 GPU
  torchrun --nproc-per-node N tools/sweep_bench.py                   N GPUs (weak data path, strong problem: 4 000 scans total)

Per rank: its scans' (pred int64, label int64, confidence fp32) maps are generated on the device before the timed region
(20 B/px, 2.6 MB per scan, 10.5 GB for the whole sweep), then the sweep = one slu_confusion_ece launch per 250-scan chunk
+ the all-reduce, timed with CUDA events (max over ranks), repeated 5 times.  The combined counts are checked against
the analytic total and are bit-identical for every N (integer sums).  Prints one JSON line on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from semanticlidarunc_b200 import dist as sdist, ops, synth  # noqa: E402

N_SCANS, CHUNK, C, H, W, N_BINS = 4000, 250, 20, 64, 2048, 15
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0

# chunk k holds scans [k*CHUNK, (k+1)*CHUNK); chunks are dealt to ranks round-robin (scan index -> rank by blocks)
my_chunks = list(range(rank, N_SCANS // CHUNK, world))
data = []
for k in my_chunks:
    g = torch.Generator(device=dev).manual_seed(1000 + k)             # a chunk's maps depend on its index only
    lab = synth.synth_coherent_labels(k, CHUNK, C, H, W, device="cpu").to(dev)
    pred = torch.where(torch.rand((CHUNK, H, W), generator=g, device=dev) < 0.85, lab, torch.randint(0, C, (CHUNK, H, W), generator=g, device=dev))
    conf = 1.0 - 0.6 * torch.rand((CHUNK, H, W), generator=g, device=dev) ** 2
    data.append((pred, lab, conf))
cm, bins = ops.new_confmat(C, dev), ops.new_ece_bins(N_BINS, dev)


def sweep():
    cm.zero_(); bins.zero_()
    for pred, lab, conf in data:
        ops.confusion_ece(pred, lab, conf, num_classes=C, ignore_index=0, confmat=cm, ece_bins=bins)
    sdist.allreduce_counts(cm, bins)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for _ in range(3):
    sweep()
times = []
for _ in range(5):
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); sweep(); b.record()
    barrier()
    t = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    times.append(float(t.item()))
ms = float(np.median(times))
ece, mce, *_ = ops.ece_from_bins(bins)
tp = cm.diag().double()
den = cm.sum(0).double() + cm.sum(1).double() - tp
iou = (tp / den.clamp_min(1))[1:]
if rank == 0:
    import hashlib
    digest = hashlib.sha256(cm.cpu().numpy().tobytes() + bins.cpu().numpy().tobytes()).hexdigest()[:16]
    nbytes = 20 * N_SCANS * H * W
    print(json.dumps({"config": "validation sweep: 4000 scans, confusion + 15-bin ECE from reduced maps", "n_gpus": world,
                      "ms_per_sweep": round(ms, 4), "scans_per_s": round(N_SCANS / ms * 1e3), "aggregate_GBps": round(nbytes / ms / 1e6, 1),
                      "frac_of_measured_peak_per_gpu": round(nbytes / ms / 1e6 / PEAK / world, 3), "launches_per_rank": len(data),
                      "confmat_sum": int(cm.sum()), "expected_confmat_sum": N_SCANS * H * W, "counts_sha256_16": digest,
                      "mIoU": round(float(iou.mean()), 6), "ece": round(ece, 6), "times_ms": [round(t, 4) for t in times]}))
if world > 1:
    dist.destroy_process_group()
