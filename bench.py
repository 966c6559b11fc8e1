#!/usr/bin/env python
"""bench.py -- scans/s of the per-scan hot path (project + uncertainty + metrics + back-project).

Workload (BASELINE.json configs[1]): a batch of 16 SemanticKITTI-shaped scans (HDL-64, 120 000
points -> 64x2048) with MC-dropout logits [T=20, B=16, C=20, 64, 2048] fp32.  One step = one batch:
  projection (4 launches) -> fused MC reduction + confusion/ECE histograms (1) -> label back-projection (1).
Synthetic seeded inputs (datasets are not available offline).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path
  python bench.py --impl reference [...]                          the reference's CPU algorithm
                                                                  (oracle port) on the host cores
Under torchrun (N>1) every rank runs its own batch (weak scaling, scans shard by index, no
data-path collective) and the integer counters are combined with one NCCL all-reduce inside the
timed region.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

T, B, C, H, W = 20, 16, 20, 64, 2048
SENSOR = "hdl64"
N_BINS = 15
BYTES_PER_PIXEL = 4 * T * C + 8 + 8 + 4 + 4 + 4          # 1628 B/px (logits + label in, 4 maps out)
WORKLOAD = "batch16_hdl64_mc_T20_C20_64x2048"


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def make_scans(rank: int, n: int):
    from semanticlidarunc_b200 import synth
    scans = [synth.synth_scan(1000 * rank + i, SENSOR) for i in range(n)]
    offs = np.concatenate([[0], np.cumsum([s[0].shape[0] for s in scans])]).astype(np.int64)
    return scans, offs


def run_ours(args):
    import torch.distributed as dist
    from semanticlidarunc_b200 import _lib
    from semanticlidarunc_b200.pipeline import ScanEvaluator

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    from semanticlidarunc_b200.dist import bind_to_gpu_numa_node
    local_cpus = bind_to_gpu_numa_node(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()

    scans, offs = make_scans(rank, B)
    xyzi = torch.from_numpy(np.concatenate([s[0] for s in scans])).to(dev)
    raw = torch.from_numpy(np.concatenate([s[1] for s in scans]).view(np.int32)).to(dev)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    logits = torch.randn((T, B, C, H, W), generator=g, device=dev, dtype=torch.float32) * 3.0
    ev = ScanEvaluator(H, W, C, n_bins=N_BINS, ignore_index=0, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        ev.step_device(xyzi, raw, offs, logits)
    ev.summary()
    ev.reset()
    ev.launches = 0
    kt = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    time.sleep(0.0)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with ClockSampler(local) as clk:
        t0.record()
        for k in range(args.steps):
            out = ev.step_device(xyzi, raw, offs, logits, timing=kt[k])
        from semanticlidarunc_b200 import dist as sdist
        sdist.allreduce_counts(ev.confmat, ev.ece_bins)          # the sweep's single collective
        t1.record()
        barrier()
    ms = t0.elapsed_time(t1)
    launches = ev.launches
    if world > 1:
        tmax = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kt]))
    summ = ev.summary(reduce_across_ranks=False)
    value = world * B * args.steps / (ms / 1e3)

    # ---- roofline of the dominant kernel (fused reduction + metrics), timed live above
    peak, peak_src = measured_peak_gbs()
    algo_bytes = BYTES_PER_PIXEL * B * H * W
    achieved = algo_bytes / (kernel_ms / 1e3) / 1e9
    roofline = {"bound": "hbm", "kernel": "reduce_staged_kernel<20,logits>", "achieved": round(achieved, 1),
                "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": round(kernel_ms, 4), "traffic": None,
                "kernel_share_of_step": round(kernel_ms / (ms / args.steps), 4),
                "frac_of_8TBps_datasheet": round(achieved / 8000.0, 4)}
    tr = os.path.join(ROOT, "profiles", "traffic_r01.json")
    if os.path.exists(tr):
        with open(tr) as f:
            roofline["traffic"] = json.load(f).get("reduce_staged_kernel_dram_bytes_per_launch")

    # ---- end to end through ScanEvaluator.step_host: pinned host buffers, H2D + D2H inside the timed region
    if args.no_e2e:
        return finish(args, world, rank, ms, value, launches, clk, None, roofline, summ, offs)
    host = [(torch.from_numpy(s[0]).pin_memory(), torch.from_numpy(s[1].view(np.int32)).pin_memory(),
             logits[:, i:i + 1].contiguous().cpu().pin_memory()) for i, s in enumerate(scans)]
    h2d = sum(a.numel() * 4 + b_.numel() * 4 + c_.numel() * 4 for a, b_, c_ in host)
    d2h = sum(a.size(0) * 8 for a, _, _ in host)
    ev.reset()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        ev.step_host(host)
    barrier()
    w0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_steps):
        labels_host = ev.step_host(host)
    e1.record()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - w0) * 1e3 if world == 1 else 0.0)
    if world > 1:
        tmax = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_ms = float(tmax.item())
    e2e = {"value": round(world * B * e2e_steps / (e2e_ms / 1e3), 2), "unit": "scans/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "ms_per_step": round(e2e_ms / e2e_steps, 3),
           "api": "ScanEvaluator.step_host (pinned host buffers, one scan per chunk, copy/compute overlap)",
           "cpus_local_to_gpu": local_cpus}

    finish(args, world, rank, ms, value, launches, clk, e2e, roofline, summ, offs)


def finish(args, world, rank, ms, value, launches, clk, e2e, roofline, summ, offs):
    import torch.distributed as dist
    line = {
        "metric": "scans/sec (project+uncertainty+metrics)", "value": round(value, 2), "unit": "scans/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "scans_per_step_per_gpu": B, "T": T, "C": C, "H": H, "W": W,
                   "points_per_scan": int(offs[1]), "n_bins": N_BINS,
                   "l2": "inputs (3.36 GB logits per step) exceed the 126 MB L2; no flush needed",
                   "sharding": "by scan index, one int64 all-reduce of counts per sweep"},
        "gpu_launches": int(launches), "clocks": clk.summary(), "e2e": e2e, "roofline": roofline,
        "result_check": {"mIoU": summ["mIoU"], "ece": summ["ece"], "confmat_sum": int(summ["confmat"].sum())},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            os.sched_setaffinity(0, range(os.cpu_count() or 1))      # the CPU baseline may use every host core
        except Exception:
            pass
        line["cpu_baseline"] = cpu_baseline(sample_scans=args.cpu_scans, budget_s=20.0)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def oracle_scan_pass(scan, logits_1, lut, confmat, samples):
    """The reference's CPU algorithm for ONE scan (oracle port): loader projection -> MC block ->
    IoU / ECE updates -> back-projection."""
    from oracle import metrics as om
    from oracle import projection as oproj
    from oracle import uncertainty as ou
    fr = oproj.kitti_frame(scan[0], scan[1], H, W, lut)
    labels = torch.from_numpy(fr["semantics"])
    r = ou.mc_reduce(logits_1)
    confmat += om.confusion_counts(r["pred"], labels, C)
    samples.append(om.ece_samples(r["p_bar"], labels, "probs", ignore_index=0))
    pix = fr["pix"]
    return r["pred"][0].reshape(-1)[torch.from_numpy(pix)]


def cpu_baseline(sample_scans: int, budget_s: float):
    from semanticlidarunc_b200 import synth
    from semanticlidarunc_b200.dataset.definitions import build_id_lut
    from oracle import metrics as om
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    lut = build_id_lut()
    g = torch.Generator().manual_seed(99)
    logits_1 = torch.randn((T, 1, C, H, W), generator=g) * 3.0
    confmat, samples = torch.zeros((C, C), dtype=torch.long), []
    oracle_scan_pass(synth.synth_scan(0, SENSOR), logits_1, lut, confmat, samples)      # warm-up
    samples.clear()
    done, t0 = 0, time.perf_counter()
    while done < sample_scans and time.perf_counter() - t0 < budget_s:
        oracle_scan_pass(synth.synth_scan(1 + done, SENSOR), logits_1, lut, confmat, samples)
        done += 1
    conf = torch.cat([s[0] for s in samples]).numpy()
    corr = torch.cat([s[1] for s in samples]).numpy()
    om.ece_from_stats(*om.ece_reference_stats(conf, corr, N_BINS))
    dt = time.perf_counter() - t0
    return {"value": round(done / dt, 3), "unit": "scans/s", "cores": cores, "kind": "port",
            "sample": f"{done} HDL-64 scans (T={T}, C={C}, {H}x{W}) through the oracle port of the reference path, "
                      f"{dt:.1f} s wall, torch threads={torch.get_num_threads()}, numpy projection single-threaded"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation (oracle port; the reference tree does not
    travel to the GPU box and is pure Python) on all host threads.  A step is a bounded sample of the
    workload: `--cpu-scans-per-step` scans of the 16-scan batch."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from semanticlidarunc_b200 import synth
    from semanticlidarunc_b200.dataset.definitions import build_id_lut
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    lut = build_id_lut()
    g = torch.Generator().manual_seed(99)
    logits_1 = torch.randn((T, 1, C, H, W), generator=g) * 3.0
    per_step = args.cpu_scans_per_step
    scans = [synth.synth_scan(i, SENSOR) for i in range(per_step)]
    confmat, samples = torch.zeros((C, C), dtype=torch.long), []
    steps, warm = min(args.steps, 8), min(max(args.warmup, 1), 2)
    for _ in range(warm):
        for s in scans:
            oracle_scan_pass(s, logits_1, lut, confmat, samples)
    samples.clear()
    t0 = time.perf_counter()
    for _ in range(steps):
        for s in scans:
            oracle_scan_pass(s, logits_1, lut, confmat, samples)
        samples.clear()
    dt = time.perf_counter() - t0
    v = round(steps * per_step / dt, 3)
    sample = (f"{per_step} of the {B} scans per step x {steps} steps, oracle port of the reference CPU path, "
              f"torch threads={torch.get_num_threads()}")
    print(json.dumps({
        "impl": "reference", "metric": "scans/sec (project+uncertainty+metrics)", "value": v, "unit": "scans/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": steps, "warmup": warm,
        "ms_per_step": round(dt / steps * 1e3, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "scans_per_step": per_step, "T": T, "C": C, "H": H, "W": W},
        "cpu_baseline": {"value": v, "unit": "scans/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-scans", type=int, default=32, help="upper bound of the cpu_baseline sample")
    ap.add_argument("--cpu-scans-per-step", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-buffer leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
