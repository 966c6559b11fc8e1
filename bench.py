#!/usr/bin/env python
"""bench.py -- scans/s of the per-scan hot path (project + uncertainty + metrics + back-project), plus one leg per
BASELINE.json config so the driver's BENCH / SCALE records witness all five.

Headline workload (BASELINE.json configs[1]): a batch of 16 SemanticKITTI-shaped scans (HDL-64, 120 000 points ->
64x2048) with MC-dropout logits [T=20, B=16, C=20, 64, 2048] fp32.  One step = one batch =
  loader item on the device (projection 4 launches + normals 1)  ->  fused MC reduction + confusion / ECE histograms (1)
  ->  label back-projection (1)
i.e. what the reference does per scan with SemanticKitti.__getitem__ + the MC block of Tester.test_epoch + IoUEvaluator /
ECEAggregator updates.  Synthetic seeded inputs (datasets are not available offline).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path
  python bench.py --impl reference [...]                          the UNMODIFIED reference's CPU code on the host cores
                                                                  (oracle/_ref/reference_src.zip; oracle port if absent)
Under torchrun (N>1) every rank runs its own batch (weak scaling, scans shard by index, no data-path collective) and the
integer counters are combined with one NCCL all-reduce inside every timed window.  Prints ONE JSON line on rank 0.

Timing: W >= 3 warm-up steps, then R windows (default 15) of EXACTLY K steps, each bracketed by barrier + synchronize and
timed with CUDA events (max over ranks); `value` is taken from the MEDIAN window, min / max are reported beside it.
"""
from __future__ import annotations

import argparse
import hashlib
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


class Ctx:
    """rank / device / collective plumbing shared by the legs"""

    def __init__(self):
        import torch.distributed as dist
        from semanticlidarunc_b200 import _lib
        from semanticlidarunc_b200.dist import bind_to_gpu_numa_node
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.local_cpus = bind_to_gpu_numa_node(self.local)
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        _lib.lib()
        self.peak, self.peak_src = measured_peak_gbs()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, v: float) -> float:
        if self.world == 1:
            return float(v)
        t = torch.tensor([v], device=self.dev, dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, reps: int, inner: int = 1):
        """`reps` windows of `inner` calls: barrier + sync on both sides, CUDA events, max over ranks -> list of ms per window"""
        out = []
        for _ in range(reps):
            self.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(inner):
                fn()
            b.record()
            self.barrier()
            out.append(self.max_over_ranks(a.elapsed_time(b)))
        return out


def stats(ms_list, per=1.0):
    a = np.asarray(ms_list, dtype=np.float64) / per
    return {"median": round(float(np.median(a)), 5), "min": round(float(a.min()), 5), "max": round(float(a.max()), 5), "n": int(a.size)}


# ------------------------------------------------------------------------------------------------ headline (config 2)
def make_scans(rank: int, n: int, sensor=SENSOR):
    from semanticlidarunc_b200 import synth
    scans = [synth.synth_scan(1000 * rank + i, sensor) for i in range(n)]
    offs = np.concatenate([[0], np.cumsum([s[0].shape[0] for s in scans])]).astype(np.int64)
    return scans, offs


def leg_headline(cx: Ctx, args):
    from semanticlidarunc_b200 import _lib
    from semanticlidarunc_b200.pipeline import ScanEvaluator
    dev = cx.dev
    scans, offs = make_scans(cx.rank, B)
    xyzi = torch.from_numpy(np.concatenate([s[0] for s in scans])).to(dev)
    raw = torch.from_numpy(np.concatenate([s[1] for s in scans]).view(np.int32)).to(dev)
    g = torch.Generator(device=dev).manual_seed(1234 + cx.rank)
    logits = torch.randn((T, B, C, H, W), generator=g, device=dev, dtype=torch.float32) * 3.0
    ev = ScanEvaluator(H, W, C, n_bins=N_BINS, ignore_index=0, device=dev)

    for _ in range(max(args.warmup, 3)):
        out = ev.step_device(xyzi, raw, offs, logits)
    ev.counts()
    ev.reset()
    K, R = args.steps, args.windows
    kt = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    win, launches = [], []
    with ClockSampler(cx.local) as clk:
        for r in range(R):
            cx.barrier()
            n0 = _lib.launch_count()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for k in range(K):
                out = ev.step_device(xyzi, raw, offs, logits, timing=kt[k] if r == R - 1 else None)
            ev.counts()                                   # the sweep's single collective: packed int64 all-reduce of a copy
            t1.record()
            cx.barrier()
            launches.append(_lib.launch_count() - n0)
            win.append(cx.max_over_ranks(t0.elapsed_time(t1)))
    ms = float(np.median(win))
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kt]))
    summ = ev.summary(reduce_across_ranks=False)
    value = cx.world * B * K / (ms / 1e3)

    algo_bytes = BYTES_PER_PIXEL * B * H * W
    achieved = algo_bytes / (kernel_ms / 1e3) / 1e9
    roofline = {"bound": "hbm", "kernel": "reduce_staged_kernel<20,logits>", "achieved": round(achieved, 1),
                "peak": cx.peak, "peak_source": cx.peak_src, "unit": "GB/s", "frac": round(achieved / cx.peak, 4),
                "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": round(kernel_ms, 4), "traffic": None,
                "kernel_share_of_step": round(kernel_ms / (ms / K), 4),
                "frac_of_8TBps_datasheet": round(achieved / 8000.0, 4)}

    # argmax flips against torch: how many pixels of this rank's batch get another class than an fp32 / fp64 torch
    # evaluation of the same logits (near-ties of p_bar; any two correct fp32 implementations differ on some)
    flips = {"pixels": B * H * W, "vs_torch_fp32": 0, "vs_torch_fp64": 0, "torch_fp32_vs_fp64": 0}
    pred = out["pred"]
    for b in range(B):
        lb = logits[:, b]
        p32 = torch.softmax(lb, dim=1).mean(0).argmax(0)
        p64 = torch.softmax(lb.double(), dim=1).mean(0).argmax(0)
        flips["vs_torch_fp32"] += int((pred[b] != p32).sum())
        flips["vs_torch_fp64"] += int((pred[b] != p64).sum())
        flips["torch_fp32_vs_fp64"] += int((p32 != p64).sum())
    flips["rate_vs_torch_fp32"] = flips["vs_torch_fp32"] / flips["pixels"]

    head = {"value": value, "ms": ms, "windows": {"R": R, "steps_per_window": K, "ms_per_step": stats(win, K)},
            "gpu_launches": int(np.median(launches)), "roofline": roofline, "clocks": clk.summary(),
            "result_check": {"mIoU": summ["mIoU"], "ece": summ["ece"], "confmat_sum": int(summ["confmat"].sum())},
            "argmax_flips": flips, "points_per_scan": int(offs[1])}
    return head, (ev, scans, logits)


def leg_e2e(cx: Ctx, args, ev, scans, logits):
    """The same metric through the public host-buffer API: pinned host inputs, H2D + D2H inside the timed region."""
    host = [(torch.from_numpy(s[0]).pin_memory(), torch.from_numpy(s[1].view(np.int32)).pin_memory(),
             logits[:, i:i + 1].contiguous().cpu().pin_memory()) for i, s in enumerate(scans)]
    h2d = d2h = 0
    for s in scans:
        a, b_ = ev.host_bytes_per_scan(s[0].shape[0], T)
        h2d, d2h = h2d + a, d2h + b_
    ev.reset()
    steps = max(3, min(args.steps, 10))
    for _ in range(2):
        ev.step_host(host)
    cx.barrier()
    w0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        outs = ev.step_host(host)
    e1.record()
    cx.barrier()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - w0) * 1e3 if cx.world == 1 else 0.0)
    e2e_ms = cx.max_over_ranks(e2e_ms)
    # ceiling of the host side: the same pinned bytes copied to the device, all ranks at once, no kernels, no D2H
    devbuf = [tuple(torch.empty_like(t, device=cx.dev) for t in h) for h in host[:3]]
    def h2d_only():
        for i, h in enumerate(host):
            for src, dst in zip(h, devbuf[i % 3]):
                dst.copy_(src, non_blocking=True)
    h2d_only()
    cop = cx.timed(h2d_only, 3)
    ceil_gbs = h2d / (float(np.median(cop)) / 1e3) / 1e9
    got_gbs = h2d / (e2e_ms / steps / 1e3) / 1e9
    res = {"value": round(cx.world * B * steps / (e2e_ms / 1e3), 2), "unit": "scans/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "steps": steps, "ms_per_step": round(e2e_ms / steps, 3),
           "api": "ScanEvaluator.step_host: pinned host scans + logits in, per-point labels and pred / conf / H_norm / MI_norm "
                  "maps back to pinned host buffers; one scan per chunk, copy / compute overlap on two streams",
           "outputs_returned": sorted(outs[0].keys()),
           "h2d_gbs_per_gpu": round(got_gbs, 2), "h2d_ceiling_gbs_per_gpu": round(ceil_gbs, 2),
           "frac_of_ceiling": round(got_gbs / ceil_gbs, 4),
           "ceiling_note": "pure pinned H2D copies of the same bytes, all ranks concurrently, no kernels",
           "cpus_local_to_gpu": cx.local_cpus}
    del host, devbuf
    return res


# ------------------------------------------------------------------------------------------------ config 1
def leg_config1(cx: Ctx):
    """configs[0]: ONE HDL-64 scan -> 64x2048 through the literal drop-in spherical_projection() (numpy in, numpy out),
    then 20-class softmax entropy + ECE from single-pass logits (T = 1).  Wall clock including every copy."""
    from semanticlidarunc_b200 import ops, synth
    from semanticlidarunc_b200.dataset.definitions import build_id_lut
    from semanticlidarunc_b200.dataset.utils import spherical_projection
    dev = cx.dev
    xyzi, raw = synth.synth_scan(7, SENSOR)
    lut = build_id_lut()
    pc = np.concatenate([xyzi.astype(np.float64), lut[(raw & 0xFFFF).astype(np.int64)][:, None].astype(np.float64)], axis=1)
    logits_h = (torch.randn((1, C, H, W), generator=torch.Generator().manual_seed(3)) * 3.0).pin_memory()
    confmat, bins = ops.new_confmat(C, dev), ops.new_ece_bins(N_BINS, dev)

    def project():
        return spherical_projection(pc, H, W)

    def whole():
        img, _, _, _ = spherical_projection(pc, H, W)
        labels = torch.from_numpy(img[..., 4].astype(np.int64)).to(dev, non_blocking=True)[None]
        lg = logits_h.to(dev, non_blocking=True)
        r = ops.reduce_metrics(lg, labels, kind="logits", conf_mode=ops.CONF_RAW, ignore_index=0, confmat=confmat,
                               ece_bins=bins, want=("pred", "H_norm"))
        return r["H_norm"].cpu(), r["pred"].cpu()

    def wall(fn, n=12):
        fn(); fn()
        ts = []
        for _ in range(n):
            torch.cuda.synchronize()
            t = time.perf_counter(); fn(); torch.cuda.synchronize()
            ts.append((time.perf_counter() - t) * 1e3)
        return ts
    t_proj, t_all = wall(project), wall(whole)
    # the T = 1 kernel on a device-resident batch of 16 (roofline figure: 108 B/px)
    x16 = torch.randn((B, C, H, W), device=dev) * 3.0
    lab16 = torch.randint(0, C, (B, H, W), device=dev)
    fn16 = lambda: ops.reduce_metrics(x16, lab16, kind="logits", ignore_index=0, confmat=confmat, ece_bins=bins)
    fn1 = lambda: ops.reduce_metrics(x16[:1], lab16[:1], kind="logits", ignore_index=0, confmat=confmat, ece_bins=bins)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # larger than the 126 MB L2

    def graphed(fn):                  # GPU-side time of the kernel alone: CUDA graph replay, L2 flushed before every replay
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.replay()
        ts = []
        for _ in range(15):
            flush.zero_()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.replay(); b_.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b_))
        return float(np.median(ts))
    k16, k1 = graphed(fn16), graphed(fn1)
    del flush
    bpp = 4 * C + 8 + 8 + 4 + 4 + 4
    return {"workload": "single HDL-64 scan (120 000 pts) -> 64x2048, softmax entropy + ECE (T=1, C=20)",
            "spherical_projection_ms_wall": stats(t_proj), "scan_ms_wall": stats(t_all),
            "scans_per_s_wall": round(1e3 / float(np.median(t_all)), 1),
            "api": "dataset.utils.spherical_projection(numpy) + ops.reduce_metrics; H2D of the cloud / logits and D2H of image, H_norm, pred inside",
            "reduce_single_kernel": {"ms_b16": round(k16, 5), "ms_b1": round(k1, 5), "bytes_per_px": bpp, "timing": "CUDA graph replay, L2 flushed (256 MB write) before each replay, median of 15",
                                     "gbs_b16": round(bpp * B * H * W / k16 / 1e6, 1),
                                     "frac_of_peak_b16": round(bpp * B * H * W / k16 / 1e6 / cx.peak, 4),
                                     "frac_of_peak_b1": round(bpp * H * W / k1 / 1e6 / cx.peak, 4)}}


# ------------------------------------------------------------------------------------------------ config 3
def leg_config3(cx: Ctx):
    """configs[2]: OS1-128 scans (128x2048, 262 144 pts): projection + label back-projection, automatic elevation range
    (spherical_projection's default) and the fixed +-pi/8 range SemanticCUDAL uses; 16 scans per call, device-resident."""
    from semanticlidarunc_b200 import ops, synth
    from semanticlidarunc_b200.dataset.definitions import build_id_lut
    dev = cx.dev
    Hh, Ww, nb = 128, 2048, 16
    scans = [synth.synth_scan(300 + i, "os1-128") for i in range(nb)]
    offs = np.concatenate([[0], np.cumsum([s[0].shape[0] for s in scans])]).astype(np.int64)
    xyzi = torch.from_numpy(np.concatenate([s[0] for s in scans])).to(dev)
    raw = torch.from_numpy(np.concatenate([s[1] for s in scans]).view(np.int32)).to(dev)
    lut = torch.from_numpy(build_id_lut()).to(dev)
    pred = torch.randint(0, C, (nb, Hh, Ww), device=dev)
    ws = [None]
    n_pts = int(offs[-1])
    bytes_p = 20 * n_pts + 24 * nb * Hh * Ww                     # SURVEY 8d "P"
    bytes_b = 8 * n_pts + 8 * nb * Hh * Ww                       # SURVEY 8d "B"
    res = {"workload": "16 OS1-128 scans (262 144 pts each) -> 128x2048: projection + label back-projection",
           "algorithmic_bytes": {"projection": bytes_p, "backprojection": bytes_b}}
    for name, tr in (("auto_range", None), ("fixed_range_pi_8", (-np.pi / 8, np.pi / 8))):
        def step():
            p = ops.project_batch(xyzi, raw, offs, Hh, Ww, lut=lut, theta_range=tr, workspace=ws[0])
            ws[0] = p["workspace"]
            return ops.backproject(pred, p["pix"], offs), p
        for _ in range(3):
            back, p = step()
        def proj_only():
            q = ops.project_batch(xyzi, raw, offs, Hh, Ww, lut=lut, theta_range=tr, workspace=ws[0])
        def back_only():
            ops.backproject(pred, p["pix"], offs)
        t_all = float(np.median(cx.timed(step, 7, inner=10))) / 10
        t_p = float(np.median(cx.timed(proj_only, 7, inner=10))) / 10
        t_b = float(np.median(cx.timed(back_only, 7, inner=10))) / 10
        digest = hashlib.sha256(p["pix"].cpu().numpy().tobytes() + p["winner"].cpu().numpy().tobytes()).hexdigest()[:16]
        res[name] = {"ms_per_16_scans": round(t_all, 5), "scans_per_s": round(nb / t_all * 1e3, 1),
                     "projection_ms": round(t_p, 5), "backprojection_ms": round(t_b, 5),
                     "projection_frac_of_peak": round(bytes_p / t_p / 1e6 / cx.peak, 4),
                     "backprojection_frac_of_peak": round(bytes_b / t_b / 1e6 / cx.peak, 4),
                     "near_edge_points": int(p["diag"][:, 1].sum()), "pix_winner_sha256_16": digest}
    return res


# ------------------------------------------------------------------------------------------------ config 4
def leg_config4(cx: Ctx):
    """configs[3]: validation-sized sweep (4 000 scans of 64x2048): confusion-matrix mIoU + 15-bin ECE from reduced maps,
    STRONG scaling -- the 16 chunks of 250 scans are dealt to the ranks, ONE int64 all-reduce of the counts at the end.
    Rank 0 also runs the whole sweep alone (what N=1 computes) and the combined counters must equal it bit for bit."""
    from semanticlidarunc_b200 import dist as sdist, ops, synth
    dev = cx.dev
    N_SCANS, CHUNK = 4000, 250
    n_chunks = N_SCANS // CHUNK

    def chunk_data(k):
        g = torch.Generator(device=dev).manual_seed(1000 + k)             # a chunk's maps depend on its index only
        coarse = torch.randint(0, C, (CHUNK, H // 8, W // 32), generator=g, device=dev)
        lab = coarse.repeat_interleave(8, dim=1).repeat_interleave(32, dim=2).contiguous()     # spatially coherent labels
        pred = torch.where(torch.rand((CHUNK, H, W), generator=g, device=dev) < 0.85, lab,
                           torch.randint(0, C, (CHUNK, H, W), generator=g, device=dev))
        conf = 1.0 - 0.6 * torch.rand((CHUNK, H, W), generator=g, device=dev) ** 2
        return pred, lab, conf

    mine = list(range(cx.rank, n_chunks, cx.world))
    data = {k: chunk_data(k) for k in mine}
    cm, bins = ops.new_confmat(C, dev), ops.new_ece_bins(N_BINS, dev)
    result = {}

    def sweep():
        cm.zero_(); bins.zero_()
        for k in mine:
            pred, lab, conf = data[k]
            ops.confusion_ece(pred, lab, conf, num_classes=C, ignore_index=0, confmat=cm, ece_bins=bins)
        result["counts"] = sdist.reduced_counts(cm, bins)
    for _ in range(3):
        sweep()
    times = cx.timed(sweep, 9)
    gcm, gbins = result["counts"]
    digest = hashlib.sha256(gcm.cpu().numpy().tobytes() + gbins.cpu().numpy().tobytes()).hexdigest()[:16]
    # the single-process sweep over ALL chunks, on rank 0 (at N=1 this is the timed sweep itself)
    digest_n1, equal = digest, True
    if cx.world > 1:
        if cx.rank == 0:
            cm1, b1 = ops.new_confmat(C, dev), ops.new_ece_bins(N_BINS, dev)
            for k in range(n_chunks):
                pred, lab, conf = data[k] if k in data else chunk_data(k)
                ops.confusion_ece(pred, lab, conf, num_classes=C, ignore_index=0, confmat=cm1, ece_bins=b1)
            digest_n1 = hashlib.sha256(cm1.cpu().numpy().tobytes() + b1.cpu().numpy().tobytes()).hexdigest()[:16]
            equal = bool(torch.equal(cm1, gcm) and torch.equal(b1, gbins))
        cx.barrier()
    # int32 variant of the same sweep (12 B/px): same counters
    data32 = {k: (v[0].int(), v[1].int(), v[2]) for k, v in data.items()}
    cm32, b32 = ops.new_confmat(C, dev), ops.new_ece_bins(N_BINS, dev)
    def sweep32():
        cm32.zero_(); b32.zero_()
        for k in mine:
            pred, lab, conf = data32[k]
            ops.confusion_ece(pred, lab, conf, num_classes=C, ignore_index=0, confmat=cm32, ece_bins=b32)
        result["c32"] = sdist.reduced_counts(cm32, b32)
    for _ in range(2):
        sweep32()
    times32 = cx.timed(sweep32, 7)
    same32 = bool(torch.equal(result["c32"][0], gcm) and torch.equal(result["c32"][1], gbins))
    ms, ms32 = float(np.median(times)), float(np.median(times32))
    ece, mce, *_ = ops.ece_from_bins(gbins)
    tp = gcm.diag().double()
    den = gcm.sum(0).double() + gcm.sum(1).double() - tp
    nbytes = 20 * N_SCANS * H * W
    del data, data32
    torch.cuda.empty_cache()
    assert int(gcm.sum()) == N_SCANS * H * W, "sweep lost pixels"
    assert equal, "sharded counts differ from the single-process sweep"
    return {"workload": "validation sweep: 4000 scans of 64x2048, confusion + 15-bin ECE from reduced maps (pred i64, label i64, conf f32)",
            "scaling": "strong", "chunks_per_rank": len(mine), "ms_per_sweep": stats(times), "scans_per_s": round(N_SCANS / ms * 1e3),
            "aggregate_gbs": round(nbytes / ms / 1e6, 1), "frac_of_peak_per_gpu": round(nbytes / ms / 1e6 / cx.peak / cx.world, 4),
            "counts_sha256_16": digest, "n1_counts_sha256_16": digest_n1, "equals_single_process_counts": equal,
            "count_transport": sdist.count_transport(),
            "confmat_sum": int(gcm.sum()), "mIoU": round(float((tp / den.clamp_min(1))[1:].mean()), 6), "ece": round(ece, 6),
            "int32_maps": {"ms_per_sweep": stats(times32), "scans_per_s": round(N_SCANS / ms32 * 1e3),
                           "frac_of_peak_per_gpu_12B_px": round(12 * N_SCANS * H * W / ms32 / 1e6 / cx.peak / cx.world, 4),
                           "same_counts": same32}}


# ------------------------------------------------------------------------------------------------ config 5
def leg_config5(cx: Ctx):
    """configs[4]: the training-step data path, the global batch of 16 HDL-64 scans dealt to the ranks (STRONG scaling).
    Per rank: projection (4 launches) -> loader tensors (views of the projection's planes + the normals kernel, 1) -> fused evidential loss forward + backward from
    the head output [B_local, C+1, 64, 2048]; the ranks exchange ONE number (the valid-pixel count), launched on a side
    stream as soon as the labels exist -- by the count kernel itself over NVLink peer memory (csrc/slu_peer.cu) when all
    ranks share a node, by an NCCL all-reduce otherwise (`count_transport` says which).  Every rank asserts that its gradient equals, bit for bit, the matching slice of
    the single-process full-batch gradient.  Timed eager (through the Python API) and as one CUDA graph."""
    from semanticlidarunc_b200 import ops, synth
    from semanticlidarunc_b200.dataset.definitions import build_id_lut
    from semanticlidarunc_b200.losses.evidential import EvidentialLoss
    dev = cx.dev
    GB = 16
    lut = torch.from_numpy(build_id_lut()).to(dev)

    def load(ids):
        scans = [synth.synth_scan(s, SENSOR) for s in ids]
        offs = np.concatenate([[0], np.cumsum([s[0].shape[0] for s in scans])]).astype(np.int64)
        xyzi = torch.from_numpy(np.concatenate([s[0] for s in scans])).to(dev)
        raw = torch.from_numpy(np.concatenate([s[1] for s in scans]).view(np.int32)).to(dev)
        outs = torch.stack([torch.randn((C + 1, H, W), generator=torch.Generator().manual_seed(500 + s)) * 3.0 for s in ids]).to(dev)
        return xyzi, raw, offs, outs

    def step(xyzi, raw, offs, outs, crit, ws):
        proj = ops.project_batch(xyzi, raw, offs, H, W, lut=lut, workspace=ws[0])
        ws[0] = proj["workspace"]
        crit.prefetch_count(proj["label"])                 # count + its all-reduce on a side stream, off the critical path
        fr = ops.frame_tensors(proj["img"], label=proj["label"])   # range / reflectivity / xyz / semantics as views of the projection's planes, normals computed
        loss4, grad = crit.forward_backward(outs, proj["label"])
        return loss4, grad, fr

    mine = list(range(cx.rank, GB, cx.world))
    xyzi, raw, offs, outs = load(mine)
    crit = EvidentialLoss(1.0, 0.05, ignore_index=0, group=True)
    ws = [None]
    loss4, grad, _ = step(xyzi, raw, offs, outs, crit, ws)
    # ---- against the single-process full batch
    fx, fr_, fo, fouts = load(list(range(GB)))
    full4, fgrad, _ = step(fx, fr_, fo, fouts, EvidentialLoss(1.0, 0.05, ignore_index=0), [None])
    grad_equal = bool(torch.equal(grad, fgrad[mine]))
    tot = loss4[:3].double().clone()
    if cx.world > 1:
        cx.dist.all_reduce(tot)
    loss_rel = abs(float(tot[0]) - float(full4[0])) / abs(float(full4[0]))
    del fx, fr_, fouts, fgrad
    ok = torch.tensor([1.0 if grad_equal else 0.0], device=dev)
    if cx.world > 1:
        cx.dist.all_reduce(ok, op=cx.dist.ReduceOp.MIN)
    all_equal = bool(ok.item() == 1.0)

    for _ in range(3):
        step(xyzi, raw, offs, outs, crit, ws)
    eager = cx.timed(lambda: step(xyzi, raw, offs, outs, crit, ws), 15)
    # ---- the same step as one CUDA graph (kernels + the NCCL all-reduce)
    graph_stats, graph_equal = None, None
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                step(xyzi, raw, offs, outs, crit, ws)
        torch.cuda.current_stream().wait_stream(side)
        cx.barrier()
        cg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cg):
            g4, ggrad, _ = step(xyzi, raw, offs, outs, crit, ws)
        cg.replay()
        cx.barrier()
        graph_equal = bool(torch.equal(ggrad, grad))
        graph_stats = cx.timed(cg.replay, 15)
    except Exception as e:                                    # capture not possible here: the eager number stands alone
        graph_equal = "capture failed: %s" % (str(e).splitlines()[0][:160],)
    # kernel-only time of the loss on this rank's shard (packed f32x2 kernel), for the roofline table
    def loss_only():
        crit_local.forward_backward(outs, lab_local)
    crit_local = EvidentialLoss(1.0, 0.05, ignore_index=0)
    lab_local = ops.project_batch(xyzi, raw, offs, H, W, lut=lut, workspace=ws[0])["label"]
    for _ in range(3):
        loss_only()
    t_loss = float(np.median(cx.timed(loss_only, 7, inner=10))) / 10
    ms = float(np.median(eager))
    res = {"workload": "training-step data path: projection + loader tensors + fused evidential loss fwd+bwd, global batch 16 HDL-64 scans",
           "scaling": "strong", "scans_per_rank": len(mine), "eager_ms_per_step": stats(eager),
           "eager_scans_per_s": round(GB / ms * 1e3, 1),
           "graph_ms_per_step": None if graph_stats is None else stats(graph_stats),
           "graph_scans_per_s": None if graph_stats is None else round(GB / float(np.median(graph_stats)) * 1e3, 1),
           "graph_grad_equals_eager": graph_equal,
           "count_transport": crit.count_transport(dev),
           "shard_grad_equals_full_batch_slice_bitwise": all_equal, "loss_sum_of_shares_rel_err": loss_rel,
           "full_batch_loss": float(full4[0]),
           "loss_kernel": {"ms_local_shard": round(t_loss, 5), "bytes_per_px": 176,
                           "frac_of_peak": round(176 * len(mine) * H * W / t_loss / 1e6 / cx.peak, 4),
                           "note": "count kernel + packed (f32x2) forward+backward kernel through EvidentialLoss.forward_backward"}}
    assert all_equal, "a rank's gradient differs from the single-process slice"
    assert loss_rel < 1e-6, "loss shares do not add up to the full-batch loss"
    return res


# ------------------------------------------------------------------------------------------------ ncu traffic of this run
def measure_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE reduce_staged_kernel launch of the headline shape, measured now
    by running tools/run_reduce_once.py under ncu (rank 0, N=1).  None when ncu is unavailable / closed on this pool."""
    import csv
    import io
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    script = os.path.join(ROOT, "tools", "run_reduce_once.py")
    if not os.path.exists(ncu) or not os.path.exists(script):
        return None, "ncu or tools/run_reduce_once.py missing"
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none",
           "-k", "regex:reduce_staged", "-c", "1", "--csv", sys.executable, script]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT)
    except Exception as e:
        return None, "ncu run failed: %s" % (str(e)[:100],)
    tot, unit_mult = 0.0, {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rows = [l for l in r.stdout.splitlines() if l.startswith('"')]
    try:
        for row in csv.DictReader(io.StringIO("\n".join(rows))):
            if row.get("Metric Name", "").startswith("dram__bytes"):
                tot += float(row["Metric Value"].replace(",", "")) * unit_mult.get(row.get("Metric Unit", "byte"), 1.0)
    except Exception as e:
        return None, "ncu output not understood: %s" % (str(e)[:100],)
    if tot <= 0:
        tail = (r.stdout + r.stderr).strip().splitlines()[-1:] or [""]
        return None, "ncu gave no dram counters: %s" % tail[0][:160]
    return int(tot), "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum on tools/run_reduce_once.py, this run"


# ------------------------------------------------------------------------------------------------ CPU baselines
def reference_inputs(n_scans, n_logits):
    from semanticlidarunc_b200 import synth
    scans = [synth.synth_scan(i, SENSOR) for i in range(n_scans)]
    g = torch.Generator().manual_seed(99)
    lg = [torch.randn((T, 1, C, H, W), generator=g) * 3.0 for _ in range(n_logits)]
    return scans, [lg[i % n_logits] for i in range(n_scans)]


def versions():
    return {"numpy": np.__version__, "torch": torch.__version__, "cpu_count": os.cpu_count(),
            "torch_threads": torch.get_num_threads()}


def oracle_port_rate(n_scans, budget_s):
    """Fallback when the packed reference is absent: the oracle PORT of the same path."""
    from oracle import metrics as om, projection as oproj, uncertainty as ou
    from semanticlidarunc_b200.dataset.definitions import build_id_lut
    scans, logits = reference_inputs(min(n_scans, 4), 1)
    lut = build_id_lut()
    confmat = torch.zeros((C, C), dtype=torch.long)
    done, t0 = 0, time.perf_counter()
    while done < n_scans and time.perf_counter() - t0 < budget_s:
        s = scans[done % len(scans)]
        fr = oproj.kitti_frame(s[0], s[1], H, W, lut)
        labels = torch.from_numpy(fr["semantics"])
        r = ou.mc_reduce(logits[0])
        confmat += om.confusion_counts(r["pred"], labels, C)
        om.ece_samples(r["p_bar"], labels, "probs", ignore_index=0)
        r["pred"][0].reshape(-1)[torch.from_numpy(fr["pix"])]
        done += 1
    return done / (time.perf_counter() - t0), done


def cpu_baseline(budget_s=12.0):
    """The reference's CPU path on this box's host cores, bounded sample of the headline workload, in the two forms
    SURVEY 8d names: single process (loader in line) and loader in DataLoader worker processes."""
    from oracle import ref_arm
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    if not ref_arm.available():
        v, n = oracle_port_rate(64, 2 * budget_s)
        return {"value": round(v, 3), "unit": "scans/s", "cores": cores, "kind": "port", "versions": versions(),
                "sample": f"{n} HDL-64 scans through the oracle port (reference archive oracle/_ref absent)"}
    scans, logits = reference_inputs(8, 2)
    workers = max(1, min(16, cores // 2))
    per_step = 8
    single = ref_arm.time_reference(scans, logits, H=H, W=W, C=C, steps=64, warmup=1, scans_per_step=per_step, workers=0, budget_s=budget_s)
    pooled = ref_arm.time_reference(scans, logits, H=H, W=W, C=C, steps=64, warmup=1, scans_per_step=per_step, workers=workers, budget_s=budget_s)
    best = max(single["scans_per_s"], pooled["scans_per_s"])
    return {"value": round(best, 3), "unit": "scans/s", "cores": cores, "kind": "reference", "versions": versions(),
            "single_process_scans_per_s": round(single["scans_per_s"], 3),
            "dataloader_workers": workers, "dataloader_workers_scans_per_s": round(pooled["scans_per_s"], 3),
            "sample": f"{single['steps'] * per_step} + {pooled['steps'] * per_step} HDL-64 scan passes (T={T}, C={C}, {H}x{W}) through the UNMODIFIED reference "
                      f"(SemanticKitti.__getitem__, Tester's MC block, IoUEvaluator, ECEAggregator; back-projection by definition), "
                      f"{single['seconds']:.1f} s + {pooled['seconds']:.1f} s wall, torch threads={torch.get_num_threads()}"}


def run_reference(args):
    """--impl reference: the UNMODIFIED reference's CPU implementation of the path (oracle/_ref/reference_src.zip, packed by
    oracle/make_ref.sh) on all host threads: a step = the FULL 16-scan batch, scan by scan as the reference's batch_size=1
    test loop does, loader in DataLoader workers as the reference runs it.  Rank 0 only (the reference is single-process)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_arm
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    steps, warm = args.steps, max(1, args.warmup)
    per_step = args.cpu_scans_per_step
    if ref_arm.available():
        scans, logits = reference_inputs(per_step, min(per_step, 4))
        workers = max(1, min(16, cores // 2))
        # bounded: stop after the budget; every step is the same work, so the rate of the completed steps stands
        res = ref_arm.time_reference(scans, logits, H=H, W=W, C=C, steps=steps, warmup=warm, scans_per_step=per_step,
                                     workers=workers, budget_s=args.reference_budget_s)
        v, done_steps, kind = res["scans_per_s"], res["steps"], "reference"
        sample = (f"{per_step} scans per step x {done_steps} steps through the UNMODIFIED reference (oracle/_ref archive of the reference src tree: "
                  f"SemanticKitti.__getitem__ in {workers} DataLoader workers, Tester's MC block closures, IoUEvaluator, ECEAggregator; "
                  f"back-projection by definition), torch threads={torch.get_num_threads()}, mIoU={res['mIoU']:.4f} ece={res['ece']:.4f}")
    else:
        v, n = oracle_port_rate(steps * per_step, args.reference_budget_s)
        done_steps, kind = max(1, n // per_step), "port"
        sample = f"{n} scan passes through the oracle port (reference archive oracle/_ref absent)"
    v = round(v, 3)
    print(json.dumps({
        "impl": "reference", "metric": "scans/sec (project+uncertainty+metrics)", "value": v, "unit": "scans/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": done_steps, "warmup": warm,
        "ms_per_step": round(per_step / v * 1e3, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "scans_per_step_per_gpu": per_step, "T": T, "C": C, "H": H, "W": W,
                   "note": "single host process + DataLoader workers, whatever N is: the reference is not distributed"},
        "cpu_baseline": {"value": v, "unit": "scans/s", "cores": cores, "kind": kind, "sample": sample, "versions": versions()},
        "e2e": {"value": v, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------ main
def run_ours(args):
    cx = Ctx()
    head, (ev, scans, logits) = leg_headline(cx, args)
    legs = {}
    e2e = None
    if not args.no_e2e:
        e2e = leg_e2e(cx, args, ev, scans, logits)
    del ev, logits
    torch.cuda.empty_cache()
    skip = set(args.skip.split(",")) if args.skip else set()
    for name, fn in (("config1_single_scan", leg_config1), ("config3_os1_128", leg_config3),
                     ("config4_validation_sweep", leg_config4), ("config5_training_step", leg_config5)):
        if name.split("_")[0] in skip:
            continue
        try:
            legs[name] = fn(cx)
        except AssertionError as e:
            legs[name] = {"FAILED_ASSERTION": str(e)}
        except Exception as e:
            legs[name] = {"error": "%s: %s" % (type(e).__name__, str(e).splitlines()[0][:200] if str(e) else "")}
        cx.barrier()
        torch.cuda.empty_cache()
    line = {
        "metric": "scans/sec (project+uncertainty+metrics)", "value": round(head["value"], 2), "unit": "scans/s",
        "n_gpus": cx.world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(head["ms"] / args.steps, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "scans_per_step_per_gpu": B, "T": T, "C": C, "H": H, "W": W,
                   "points_per_scan": head["points_per_scan"], "n_bins": N_BINS,
                   "step": "projection (4 launches) + normals (1) + fused MC reduction with confusion / ECE histograms (1) + back-projection (1)",
                   "l2": "inputs (3.36 GB logits per step) exceed the 126 MB L2; no flush needed",
                   "sharding": "by scan index, one int64 all-reduce of the counts per window",
                   "value_from": "median of R windows of exactly K steps"},
        "gpu_launches": head["gpu_launches"], "clocks": head["clocks"], "e2e": e2e, "roofline": head["roofline"],
        "windows": head["windows"], "result_check": head["result_check"], "argmax_flips": head["argmax_flips"],
        "legs": legs,
    }
    if cx.rank == 0 and cx.world == 1:
        if not args.no_ncu_traffic:
            tr, how = measure_traffic()
            if tr is None:
                fb = os.path.join(ROOT, "profiles", "traffic_r01.json")
                if os.path.exists(fb):
                    with open(fb) as f:
                        tr = json.load(f).get("reduce_staged_kernel_dram_bytes_per_launch")
                    how = "committed round-1 ncu capture (profiles/traffic_r01.json); live capture unavailable: " + how
            line["roofline"]["traffic"], line["roofline"]["traffic_source"] = tr, how
        if not args.no_cpu_baseline:
            try:
                os.sched_setaffinity(0, range(os.cpu_count() or 1))      # the CPU baseline may use every host core
            except Exception:
                pass
            try:
                line["cpu_baseline"] = cpu_baseline()
            except Exception as e:
                line["cpu_baseline"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
    if cx.rank == 0:
        print(json.dumps(line))
        sys.stdout.flush()
    cx.barrier()
    # a captured graph that holds an NCCL collective keeps the communicator busy at teardown (destroy_process_group can
    # wait forever): leave without the orderly shutdown once every rank has passed the barrier
    os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--windows", type=int, default=15, help="R: number of timed windows of K steps (median reported)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-scans-per-step", type=int, default=B, help="reference arm: scans per step (the full batch by default)")
    ap.add_argument("--reference-budget-s", type=float, default=150.0, help="reference arm: stop after this many seconds of timed work")
    ap.add_argument("--skip", default="", help="comma-separated legs to skip: config1,config3,config4,config5")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ncu-traffic", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-buffer leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
