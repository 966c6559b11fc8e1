"""The per-scan hot path end to end: project -> (backbone logits) -> uncertainty -> metrics -> back-project.

`ScanEvaluator` is the call a user of the reference's test loop makes instead of the MC block of
Tester.test_epoch (src/models/tester.py:395-471) plus the loader's projection
(src/dataset/dataloader_semantic_KITTI.py:35-97): it owns the device accumulators (confusion
matrix, reliability bins), runs a batch of scans with 6 kernel launches and never synchronises
until `summary()`.

  step_device(...)  inputs already in HBM (what a GPU backbone hands over)
  step_host(...)    inputs in pinned host memory, one scan per chunk: the H2D copy of chunk i+1
                    overlaps the kernels of chunk i on a second stream; per-point labels come back
                    to the host
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib, dist as sdist, ops
from .dataset.definitions import build_id_lut


class ScanEvaluator:
    def __init__(self, height: int, width: int, num_classes: int, *, n_bins: int = 15, ignore_index: Optional[int] = 0,
                 lut: Optional[np.ndarray] = None, theta_range=None, eps: float = 1e-12, device=None):
        self.device = _lib.require_cuda(device)
        _lib.lib()
        self.H, self.W, self.C = int(height), int(width), int(num_classes)
        self.n_bins, self.ignore_index, self.eps = int(n_bins), ignore_index, float(eps)
        self.theta_range = theta_range
        self.lut = torch.from_numpy(build_id_lut() if lut is None else np.asarray(lut, dtype=np.int32)).to(self.device)
        self.edges = ops.uniform_edges(self.n_bins)
        self.confmat = ops.new_confmat(self.C, self.device)
        self.ece_bins = ops.new_ece_bins(self.n_bins, self.device)
        self._ws = None
        self._copy_stream = None
        self._stage = None
        self.launches = 0                 # kernels of libslu launched so far

    def reset(self):
        self.confmat.zero_()
        self.ece_bins.zero_()

    # ---------------------------------------------------------------- device-resident batch
    @torch.no_grad()
    def step_device(self, xyzi: torch.Tensor, raw_label: torch.Tensor, offsets: Sequence[int],
                    mc_logits: torch.Tensor, want=("pred", "conf", "H_norm", "MI_norm"), timing=None) -> dict:
        """xyzi [n_total,4] f32, raw_label [n_total] i32, mc_logits [T,B,C,H,W] f32: all CUDA.

        Returns the image planes, per-pixel uncertainty maps and per-point predicted labels; adds
        this batch to the confusion matrix and reliability bins."""
        proj = ops.project_batch(xyzi, raw_label, offsets, self.H, self.W, lut=self.lut, theta_range=self.theta_range,
                                 workspace=self._ws)
        self._ws = proj["workspace"]
        if timing is not None:
            timing[0].record()
        red = ops.reduce_metrics(mc_logits, proj["label"], kind="logits", conf_mode=ops.CONF_RENORM, eps=self.eps,
                                 ignore_index=self.ignore_index, edges=self.edges, confmat=self.confmat,
                                 ece_bins=self.ece_bins, want=want)
        if timing is not None:
            timing[1].record()
        red["point_labels"] = ops.backproject(red["pred"], proj["pix"], offsets)
        red["img"], red["label"], red["pix"] = proj["img"], proj["label"], proj["pix"]
        self.launches += 6                # angles(+init,+extremes), rows, ties, resolve | reduce | back-project
        return red

    # ---------------------------------------------------------------- host buffers, pipelined
    @torch.no_grad()
    def step_host(self, scans: Sequence[tuple]) -> list:
        """scans: sequence of (xyzi [N,4] f32, raw_label [N] i32/u32, mc_logits [T,1,C,H,W] f32), all
        PINNED host tensors (one entry per scan, as the reference's batch_size=1 test loop produces).
        Returns the per-point predicted labels of every scan as host int64 tensors (ready after the
        returned event list's last event; this method synchronises once at the end)."""
        dev = self.device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        cs = self._copy_stream
        nbuf = 3
        if self._stage is None or self._stage[0][2].shape != scans[0][2].shape or self._stage[0][0].size(0) < max(s[0].size(0) for s in scans):
            nmax = max(s[0].size(0) for s in scans)
            self._stage = [(torch.empty((nmax, 4), dtype=torch.float32, device=dev),
                            torch.empty((nmax,), dtype=torch.int32, device=dev),
                            torch.empty(scans[0][2].shape, dtype=torch.float32, device=dev)) for _ in range(nbuf)]
            self._free = [torch.cuda.Event() for _ in range(nbuf)]
            self._out_host = {}
        outs, ready = [], []
        cs.wait_stream(main)
        for i, (xyzi, raw, logits) in enumerate(scans):
            sx, sr, sl = self._stage[i % nbuf]
            n = xyzi.size(0)
            with torch.cuda.stream(cs):
                if i >= nbuf:
                    cs.wait_event(self._free[i % nbuf])          # kernels that read this slot are done
                sx[:n].copy_(xyzi, non_blocking=True)
                sr[:n].copy_(raw.view(torch.int32) if raw.dtype != torch.int32 else raw, non_blocking=True)
                sl.copy_(logits, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cs)
            main.wait_event(ev)
            red = self.step_device(sx[:n], sr[:n], [0, n], sl, want=("pred",))
            self._free[i % nbuf].record(main)
            key = (i, n)
            if key not in self._out_host:
                self._out_host[key] = torch.empty((n,), dtype=torch.int64, pin_memory=True)
            self._out_host[key].copy_(red["point_labels"], non_blocking=True)
            outs.append(self._out_host[key])
        main.synchronize()
        return outs

    # ---------------------------------------------------------------- results
    def summary(self, class_names=None, test_mask=None, ignore_gt=(0,), group=None, reduce_across_ranks=True) -> dict:
        """Combine the counters over ranks (one integer all-reduce) and turn them into mIoU / ECE / MCE."""
        if reduce_across_ranks:
            sdist.allreduce_counts(self.confmat, self.ece_bins, group=group)
        cm = self.confmat.cpu().double()
        if ignore_gt:
            cm[list(ignore_gt), :] = 0.0
        tp = cm.diag()
        denom = cm.sum(0) + cm.sum(1) - tp
        iou = torch.where(denom > 0, tp / denom.clamp_min(1), torch.full_like(tp, float("nan")))
        mask = torch.isfinite(iou)
        if test_mask is not None:
            mask &= torch.as_tensor(test_mask, dtype=torch.bool)
        ece, mce, n, acc, avg = ops.ece_from_bins(self.ece_bins)
        return {"mIoU": float(iou[mask].mean()) if mask.any() else float("nan"), "iou": iou.numpy(),
                "ece": ece, "mce": mce, "bin_n": n, "bin_acc": acc, "bin_conf": avg,
                "confmat": self.confmat.cpu().numpy()}
