"""The per-scan hot path end to end: project -> (backbone logits) -> uncertainty -> metrics -> back-project.

`ScanEvaluator` is the call a user of the reference's test loop makes instead of the loader item
(src/dataset/dataloader_semantic_KITTI.py:31-99: remap, projection, range, normals) plus the MC block of
Tester.test_epoch (src/models/tester.py:395-471): it owns the device accumulators (confusion matrix,
reliability bins), runs a batch of scans with 7 kernel launches and never synchronises until `summary()`.

  step_device(...)  inputs already in HBM (what a GPU backbone hands over)
  step_host(...)    inputs in pinned host memory, one scan per chunk: the H2D copy of chunk i+1 overlaps the
                    kernels of chunk i on a second stream; the per-point labels and the four per-pixel maps
                    (pred, confidence, H_norm, MI_norm) come back to pinned host buffers
  capture(...) / replay()   the device step of a FIXED batch shape as one CUDA graph (one launch per step)
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib, dist as sdist, ops
from .dataset.definitions import build_id_lut

MAPS = ("pred", "conf", "H_norm", "MI_norm")


class ScanEvaluator:
    def __init__(self, height: int, width: int, num_classes: int, *, n_bins: int = 15, ignore_index: Optional[int] = 0,
                 lut: Optional[np.ndarray] = None, theta_range=None, eps: float = 1e-12, normals: bool = True, device=None):
        self.device = _lib.require_cuda(device)
        _lib.lib()
        self.H, self.W, self.C = int(height), int(width), int(num_classes)
        self.n_bins, self.ignore_index, self.eps = int(n_bins), ignore_index, float(eps)
        self.theta_range = theta_range
        self.want_normals = bool(normals)
        self.lut = torch.from_numpy(build_id_lut() if lut is None else np.asarray(lut, dtype=np.int32)).to(self.device)
        self.edges = ops.uniform_edges(self.n_bins)
        self.confmat = ops.new_confmat(self.C, self.device)
        self.ece_bins = ops.new_ece_bins(self.n_bins, self.device)
        self._ws = None
        self._copy_stream = None
        self._stage = None
        self._host_out = None
        self._graph = None
        self.launches = 0                 # kernels of libslu launched by this object's steps (slu_launch_count deltas)

    def reset(self):
        self.confmat.zero_()
        self.ece_bins.zero_()

    # ---------------------------------------------------------------- device-resident batch
    @torch.no_grad()
    def step_device(self, xyzi: torch.Tensor, raw_label: torch.Tensor, offsets: Sequence[int],
                    mc_logits: torch.Tensor, want=MAPS, timing=None) -> dict:
        """xyzi [n_total,4] f32, raw_label [n_total] i32, mc_logits [T,B,C,H,W] f32: all CUDA.

        Returns the image planes (x, y, z, range, intensity, label), the label map, the normals, the per-pixel
        uncertainty maps and the per-point predicted labels; adds this batch to the confusion matrix and
        reliability bins."""
        n0 = _lib.launch_count()
        with torch.cuda.device(self.device):
            proj = ops.project_batch(xyzi, raw_label, offsets, self.H, self.W, lut=self.lut, theta_range=self.theta_range,
                                     workspace=self._ws)
            self._ws = proj["workspace"]
            normals = ops.frame_normals(proj["img"]) if self.want_normals else None      # dataloader_semantic_KITTI.py:85
            if timing is not None:
                timing[0].record()
            red = ops.reduce_metrics(mc_logits, proj["label"], kind="logits", conf_mode=ops.CONF_RENORM, eps=self.eps,
                                     ignore_index=self.ignore_index, edges=self.edges, confmat=self.confmat,
                                     ece_bins=self.ece_bins, want=tuple(set(want) | {"pred"}))
            if timing is not None:
                timing[1].record()
            red["point_labels"] = ops.backproject(red["pred"], proj["pix"], offsets)
        red["img"], red["label"], red["pix"], red["normals"] = proj["img"], proj["label"], proj["pix"], normals
        red["near_edge"] = proj["diag"][:, 1]
        self.launches += _lib.launch_count() - n0
        return red

    # ---------------------------------------------------------------- the same step as one CUDA graph
    def capture(self, xyzi: torch.Tensor, raw_label: torch.Tensor, offsets: Sequence[int], mc_logits: torch.Tensor,
                want=MAPS) -> dict:
        """Record step_device for THIS batch shape (point offsets included: they are kernel parameters) into a CUDA
        graph.  The tensors passed here become the graph's static inputs -- write the next batch INTO them -- and the
        returned dict holds its static outputs, refreshed by every `replay()`."""
        with torch.cuda.device(self.device):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            keep = (self.confmat.clone(), self.ece_bins.clone())
            with torch.cuda.stream(side):                       # warm-up off the capture: workspace allocation, lazy init
                for _ in range(2):
                    self.step_device(xyzi, raw_label, offsets, mc_logits, want=want)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.confmat.copy_(keep[0]); self.ece_bins.copy_(keep[1])
            n0 = _lib.launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self.step_device(xyzi, raw_label, offsets, mc_logits, want=want)
            self._graph = (g, out, _lib.launch_count() - n0)
            self.launches -= self._graph[2]                     # capture enqueues nothing
        return out

    def replay(self) -> dict:
        if self._graph is None:
            raise RuntimeError("capture() first")
        self._graph[0].replay()
        self.launches += self._graph[2]
        return self._graph[1]

    # ---------------------------------------------------------------- host buffers, pipelined
    @torch.no_grad()
    def step_host(self, scans: Sequence[tuple], want=MAPS) -> list:
        """scans: sequence of (xyzi [N,4] f32, raw_label [N] i32/u32, mc_logits [T,1,C,H,W] f32), all PINNED host
        tensors (one entry per scan, as the reference's batch_size=1 test loop produces).

        Returns one dict per scan with HOST tensors: "point_labels" [N] int64 and the maps named in `want`
        ("pred" [H,W] int64, "conf", "H_norm", "MI_norm" [H,W] float32).  They are views of pinned buffers this
        object owns (one set per position in `scans`, sized for the largest scan seen): valid until the next
        step_host call -- copy what must outlive it.  Synchronises once, at the end."""
        dev = self.device
        with torch.cuda.device(dev):
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=dev)
            main = torch.cuda.current_stream(dev)
            cs = self._copy_stream
            nbuf = 3
            nmax = max(s[0].size(0) for s in scans)
            if self._stage is None or self._stage[0][2].shape != scans[0][2].shape or self._stage[0][0].size(0) < nmax:
                self._stage = [(torch.empty((nmax, 4), dtype=torch.float32, device=dev),
                                torch.empty((nmax,), dtype=torch.int32, device=dev),
                                torch.empty(scans[0][2].shape, dtype=torch.float32, device=dev)) for _ in range(nbuf)]
                self._free = [torch.cuda.Event() for _ in range(nbuf)]
            self._ensure_host_out(len(scans), nmax)
            outs = []
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
                red = self.step_device(sx[:n], sr[:n], [0, n], sl, want=want)
                self._free[i % nbuf].record(main)
                slot = self._host_out[i]
                res = {"point_labels": slot["point_labels"][:n]}
                res["point_labels"].copy_(red["point_labels"], non_blocking=True)
                for k in want:
                    slot[k].copy_(red[k][0], non_blocking=True)
                    res[k] = slot[k]
                outs.append(res)
            main.synchronize()
        return outs

    def _ensure_host_out(self, n_slots: int, nmax: int):
        """Pinned result buffers: one set per position in the batch, grown (never per distinct point count)."""
        if self._host_out is not None and len(self._host_out) >= n_slots and self._host_out[0]["point_labels"].numel() >= nmax:
            return
        cap = max(nmax, 0 if self._host_out is None else self._host_out[0]["point_labels"].numel())
        slots = max(n_slots, 0 if self._host_out is None else len(self._host_out))
        self._host_out = [{"point_labels": torch.empty((cap,), dtype=torch.int64, pin_memory=True),
                           "pred": torch.empty((self.H, self.W), dtype=torch.int64, pin_memory=True),
                           "conf": torch.empty((self.H, self.W), dtype=torch.float32, pin_memory=True),
                           "H_norm": torch.empty((self.H, self.W), dtype=torch.float32, pin_memory=True),
                           "MI_norm": torch.empty((self.H, self.W), dtype=torch.float32, pin_memory=True)} for _ in range(slots)]

    def host_bytes_per_scan(self, n_points: int, T: int, want=MAPS):
        """(h2d, d2h) bytes step_host moves for one scan of n_points."""
        h2d = n_points * 16 + n_points * 4 + T * self.C * self.H * self.W * 4
        d2h = n_points * 8 + sum(self.H * self.W * (8 if k == "pred" else 4) for k in want)
        return h2d, d2h

    # ---------------------------------------------------------------- results
    def counts(self, group=None, reduce_across_ranks=True):
        """(confmat, ece_bins) summed over the ranks with ONE int64 all-reduce of a packed COPY: the per-rank
        accumulators stay local, so calling this (or summary()) twice, or stepping on afterwards, is safe."""
        if reduce_across_ranks and sdist.world()[1] > 1:
            buf = sdist.pack_counts(self.confmat, self.ece_bins)
            sdist.allreduce_packed(buf, group=group)
            n = self.confmat.numel()
            return buf[:n].view_as(self.confmat), buf[n:].view_as(self.ece_bins)
        return self.confmat, self.ece_bins

    def summary(self, class_names=None, test_mask=None, ignore_gt=(0,), group=None, reduce_across_ranks=True) -> dict:
        """Combine the counters over ranks (one integer all-reduce of a copy) and turn them into mIoU / ECE / MCE."""
        confmat, ece_bins = self.counts(group=group, reduce_across_ranks=reduce_across_ranks)
        cm = confmat.cpu().double()
        if ignore_gt:
            cm[list(ignore_gt), :] = 0.0
        tp = cm.diag()
        denom = cm.sum(0) + cm.sum(1) - tp
        iou = torch.where(denom > 0, tp / denom.clamp_min(1), torch.full_like(tp, float("nan")))
        mask = torch.isfinite(iou)
        if test_mask is not None:
            mask &= torch.as_tensor(test_mask, dtype=torch.bool)
        ece, mce, n, acc, avg = ops.ece_from_bins(ece_bins)
        return {"mIoU": float(iou[mask].mean()) if mask.any() else float("nan"), "iou": iou.numpy(),
                "ece": ece, "mce": mce, "bin_n": n, "bin_acc": acc, "bin_conf": avg,
                "confmat": confmat.cpu().numpy()}
