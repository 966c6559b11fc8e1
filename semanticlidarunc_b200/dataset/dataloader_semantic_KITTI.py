"""`SemanticKitti` dataset with the reference's interface (src/dataset/dataloader_semantic_KITTI.py:15-99),
its per-item work done on the GPU.

Same constructor (`data_path, rotate, flip, resolution, projection, resize`) and the same five tensors
from `__getitem__`: range [1,H,W], reflectivity [1,H,W], xyz [3,H,W], normals [3,H,W] float32 and
semantics [1,H,W] int64.  The reference does the label remap in a per-point python loop, the projection
in numpy and the normals in OpenCV inside DataLoader worker processes (38 + 16 + 7 ms per scan); here
an item is: read the two files -> one H2D copy -> projection kernels (label LUT, optional yaw) ->
slu_frame_tensors (resize, flip, range, normals).  Use it with `num_workers=0`: the loader is no longer
CPU-bound, and CUDA work does not belong in forked workers.  `device_batch()` projects many scans at once.
"""
from __future__ import annotations

import numpy as np
import torch
from torch.utils.data import Dataset

from .. import _lib, ops
from .definitions import build_id_lut, id_map


class SemanticKitti(Dataset):
    THETA_RANGE = None            # per-scan theta min/max (dataloader_semantic_KITTI.py:58)
    RESIZE_TO = (128, 2048)       # cv2.resize(..., (2048,128)) at :62
    LABEL_MAP = id_map

    def __init__(self, data_path, rotate=False, flip=False, resolution=(2048, 128), projection=(64, 2048), resize=True,
                 *, device=None, return_device: bool = False, label_map=None):
        self.data_path = data_path
        self.rotate = rotate
        self.flip = flip
        self.resolution = resolution
        self.projection = projection
        self.resize = resize
        self.return_device = return_device
        self._device = device
        self._lut_host = build_id_lut(self.LABEL_MAP if label_map is None else label_map)
        self._lut = None

    def __len__(self):
        return len(self.data_path)

    # -- helpers ---------------------------------------------------------------------------------
    def _dev(self):
        dev = _lib.require_cuda(self._device)
        if self._lut is None or self._lut.device != dev:
            self._lut = torch.from_numpy(self._lut_host).to(dev)
        return dev

    def read_scan(self, frame_path, label_path):
        """The on-disk pair: float32 [N,4] x,y,z,intensity and uint32 [N] (semantic id | instance << 16)."""
        xyzi = np.fromfile(frame_path, dtype=np.float32).reshape(-1, 4)
        label = np.fromfile(label_path, dtype=np.uint32).reshape(-1)
        return xyzi, label

    def device_batch(self, scans, yaw_deg=None, flip=None, settle_edges: bool = False):
        """scans: list of (xyzi, raw_label) numpy pairs -> dict of stacked device tensors
        (range, reflectivity, xyz, normals, semantics) plus pix / offsets for back-projection.

        settle_edges: read the kernels' near-edge counters (one synchronisation) and, for a scan that has points
        within 4 ulp of a bin edge, settle those points with numpy's own arithmetic before the image is framed
        (`_settle_scan`; ~1e-14 of all points are affected, so this is almost always just the counter read).
        `__getitem__` turns it on -- it synchronises anyway; batched callers that must not synchronise leave it off
        and can inspect out["near_edge"] later."""
        dev = self._dev()
        offs = np.concatenate([[0], np.cumsum([s[0].shape[0] for s in scans])]).astype(np.int64)
        xyzi = torch.from_numpy(np.ascontiguousarray(np.concatenate([s[0] for s in scans]))).to(dev, non_blocking=True)
        raw = torch.from_numpy(np.ascontiguousarray(np.concatenate([s[1] for s in scans])).view(np.int32)).to(dev, non_blocking=True)
        proj = ops.project_batch(xyzi, raw, offs, self.projection[0], self.projection[1], lut=self._lut, yaw_deg=yaw_deg,
                                 theta_range=self.THETA_RANGE, want_label=False)
        missing = proj["diag"][:, 0]
        settled = 0
        if settle_edges:
            diag = proj["diag"].cpu()
            for b in np.nonzero(diag[:, 1].numpy())[0]:
                settled += self._settle_scan(proj, int(b), scans[int(b)], offs, None if yaw_deg is None else yaw_deg[int(b)])
        out = self._frame(proj["img"], flip)
        out["pix"], out["offsets"], out["missing_label_ids"] = proj["pix"], offs, missing
        out["near_edge"], out["settled_points"] = proj["diag"][:, 1], settled
        return self._finish(out)

    def _settle_scan(self, proj, b, scan, offs, yaw):
        """Bit-exactness by construction for scan b of a projected batch: recompute every point's bin with numpy (the
        reference's arithmetic, dataset/utils.py::_host_bins), and re-resolve on the host the few pixels whose membership
        differs from the device's; the device image planes, label plane and pix are patched in place."""
        from .utils import _host_bins
        xyzi, raw = scan
        H, W = self.projection
        xyz = xyzi[:, :3].astype(np.float64)
        if yaw is not None:                                     # rotate_z, src/dataset/utils.py:4-18
            a = np.radians(float(yaw))
            c, s_ = np.cos(a), np.sin(a)
            xyz = xyz @ np.array([[c, -s_, 0.0], [s_, c, 0.0], [0.0, 0.0, 1.0]])
        host, _ = _host_bins(xyz, H, W, self.THETA_RANGE)
        n0, n1 = int(offs[b]), int(offs[b + 1])
        pix = proj["pix"][n0:n1].cpu().numpy().astype(np.int64)
        changed = np.nonzero(host != pix)[0]
        if changed.size == 0:
            return 0
        r = np.sqrt(xyz[:, 0] ** 2 + xyz[:, 1] ** 2 + xyz[:, 2] ** 2)
        sem = self._lut_host[(raw & 0xFFFF).astype(np.int64)]
        img = proj["img"][b].reshape(6, H * W)
        for q in np.unique(np.concatenate([host[changed], pix[changed]])):
            members = np.nonzero(host == q)[0]
            vals = np.zeros(6, dtype=np.float32)
            if members.size:
                rr = r[members]
                w = int(members[rr == rr.min()].min())          # nearest wins, lowest index among exact ties
                x32 = xyz[w].astype(np.float32)
                vals[:3] = x32
                vals[3] = np.sqrt(np.float32(x32[0] * x32[0]) + np.float32(x32[1] * x32[1]) + np.float32(x32[2] * x32[2]), dtype=np.float32)
                vals[4] = xyzi[w, 3]
                vals[5] = np.float32(sem[w])
            img[:, int(q)] = torch.from_numpy(vals).to(img.device)
        proj["pix"][n0:n1][torch.from_numpy(changed).to(proj["pix"].device)] = torch.from_numpy(host[changed].astype(np.int32)).to(proj["pix"].device)
        return int(changed.size)

    def staged_batches(self, indices, batch_size: int = 16, n_slots: int = 32, n_io_threads: int = 4, max_points: int = 300_000):
        """Yield `device_batch`-style dicts for `indices` in chunks of `batch_size`, the files read by libslu's native
        I/O threads (dataset/stager.py) one batch ahead of the projection kernels.  Augmentations are drawn per scan
        from numpy's global RNG in the reference's order, as in __getitem__."""
        from .stager import ScanStager
        dev = self._dev()
        indices = list(indices)
        with ScanStager(n_slots=max(n_slots, 2 * batch_size), max_points=max_points, n_io_threads=n_io_threads, device=dev) as st:
            chunks = [indices[i:i + batch_size] for i in range(0, len(indices), batch_size)]
            tickets = [[st.submit(*self.data_path[j]) for j in chunks[0]]] if chunks else []
            for k, chunk in enumerate(chunks):
                if k + 1 < len(chunks):                                  # the next batch's files are read while this one runs
                    tickets.append([st.submit(*self.data_path[j]) for j in chunks[k + 1]])
                xyzi = torch.empty((len(chunk) * max_points, 4), dtype=torch.float32, device=dev)
                raw = torch.empty((len(chunk) * max_points,), dtype=torch.int32, device=dev)
                offs, pos = [0], 0
                for t in tickets[k]:
                    n, has = st.fetch_into(t, xyzi[pos:pos + max_points], raw[pos:pos + max_points])
                    if not has:
                        raw[pos:pos + n].zero_()
                    pos += n
                    offs.append(pos)
                aug = [self._draw_augmentation() for _ in chunk]
                yaw = [0.0 if a[0] is None else a[0] for a in aug] if self.rotate else None
                offs = np.asarray(offs, dtype=np.int64)
                proj = ops.project_batch(xyzi[:pos], raw[:pos], offs, self.projection[0], self.projection[1], lut=self._lut,
                                         yaw_deg=yaw, theta_range=self.THETA_RANGE, want_label=False)
                out = self._frame(proj["img"], [a[1] for a in aug])
                out["pix"], out["offsets"], out["missing_label_ids"] = proj["pix"], offs, proj["diag"][:, 0]
                yield self._finish(out)

    def _frame(self, img, flip):
        return ops.frame_tensors(img, out_hw=self.RESIZE_TO if self.resize else None, flip=flip)

    def _finish(self, out):
        return out

    def _draw_augmentation(self):
        """(yaw in degrees or None, flip) drawn from numpy's global RNG in the reference's order (:53, :71)."""
        yaw = float(np.random.randint(-180, 180)) if self.rotate else None
        return yaw, bool(self.flip and np.random.rand() < 0.5)

    # -- Dataset -----------------------------------------------------------------------------------
    def __getitem__(self, idx):
        frame_path, label_path = self.data_path[idx]
        xyzi, label = self.read_scan(frame_path, label_path)
        yaw, do_flip = self._draw_augmentation()
        out = self.device_batch([(xyzi, label)], yaw_deg=None if yaw is None else [yaw], flip=[do_flip], settle_edges=True)
        if int(out["missing_label_ids"][0]) != 0:
            raise KeyError("scan %s contains semantic ids that are not in the label map" % (label_path,))   # id_map[l] at :47
        items = tuple(out[k][0] for k in ("range", "reflectivity", "xyz", "normals", "semantics"))
        return items if self.return_device else tuple(t.cpu() for t in items)
