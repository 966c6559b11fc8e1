"""`SemanticTHAB` with the reference's interface (src/dataset/dataloader_semantic_THAB.py:13-84) on the GPU.

SemanticTHAB scans come from an Ouster OS-128 and are already organised: 128 x 2048 points per file, pixel =
point index, no projection (documentation/dataset.md:109).  An item is: read the two files -> one H2D copy ->
slu_organized_planes (label LUT, flip, yaw as column roll + rotate_z, float64 range) -> slu_frame_tensors
(normals, tensor packing).  Back-projection of a label image to the points is the identity reshape.
"""
from __future__ import annotations

import numpy as np
import torch
from torch.utils.data import Dataset

from .. import _lib, ops
from .definitions import build_id_lut, id_map as _id_map

id_map = _id_map


class SemanticTHAB(Dataset):
    H, W = 128, 2048

    def __init__(self, data_path, rotate=False, flip=False, id_map=_id_map, projection=None, resize=False,
                 *, device=None, return_device: bool = False):
        self.id_map = id_map
        self.data_path = data_path
        self.rotate = rotate
        self.flip = flip
        self.projection = projection
        self.resize = resize
        self.return_device = return_device
        self._device = device
        self._lut_host = build_id_lut(id_map)
        self._lut = None

    def __len__(self):
        return len(self.data_path)

    def _dev(self):
        dev = _lib.require_cuda(self._device)
        if self._lut is None or self._lut.device != dev:
            self._lut = torch.from_numpy(self._lut_host).to(dev)
        return dev

    def device_batch(self, scans, flip=None, yaw_deg=None):
        """scans: list of (xyzi [H*W,4] float32, raw_label [H*W] uint32) -> stacked device tensors."""
        dev = self._dev()
        xyzi = torch.from_numpy(np.ascontiguousarray(np.concatenate([s[0] for s in scans]))).to(dev, non_blocking=True)
        raw = torch.from_numpy(np.ascontiguousarray(np.concatenate([s[1] for s in scans])).view(np.int32)).to(dev, non_blocking=True)
        shift = None
        if yaw_deg is not None:
            # rotate_equirectangular_image (src/dataset/utils.py:21-28) is called with the angle in DEGREES but
            # divides by 2*pi: the roll is round(angle / (2*pi) * W) columns.  Reproduced as is.
            shift = [int(round((float(a) / (2 * np.pi)) * self.W)) for a in np.broadcast_to(np.asarray(yaw_deg, dtype=np.float64), (len(scans),))]
        org = ops.organized_planes(xyzi, raw, self.H, self.W, lut=self._lut, flip=flip, col_shift=shift, yaw_deg=yaw_deg)
        out = ops.frame_tensors(org["img"])
        out["missing_label_ids"] = org["missing"]
        return out

    def __getitem__(self, idx):
        frame_path, label_path = self.data_path[idx]
        xyzi = np.fromfile(frame_path, dtype=np.float32).reshape(-1, 4)
        label = np.fromfile(label_path, dtype=np.uint32).reshape(-1)
        # the reference draws the flip coin first (:52), then the angle (:56)
        do_flip = bool(np.random.choice([True, False])) if self.flip else False
        yaw = int(np.random.randint(-180, 180)) if self.rotate else None
        out = self.device_batch([(xyzi, label)], flip=[do_flip], yaw_deg=None if yaw is None else [yaw])
        if int(out["missing_label_ids"][0]) != 0:
            raise KeyError("scan %s contains semantic ids that are not in the label map" % (label_path,))
        items = tuple(out[k][0] for k in ("range", "reflectivity", "xyz", "normals", "semantics"))
        return items if self.return_device else tuple(t.cpu() for t in items)
