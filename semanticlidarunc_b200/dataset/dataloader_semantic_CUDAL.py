"""`SemanticCUDAL` with the reference's interface (src/dataset/dataloader_semantic_CUDAL.py:54-125) on the
GPU: the KITTI flow with a 128x2048 projection over the fixed elevation range +-pi/8 (:95), CUDAL's label
map (:14-51: SemanticKITTI's plus raw id 2 -> 12), the flip coin drawn with np.random.choice (:101) and the
reflectivity normalised by max(max, 1) (:107)."""
from __future__ import annotations

import numpy as np
import torch

from .dataloader_semantic_KITTI import SemanticKitti
from .definitions import id_map as _kitti_id_map

id_map = dict(_kitti_id_map)
id_map[2] = 12


class SemanticCUDAL(SemanticKitti):
    THETA_RANGE = (-np.pi / 8, np.pi / 8)
    RESIZE_TO = (128, 2048)
    LABEL_MAP = id_map

    def __init__(self, data_path, rotate=False, flip=False, resolution=(2048, 128), projection=(128, 2048), resize=True, **kw):
        super().__init__(data_path, rotate=rotate, flip=flip, resolution=resolution, projection=projection, resize=resize, **kw)

    def _draw_augmentation(self):
        yaw = float(np.random.randint(-180, 180)) if self.rotate else None          # :93
        do_flip = bool(np.random.choice([True, False])) if self.flip else False     # :101
        return yaw, do_flip

    def _finish(self, out):
        r = out["reflectivity"]
        out["reflectivity"] = r / torch.clamp(r.amax(dim=(1, 2, 3), keepdim=True), min=1.0)   # :107, one value per scan
        return out
