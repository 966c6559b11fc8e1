"""`SemanticWADS` with the reference's interface (src/dataset/dataloader_semantic_WADS.py:82-155) on the GPU:
the KITTI flow with the elevation range fixed to +-pi/2 (:124), image rows without any return dropped before the
resize (:125), a 64x1024 target size (:127), WADS' label map (:12-51: SemanticKITTI's plus 110/111 -> 20) and the
flip coin drawn with np.random.choice (:131)."""
from __future__ import annotations

import numpy as np

from .. import ops
from .dataloader_semantic_KITTI import SemanticKitti
from .definitions import id_map as _kitti_id_map

id_map = dict(_kitti_id_map)
id_map[110] = 20
id_map[111] = 20


class SemanticWADS(SemanticKitti):
    THETA_RANGE = (-np.pi / 2, np.pi / 2)
    RESIZE_TO = (64, 1024)
    LABEL_MAP = id_map

    def __init__(self, data_path, rotate=False, flip=False, resolution=(2048, 128), projection=(64, 2048), resize=True,
                 remap_adverse_label=False, **kw):
        super().__init__(data_path, rotate=rotate, flip=flip, resolution=resolution, projection=projection, resize=resize, **kw)
        self.remap_adverse_label = remap_adverse_label      # accepted and unused, as in the reference (:114-115)

    def _draw_augmentation(self):
        yaw = float(np.random.randint(-180, 180)) if self.rotate else None
        do_flip = bool(np.random.choice([True, False])) if self.flip else False
        return yaw, do_flip

    def _frame(self, img, flip):
        out = ops.frame_tensors(img, out_hw=self.RESIZE_TO if self.resize else None, flip=flip, drop_empty_rows=True)
        if not self.resize:                       # variable height: keep the rows that exist (needs the count on the host)
            n = int(out["rows_kept"].max())
            for k in ("range", "reflectivity", "xyz", "normals", "semantics"):
                out[k] = out[k][:, :, :n].contiguous()
        return out
