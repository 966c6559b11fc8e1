"""`SemanticSTF` with the reference's interface (src/dataset/dataloader_semantic_STF.py:15-95) on the GPU.

SemanticSTF files carry 5 float32 columns (the first four are x, y, z, intensity in 0..255) and labels that are
already train ids.  Decoding stays on the host, as file parsing does everywhere: intensity / 255 (:38), dropping
the sensor-clip points with ||xyz|| < 1.8 (:49-54); the label table is the identity with 20 -> 0 when
`remap_adverse_label` (:55-56).  Projection, resize, flip, range and normals run on the device as for KITTI."""
from __future__ import annotations

import numpy as np

from .dataloader_semantic_KITTI import SemanticKitti


class SemanticSTF(SemanticKitti):
    def __init__(self, data_path, rotate=False, flip=False, resolution=(2048, 128), projection=(64, 2048), resize=True,
                 remap_adverse_label=False, clip=True, **kw):
        table = {i: i for i in range(256)}
        if remap_adverse_label:
            table[20] = 0
        super().__init__(data_path, rotate=rotate, flip=flip, resolution=resolution, projection=projection, resize=resize,
                         label_map=table, **kw)
        self.remap_adverse_label = remap_adverse_label
        self.clip = clip

    def read_scan(self, frame_path, label_path):
        xyzi = np.fromfile(frame_path, dtype=np.float32).reshape(-1, 5)[:, :4].copy()
        xyzi[:, 3] /= 255.
        label = np.fromfile(label_path, dtype=np.uint32).reshape(-1)
        if self.clip:
            keep = np.where(np.linalg.norm(xyzi[:, 0:3], axis=-1) >= 1.8)
            label, xyzi = label[keep], xyzi[keep]
        return np.ascontiguousarray(xyzi), np.ascontiguousarray(label)

    def staged_batches(self, *a, **kw):
        # 5-column records and the r < 1.8 m / intensity filters happen on the host (read_scan): not the stager's 16-byte format
        raise NotImplementedError("SemanticSTF files are 5-column records filtered in read_scan(); use __getitem__ / device_batch")

    def _draw_augmentation(self):
        yaw = float(np.random.randint(-180, 180)) if self.rotate else None
        do_flip = bool(np.random.choice([True, False])) if self.flip else False
        return yaw, do_flip
