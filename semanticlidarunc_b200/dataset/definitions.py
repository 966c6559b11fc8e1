"""SemanticKITTI label tables used by the loaders (data, not code).

Same names and values as the reference's src/dataset/definitions.py
(`id_map` :3-39, `id_map_reduced` :41-77, `id_map_dynamic` :79-114,
`color_map` :116-138, `class_names` :156-178, `custom_colormap` :205-213), kept
as one table so the raw-id -> train-id LUT that the projection kernel fuses
(`build_id_lut`) and the dict views come from a single source.
"""
import numpy as np

# raw SemanticKITTI id: (train id, reduced id, dynamic id or None if absent there)
_RAW = {
    0: (0, 0, 0), 1: (0, 0, 0), 9: (0, 0, None),
    10: (1, 1, 1), 11: (2, 2, 2), 13: (5, 3, 5), 15: (3, 2, 3), 16: (5, 3, 5),
    18: (4, 3, 4), 20: (5, 3, 5),
    30: (6, 4, 6), 31: (7, 5, 7), 32: (8, 5, 8),
    40: (9, 6, 0), 44: (10, 6, 0), 48: (11, 7, 0), 49: (12, 8, 0),
    50: (13, 9, 0), 51: (14, 9, 0), 52: (0, 0, 0),
    60: (19, 6, 0),
    70: (15, 7, 0), 71: (16, 7, 0), 72: (17, 10, 0),
    80: (18, 11, 0), 81: (19, 12, 0), 99: (0, 0, 0),
    252: (1, 1, 1), 253: (7, 5, 7), 254: (6, 6, 6), 255: (8, 5, 8),
    256: (5, 3, 5), 257: (5, 3, 5), 258: (4, 3, 4), 259: (5, 3, 5),
}

id_map = {k: v[0] for k, v in _RAW.items()}
id_map_reduced = {k: v[1] for k, v in _RAW.items()}
id_map_dynamic = {k: v[2] for k, v in _RAW.items() if v[2] is not None}

_NAMES = ("unlabeled car bicycle motorcycle truck other-vehicle person bicyclist motorcyclist "
          "road parking sidewalk other-ground building fence vegetation trunk terrain pole "
          "traffic-sign snow").split()
class_names = dict(enumerate(_NAMES))

_RGB = (
    (0, 0, 0), (245, 150, 100), (245, 230, 100), (150, 60, 30), (180, 30, 80), (255, 0, 0),
    (30, 30, 255), (200, 40, 255), (90, 30, 150), (125, 125, 125), (255, 150, 255), (75, 0, 75),
    (75, 0, 175), (0, 200, 255), (50, 120, 255), (0, 175, 0), (0, 60, 135), (80, 240, 150),
    (150, 240, 255), (250, 10, 250), (255, 255, 2),
)
color_map = {i: list(c) for i, c in enumerate(_RGB)}

_RGB_REDUCED = (
    (0, 0, 0), (245, 150, 100), (245, 230, 100), (255, 0, 0), (30, 30, 255), (200, 40, 255),
    (125, 125, 125), (75, 0, 75), (255, 150, 255), (0, 175, 0), (0, 60, 135), (150, 240, 255),
    (250, 250, 250),
)
color_map_reduced = {i: list(c) for i, c in enumerate(_RGB_REDUCED)}

# 256-entry BGR lookup table, shape [256,1,3] uint8, black where undefined
custom_colormap = np.zeros((256, 1, 3), dtype=np.uint8)
custom_colormap[: len(_RGB), 0, :] = np.asarray(_RGB, dtype=np.uint8)
custom_colormap = custom_colormap[..., ::-1]


def build_id_lut(mapping=None, size: int = 65536, missing: int = -1) -> np.ndarray:
    """Dense raw-id -> train-id table for the `label & 0xFFFF` domain.

    The reference remaps with a per-point python dict lookup
    (src/dataset/dataloader_semantic_KITTI.py:47), which raises KeyError on an
    id that is not in the map; the table marks such ids with `missing` so the
    device path can report them instead of silently remapping.
    """
    mapping = id_map if mapping is None else mapping
    lut = np.full(size, missing, dtype=np.int32)
    for k, v in mapping.items():
        lut[int(k)] = int(v)
    return lut
