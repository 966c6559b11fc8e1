"""CPU restatement of the reference's spherical projection path.  TEST INFRASTRUCTURE ONLY.

Follows src/dataset/utils.py:61-67 (`to_deflection_coordinates`) and :288-349
(`spherical_projection`), and the KITTI loader glue
src/dataset/dataloader_semantic_KITTI.py:35-97.  Pinned against outputs of the
unmodified reference in tests/golden/ (see oracle/gen_golden.py).

The reference sorts points far->near and lets numpy's fancy-index scatter keep
the last write.  The restatement computes the same thing without the sort:
per-point (row, col) in the ORIGINAL point order plus an explicit depth test,
because those intermediates (which the reference hides) are what the CUDA
path is compared on.
"""
from __future__ import annotations

import numpy as np


def to_deflection_coordinates(x, y, z):
    """(phi, theta) exactly as src/dataset/utils.py:61-67 (dtype follows the input)."""
    p = np.sqrt(x ** 2 + y ** 2)
    phi = np.arctan2(y, x)
    theta = -np.arctan2(p, z) + np.pi / 2
    return phi, theta


def point_range(pc):
    """Sort key of src/dataset/utils.py:299: r = sqrt(x^2 + y^2 + z^2) in pc's dtype."""
    return np.sqrt(pc[:, 0] ** 2 + pc[:, 1] ** 2 + pc[:, 2] ** 2)


def projection_indices(pc, height, width, theta_range=None):
    """Per-point (row, col) of src/dataset/utils.py:319-339, in the caller's point order.

    `np.digitize(v, edges[::-1]) - 1` on descending edges equals
    `len(edges) - 1 - #{e <= v}`; the `-1` that results for v >= edges.max()
    wraps to the last row/column through numpy's negative indexing at :344, so
    row = (H - 1 - cnt_h) mod H, col = (W - 1 - cnt_w) mod W.
    Returns (row int64[N], col int64[N], (theta_min, theta_max)).
    """
    phi, theta = to_deflection_coordinates(pc[:, 0], pc[:, 1], pc[:, 2])
    if theta_range is None:
        theta_min, theta_max = theta.min(), theta.max()
    else:
        theta_min, theta_max = theta_range
    edges_h = np.linspace(theta_min, theta_max, height)
    edges_w = np.linspace(-np.pi, np.pi, width)
    cnt_h = np.searchsorted(edges_h, theta, side="right")
    cnt_w = np.searchsorted(edges_w, phi, side="right")
    row = (height - 1 - cnt_h) % height
    col = (width - 1 - cnt_w) % width
    return row.astype(np.int64), col.astype(np.int64), (theta_min, theta_max)


def depth_test(r, pix, n_pixels, farthest_wins=False):
    """Winning point per pixel: nearest range (farthest if `farthest_wins`), -1 where empty.

    The reference's winner among EQUAL ranges is whatever numpy's unstable
    argsort leaves last (src/dataset/utils.py:300); the contract here, and of
    the CUDA path, is lowest point index.  Tests keep ranges unique per pixel
    or compare pixel contents.
    """
    n = r.shape[0]
    key = -r if farthest_wins else r
    order = np.lexsort((np.arange(n), key))       # by key, then by index
    first = np.full(n_pixels, -1, dtype=np.int64)
    # reversed stable assignment: the first element in `order` per pixel must win
    first[pix[order[::-1]]] = order[::-1]
    return first


def spherical_projection(pc, height=64, width=2048, theta_range=None, th=1.0,
                         sort_largest_first=False, bins_h=None, max_range=None):
    """Same return tuple as src/dataset/utils.py:288-349.

    `th` and `max_range` are dead parameters in the reference; `bins_h`
    (caller-provided descending row edges) replaces the linspace at :330-331.
    """
    pc = np.asarray(pc)
    r = point_range(pc)
    phi, theta = to_deflection_coordinates(pc[:, 0], pc[:, 1], pc[:, 2])
    if theta_range is None:
        theta_min, theta_max = theta.min(), theta.max()
    else:
        theta_min, theta_max = theta_range
    if bins_h is None:
        bins_h = np.linspace(theta_min, theta_max, height)[::-1]
    bins_w = np.linspace(-np.pi, np.pi, width)[::-1]
    idx_h = np.digitize(theta, bins_h) - 1
    idx_w = np.digitize(phi, bins_w) - 1
    pix = (idx_h % height) * width + (idx_w % width)
    win = depth_test(r, pix, height * width, farthest_wins=sort_largest_first)
    pj_img = np.zeros((height * width, pc.shape[1]), dtype=np.float32)
    occ = win >= 0
    pj_img[occ] = pc[win[occ]]
    pj_img = pj_img.reshape(height, width, pc.shape[1])
    theta_img = np.stack(width * [bins_h], axis=-1)
    phi_img = np.stack(height * [bins_w], axis=0)
    alpha = np.sqrt(np.square(theta_img) + np.square(phi_img))
    return pj_img, alpha, (theta_min, theta_max), (-np.pi, np.pi)


def backproject_labels(label_img, row, col):
    """Definition of SURVEY.md 8a-2: every point reads the label of the pixel it projects to."""
    return np.asarray(label_img)[row, col]


def kitti_frame(xyzi, raw_label, height, width, lut, theta_range=None, flip=False):
    """Loader glue of src/dataset/dataloader_semantic_KITTI.py:35-97 without resize/normals.

    xyzi float32 [N,4], raw_label uint32 [N], lut int32 raw-id -> train-id.
    Returns dict with the reference's tensors as numpy arrays:
      xyz [3,H,W] f32, range [1,H,W] f32, reflectivity [1,H,W] f32, semantics [1,H,W] int64,
    plus the hidden intermediates pix int64[N] and winner int64[H*W].
    """
    sem = lut[(raw_label & 0xFFFF).astype(np.int64)].astype(np.int64)           # :40-47
    pc = np.concatenate([xyzi, sem[..., np.newaxis]], axis=-1)                   # :49 -> float64
    img, _, (tmin, tmax), _ = spherical_projection(pc, height, width, theta_range)
    row, col, _ = projection_indices(pc, height, width, theta_range)
    pix = row * width + col
    win = depth_test(point_range(pc), pix, height * width)
    if flip:                                                                     # :71-73
        img = img[:, ::-1, :].copy()
        img[..., 1] *= -1
    xyz_img = img[..., 0:3]
    rng_img = np.linalg.norm(xyz_img, axis=-1)                                   # :83 (float32)
    return {
        "xyz": np.ascontiguousarray(xyz_img.transpose(2, 0, 1)).astype(np.float32),
        "range": rng_img[None].astype(np.float32),
        "reflectivity": np.ascontiguousarray(img[..., 3][None]).astype(np.float32),
        "semantics": img[..., 4][None].astype(np.int64),
        "pix": pix, "winner": win, "theta_min": float(tmin), "theta_max": float(tmax),
    }


def rotate_z(xyz, angle_deg):
    """src/dataset/utils.py:4-18: xyz @ [[c,-s,0],[s,c,0],[0,0,1]] in float64."""
    a = np.radians(angle_deg)
    rot = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    return np.dot(xyz, rot)


def build_normal_xyz(xyz, norm_factor=0.25):
    """src/dataset/utils.py:30-59 (six 3x3 Scharr filters via OpenCV, cross product, normalise)."""
    import cv2
    x, y, z = (xyz[..., i].astype(np.float32) for i in range(3))
    sc = 1.0 / norm_factor
    Sxx = cv2.Scharr(x, cv2.CV_32FC1, 1, 0, scale=sc); Sxy = cv2.Scharr(x, cv2.CV_32FC1, 0, 1, scale=sc)
    Syx = cv2.Scharr(y, cv2.CV_32FC1, 1, 0, scale=sc); Syy = cv2.Scharr(y, cv2.CV_32FC1, 0, 1, scale=sc)
    Szx = cv2.Scharr(z, cv2.CV_32FC1, 1, 0, scale=sc); Szy = cv2.Scharr(z, cv2.CV_32FC1, 0, 1, scale=sc)
    normal = -np.dstack((Syx * Szy - Szx * Syy, Szx * Sxy - Szy * Sxx, Sxx * Syy - Syx * Sxy))
    n = np.linalg.norm(normal, axis=2) + 1e-10
    normal[:, :, 0] /= n
    normal[:, :, 1] /= n
    normal[:, :, 2] /= n
    return normal


def kitti_item(xyzi, raw_label, lut, projection=(64, 2048), resize=True, flip=False, yaw_deg=None,
               theta_range=None, normalise_reflectivity=False, drop_empty_rows=False, resize_to=(2048, 128)):
    """SemanticKitti.__getitem__ (src/dataset/dataloader_semantic_KITTI.py:31-99) with the random draws
    (yaw angle :53, flip coin :71) passed in.  Returns the five arrays in the reference's order/dtypes.
    theta_range / normalise_reflectivity give SemanticCUDAL.__getitem__ (dataloader_semantic_CUDAL.py:70-125);
    drop_empty_rows / resize_to=(1024, 64) give SemanticWADS.__getitem__ (dataloader_semantic_WADS.py:96-155)."""
    import cv2
    sem = lut[(raw_label & 0xFFFF).astype(np.int64)].astype(np.int64)
    pc = np.concatenate([xyzi, sem[..., np.newaxis]], axis=-1)
    if yaw_deg is not None:
        pc[..., 0:3] = rotate_z(pc[..., 0:3].reshape(-1, 3), float(yaw_deg))
    img, _, _, _ = spherical_projection(pc, projection[0], projection[1], theta_range=theta_range)
    if drop_empty_rows:
        img = img[~np.all(np.linalg.norm(img, axis=-1) == 0, axis=1)]            # WADS :125
    if resize:
        img = cv2.resize(img, tuple(resize_to), interpolation=cv2.INTER_NEAREST)
    if flip:
        img = img[:, ::-1, :]
        img[..., 1] *= -1
    label_img = img[..., 4:5]
    refl = img[..., 3]
    if normalise_reflectivity:
        refl = refl / np.maximum(refl.max(), 1.0)                                 # CUDAL :107
    xyz = img[..., 0:3]
    rng = np.linalg.norm(xyz, axis=-1)
    normals = build_normal_xyz(xyz[..., 0:3])
    return (rng[..., None].transpose(2, 0, 1).astype("float32"), refl[..., None].transpose(2, 0, 1).astype("float32"),
            xyz.transpose(2, 0, 1).astype("float32"), normals.transpose(2, 0, 1).astype("float32"),
            label_img.transpose(2, 0, 1).astype("int64"))


def thab_item(xyzi, raw_label, lut, flip=False, yaw_deg=None, H=128, W=2048):
    """SemanticTHAB.__getitem__ (src/dataset/dataloader_semantic_THAB.py:28-84): organised cloud, flip before
    the yaw, the yaw as np.roll by round(angle / (2 pi) * W) columns plus rotate_z, float64 range."""
    sem = lut[(raw_label & 0xFFFF).astype(np.int64)].astype(np.int64).reshape(H, W, 1)
    img = np.concatenate((xyzi.reshape(H, W, 4), sem), axis=-1)                  # float64
    if flip:
        img = img[:, ::-1, :]
        img[..., 1] = -img[..., 1]
    if yaw_deg is not None:
        img = np.roll(img, int(round((yaw_deg / (2 * np.pi)) * W)), axis=1)
        img[..., 0:3] = rotate_z(img[..., 0:3].reshape(-1, 3), yaw_deg).reshape(img[..., 0:3].shape)
    xyz = img[..., 0:3]
    rng = np.linalg.norm(xyz, axis=-1)
    normals = build_normal_xyz(xyz)
    return (rng[..., None].transpose(2, 0, 1).astype("float32"), img[..., 3][..., None].transpose(2, 0, 1).astype("float32"),
            xyz.transpose(2, 0, 1).astype("float32"), normals.transpose(2, 0, 1).astype("float32"),
            img[..., 4:5].transpose(2, 0, 1).astype("int64"))
