"""`spherical_projection` with the reference's signature (src/dataset/utils.py:288-349), on the GPU.

numpy in / numpy out, as the reference's DataLoader workers call it; the arithmetic runs in
libslu's projection kernels (csrc/slu_project.cu).  `project_device` is the same call without the
host round trip, and `ops.project_batch` the batched, loader-fused form.  `build_normal_xyz` (:30-59) is the loader
kernel's normal stage on its own.  `rotate_z` (:4-18) has no stand-alone counterpart: the yaw augmentation is applied
to the points inside the projection kernels (`ops.project_batch(yaw_deg=...)`, float64, the reference's matrix).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib, ops


def to_deflection_coordinates(x, y, z):
    """(phi, theta) of src/dataset/utils.py:61-67 for torch tensors (any device), float64 math."""
    x, y, z = (torch.as_tensor(v, dtype=torch.float64) for v in (x, y, z))
    p = torch.sqrt(x ** 2 + y ** 2)
    return torch.atan2(y, x), -torch.atan2(p, z) + np.pi / 2


def build_normal_xyz(xyz, norm_factor=0.25, ksize=3):
    """Surface normals of an organised cloud, `xyz` [h,w,3] -> [h,w,3] float32 (src/dataset/utils.py:30-59: six 3x3
    Scharr derivatives, cross product, normalisation), computed by libslu's loader kernel (csrc/slu_loader.cu).
    numpy in / numpy out like the reference; a CUDA tensor in gives a CUDA tensor out."""
    if ksize != 3:
        raise NotImplementedError("only the 3x3 Scharr stencil of the reference call sites is implemented")
    is_np = not torch.is_tensor(xyz)
    dev = _lib.require_cuda(xyz.device if (not is_np and xyz.is_cuda) else None)
    t = torch.as_tensor(np.ascontiguousarray(xyz, dtype=np.float32) if is_np else xyz).to(dev, torch.float32)
    if t.dim() != 3 or t.size(2) != 3:
        raise ValueError("xyz must be [h,w,3]")
    h, w = t.shape[:2]
    img = torch.zeros((1, 6, h, w), dtype=torch.float32, device=dev)
    img[0, :3] = t.permute(2, 0, 1)
    n = ops.frame_tensors(img, norm_factor=float(norm_factor))["normals"][0].permute(1, 2, 0).contiguous()
    return n.cpu().numpy() if is_np else n


_WS = {}          # device -> projection workspace, reused across calls (the drop-in is called once per scan)
_PIN = {}         # (H, W, Cin) -> pinned host image buffer + pinned meta buffer


def project_device(pc: torch.Tensor, height=64, width=2048, theta_range=None, sort_largest_first=False, bins_h=None):
    """pc [N,Cin] CUDA tensor -> dict(img [H,W,Cin] f32, pix [N] i32, winner [H,W] i32, theta [1,2] f64, diag)."""
    dev = pc.device
    res = ops.project_points(pc.to(torch.float64), height, width, theta_range=theta_range,
                             farthest_wins=sort_largest_first, workspace=_WS.get(dev), bins_h=bins_h)
    _WS[dev] = res["workspace"]
    return res


def _host_bins(pc64: np.ndarray, height: int, width: int, theta_range, bins_h=None):
    """Per-point pixel index in numpy, the reference's own arithmetic (src/dataset/utils.py:313-339: arctan2, linspace,
    digitize - 1 with the negative index wrapping round).  Only used to settle points the device flags as sitting within
    a few ulps of a bin edge."""
    x, y, z = pc64[:, 0], pc64[:, 1], pc64[:, 2]
    phi = np.arctan2(y, x)
    theta = -np.arctan2(np.sqrt(x ** 2 + y ** 2), z) + np.pi / 2
    tmin, tmax = (theta.min(), theta.max()) if theta_range is None else theta_range
    row = (np.digitize(theta, np.linspace(tmin, tmax, height)[::-1] if bins_h is None else np.asarray(bins_h)) - 1) % height
    col = (np.digitize(phi, np.linspace(-np.pi, np.pi, width)[::-1]) - 1) % width
    return row * width + col, (float(tmin), float(tmax))


def _settle_edge_points(pc64: np.ndarray, img: np.ndarray, pix: np.ndarray, height, width, theta_range, farthest_wins, bins_h=None):
    """Bit-exactness by construction: CUDA's and numpy's atan2 are both good to ~2 ulp, so a point within 4 ulp of a bin
    edge (the kernel counts them; ~1e-14 of all points) could land on either side.  When a scan has such points, every
    point's bin is recomputed with numpy and the pixels whose membership changed are re-resolved on the host.
    Returns (number of points that changed pixel, numpy's (theta_min, theta_max))."""
    host, trange = _host_bins(pc64, height, width, theta_range, bins_h)
    changed = np.nonzero(host != pix)[0]
    if changed.size == 0:
        return 0, trange
    r = np.sqrt(pc64[:, 0] ** 2 + pc64[:, 1] ** 2 + pc64[:, 2] ** 2)
    flat = img.reshape(height * width, -1)
    for q in np.unique(np.concatenate([host[changed], pix[changed].astype(np.int64)])):
        members = np.nonzero(host == q)[0]
        if members.size == 0:
            flat[q] = 0.0
        else:
            rr = r[members]
            best = members[rr == (rr.max() if farthest_wins else rr.min())].min()     # lowest index among exact ties
            flat[q] = pc64[best].astype(np.float32)
    pix[changed] = host[changed]
    return int(changed.size), trange


def spherical_projection(pc, height=64, width=2048, theta_range=None, th=1.0, sort_largest_first=False,
                         bins_h=None, max_range=None, *, return_alpha=True):
    """Drop-in for the reference function: returns (pj_img, alpha, (theta_min, theta_max), (phi_min, phi_max)).

    pc: array [N,Cin] (x,y,z first).  The reference computes in pc's dtype, float64 in every loader
    (np.concatenate of float32 xyzi with int64 labels); this path always computes in float64.
    `th` and `max_range` are dead parameters in the reference and are ignored here too.
    `bins_h`: caller-supplied row edges (H monotone values), used exactly as the reference uses them
    (np.digitize(theta, bins_h) - 1); the device bisects the same float64 array.
    `alpha` [H,W] float64 depends only on the bin edges and every reference caller discards it: pass
    return_alpha=False to skip building it (None is returned in its place).
    One H2D copy of the cloud, one D2H copy of the image into a reused pinned buffer, one synchronisation.  Points the
    kernel flags as within 4 ulp of a bin edge are settled with numpy's own arithmetic (`_settle_edge_points`); the
    count of the last call is left in `spherical_projection.last_near_edge` / `.last_settled`."""
    dev = _lib.require_cuda()
    pc64 = np.ascontiguousarray(pc, dtype=np.float64)
    pc_t = torch.from_numpy(pc64).to(dev, non_blocking=True)
    res = project_device(pc_t, height, width, theta_range, sort_largest_first, bins_h)
    key = (height, width, pc64.shape[1])
    if key not in _PIN:
        _PIN[key] = (torch.empty((height, width, pc64.shape[1]), dtype=torch.float32, pin_memory=True),
                     torch.empty(4, dtype=torch.float64, pin_memory=True))
    h_img, h_meta = _PIN[key]
    h_img.copy_(res["img"], non_blocking=True)
    meta = torch.cat([res["theta"].reshape(-1), res["diag"].reshape(-1).to(torch.float64)])
    h_meta.copy_(meta, non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()
    pj_img = h_img.numpy().copy()
    tmin, tmax, _, near_edge = (float(v) for v in h_meta)
    spherical_projection.last_near_edge, spherical_projection.last_settled = int(near_edge), 0
    if near_edge > 0:
        pix = res["pix"].cpu().numpy()
        spherical_projection.last_settled, (tmin, tmax) = _settle_edge_points(pc64, pj_img, pix, height, width, theta_range,
                                                                              sort_largest_first, bins_h)
    if theta_range is not None:
        tmin, tmax = theta_range
    alpha = None
    if return_alpha:
        bh = np.linspace(tmin, tmax, height)[::-1] if bins_h is None else np.asarray(bins_h, dtype=np.float64)
        bw = np.linspace(-np.pi, np.pi, width)[::-1]
        alpha = np.sqrt(np.square(bh)[:, None] + np.square(bw)[None, :])
    return pj_img, alpha, (tmin, tmax), (-np.pi, np.pi)


spherical_projection.last_near_edge = 0
spherical_projection.last_settled = 0
