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


def project_device(pc: torch.Tensor, height=64, width=2048, theta_range=None, sort_largest_first=False):
    """pc [N,Cin] CUDA tensor -> dict(img [H,W,Cin] f32, pix [N] i32, winner [H,W] i32, theta [1,2] f64, diag)."""
    return ops.project_points(pc.to(torch.float64), height, width, theta_range=theta_range,
                              farthest_wins=sort_largest_first)


def spherical_projection(pc, height=64, width=2048, theta_range=None, th=1.0, sort_largest_first=False,
                         bins_h=None, max_range=None, *, return_alpha=True):
    """Drop-in for the reference function: returns (pj_img, alpha, (theta_min, theta_max), (phi_min, phi_max)).

    pc: array [N,Cin] (x,y,z first).  The reference computes in pc's dtype, float64 in every loader
    (np.concatenate of float32 xyzi with int64 labels); this path always computes in float64.
    `th` and `max_range` are dead parameters in the reference and are ignored here too.
    `bins_h` (caller-supplied row edges) is not supported on the device path.
    `alpha` [H,W] float64 depends only on the bin edges; pass return_alpha=False to skip building it
    (every reference caller discards it).
    """
    if bins_h is not None:
        raise NotImplementedError("caller-supplied bins_h is not supported by the CUDA projection")
    dev = _lib.require_cuda()
    pc_t = torch.as_tensor(np.ascontiguousarray(pc, dtype=np.float64)).to(dev, non_blocking=True)
    res = project_device(pc_t, height, width, theta_range, sort_largest_first)
    pj_img = res["img"].cpu().numpy()
    tmin, tmax = (float(v) for v in res["theta"][0].cpu())
    if theta_range is not None:
        tmin, tmax = theta_range
    alpha = None
    if return_alpha:
        bh = np.linspace(tmin, tmax, height)[::-1]
        bw = np.linspace(-np.pi, np.pi, width)[::-1]
        alpha = np.sqrt(np.square(bh)[:, None] + np.square(bw)[None, :])
    return pj_img, alpha, (tmin, tmax), (-np.pi, np.pi)
