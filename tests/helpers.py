"""Test-only helpers (may use the oracle)."""
import numpy as np
import torch

from oracle import metrics as om
from oracle import uncertainty as ou


def enforce_margins(mc_logits, labels, n_bins=15, ignore_index=0, top2=1e-3, edge=1e-4, conf_renorm=True):
    """Make argmax and bin membership insensitive to ulp-level differences between implementations.

    Pixels whose top-2 gap of p_bar is below `top2` get their arg-max logit boosted in every sample;
    pixels whose confidence is within `edge` of a bin edge get label = ignore_index (so both sides
    drop them from the ECE bins).  Returns (logits, labels) modified copies.
    """
    x = mc_logits.clone()
    lab = labels.clone()
    for _ in range(8):
        r = ou.mc_reduce(x.double())
        top = r["p_bar"].topk(2, dim=1)
        bad = (top.values[:, 0] - top.values[:, 1]) < top2          # [B,H,W]
        if not bad.any():
            break
        idx = top.indices[:, 0]                                      # [B,H,W]
        boost = torch.zeros_like(x[0])
        boost.scatter_(1, idx.unsqueeze(1), bad.unsqueeze(1).to(x.dtype) * 1.0)
        x = x + boost.unsqueeze(0)
    r = ou.mc_reduce(x.double())
    p = r["p_bar"]
    if conf_renorm:
        p = om.to_probs(p, "probs")
    conf = p.max(dim=1).values.numpy()
    edges = om.ece_edges(n_bins).astype(np.float64)
    near = (np.abs(conf[..., None] - edges[None, None, None, :]).min(axis=-1) < edge)
    lab[torch.from_numpy(near)] = ignore_index
    return x, lab


def rel_close(a, b, rtol, atol):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    return float(err.max()) <= 0.0, float(np.abs(a - b).max()), float((np.abs(a - b) / np.maximum(np.abs(b), 1e-30)).max())


def normals_condition_mask(xyz_chw, thresh=1e-2):
    """Pixels where build_normal_xyz is well conditioned.

    The normal is the normalised cross product of the Scharr d/dx and d/dy vectors.  Where those two
    vectors are (nearly) parallel -- e.g. the block corners of a nearest-neighbour upsampled image --
    the cross product is pure float32 rounding noise and its normalised direction is arbitrary in ANY
    implementation (cv2 itself changes with its SIMD dispatch).  mask = |gx x gy| > thresh * |gx| |gy|,
    or both gradients zero (normal exactly 0)."""
    import cv2
    g = []
    for c in range(3):
        p = np.ascontiguousarray(xyz_chw[c], dtype=np.float32)
        g.append((cv2.Scharr(p, cv2.CV_32FC1, 1, 0, scale=4.0).astype(np.float64),
                  cv2.Scharr(p, cv2.CV_32FC1, 0, 1, scale=4.0).astype(np.float64)))
    gx = np.stack([a for a, _ in g], -1)
    gy = np.stack([b for _, b in g], -1)
    cr = np.linalg.norm(np.cross(gx, gy), axis=-1)
    den = np.linalg.norm(gx, axis=-1) * np.linalg.norm(gy, axis=-1)
    return (cr > thresh * den) | (den == 0)
