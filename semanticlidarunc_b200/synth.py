"""Seeded synthetic inputs of the shapes BASELINE.json names.

Datasets are not available offline, so tests and bench.py use these generators.
Host-side numpy/torch only; nothing here is on the timed path.
"""
from __future__ import annotations

import numpy as np
import torch

from .dataset.definitions import id_map

# sensor presets: (n_beams, n_azimuth, fov_up_deg, fov_down_deg, image H, image W)
SENSORS = {
    "hdl64": (64, 1875, 2.0, -24.8, 64, 2048),     # 120 000 points -> 64 x 2048
    "os1-128": (128, 2048, 22.5, -22.5, 128, 2048),  # 262 144 points -> 128 x 2048
    "tiny": (16, 200, 10.0, -20.0, 16, 256),        # 3 200 points  -> 16 x 256 (oracle-speed cases)
}


def synth_scan(seed: int, sensor: str = "hdl64", n_points: int | None = None):
    """One beam-structured scan in KITTI on-disk layout.

    Returns (xyzi float32 [N,4], raw_label uint32 [N]).  raw_label carries a
    random instance id in the upper 16 bits, as SemanticKITTI .label files do
    (reference: src/dataset/dataloader_semantic_KITTI.py:37-44).
    No (0,0,0) points and no exact duplicate points are emitted.
    """
    n_beams, n_az, up, down, _, _ = SENSORS[sensor]
    rng = np.random.default_rng(seed)
    elev = np.deg2rad(np.linspace(up, down, n_beams))
    azim = np.linspace(-np.pi, np.pi, n_az, endpoint=False)
    el = np.repeat(elev, n_az) + rng.normal(0.0, 2e-4, n_beams * n_az)
    az = np.tile(azim, n_beams) + rng.normal(0.0, 2e-4, n_beams * n_az)
    az = np.clip(az, -np.pi + 1e-6, np.pi - 1e-6)
    rg = rng.uniform(2.0, 80.0, n_beams * n_az)
    x = rg * np.cos(el) * np.cos(az)
    y = rg * np.cos(el) * np.sin(az)
    z = rg * np.sin(el)
    inten = rng.uniform(0.0, 1.0, n_beams * n_az)
    xyzi = np.stack([x, y, z, inten], axis=1).astype(np.float32)
    keys = np.fromiter(id_map.keys(), dtype=np.uint32)
    sem = keys[rng.integers(0, len(keys), n_beams * n_az)]
    inst = rng.integers(0, 1 << 16, n_beams * n_az).astype(np.uint32)
    raw = (sem | (inst << np.uint32(16))).astype(np.uint32)
    if n_points is not None and n_points < xyzi.shape[0]:
        keep = np.sort(rng.choice(xyzi.shape[0], size=n_points, replace=False))
        xyzi, raw = xyzi[keep], raw[keep]
    return np.ascontiguousarray(xyzi), np.ascontiguousarray(raw)


def synth_mc_logits(seed: int, T: int, B: int, C: int, H: int, W: int, device="cpu", scale: float = 3.0):
    """MC-dropout logits [T,B,C,H,W] fp32 = randn * scale, and labels [B,H,W] int64 in [0,C)."""
    g = torch.Generator(device=device).manual_seed(seed)
    logits = torch.randn((T, B, C, H, W), generator=g, device=device, dtype=torch.float32) * scale
    labels = torch.randint(0, C, (B, H, W), generator=g, device=device, dtype=torch.int64)
    return logits, labels


def synth_evidential_logits(seed: int, B: int, C: int, H: int, W: int, device="cpu", scale: float = 3.0):
    """Evidential head output [B,C+1,H,W] (C shape logits + 1 scale logit) and labels [B,H,W]."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = torch.randn((B, C + 1, H, W), generator=g, device=device, dtype=torch.float32) * scale
    labels = torch.randint(0, C, (B, H, W), generator=g, device=device, dtype=torch.int64)
    return out, labels


def synth_coherent_labels(seed: int, B: int, C: int, H: int, W: int, block: int = 32, device="cpu"):
    """Spatially coherent label maps (block x block patches of one class), as real scans are."""
    g = torch.Generator(device=device).manual_seed(seed)
    hb, wb = (H + block - 1) // block, (W + block - 1) // block
    coarse = torch.randint(0, C, (B, hb, wb), generator=g, device=device, dtype=torch.int64)
    return coarse.repeat_interleave(block, 1).repeat_interleave(block, 2)[:, :H, :W].contiguous()
