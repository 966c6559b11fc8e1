#!/usr/bin/env python
"""CPU emulation (numpy float64 = the device's IEEE arithmetic) of the cross-product decision the projection kernels can use
for points near a bin edge (csrc/slu_project.cu: col_count_by_cross / row_count_by_cross, SLU_PROJECT_CROSS).

For points placed at offsets of 1e-17 ... 1e-5 rad around column / row edges it compares
    cnt = k + [cross >= 0]      whenever |cross| > 1e-14 (|x| + |y|)   (k = the edge nearest to the fp32 angle)
with the reference's np.digitize(arctan2(...)) count and reports the disagreements among the DECIDED points (must be zero) and
how many points fall inside the band (they take the exact fp64 path on the device).  `--float32-inputs` rounds the coordinates to
float32 first (file coordinates); without it the coordinates are arbitrary float64 values (yaw-rotated points), which is what
probes the band itself.

    python tools/emulate_cross_decision.py [n_points] [--float32-inputs]"""
import sys

import numpy as np

PI, HALF_PI = np.pi, np.pi / 2
BAND = 1.0e-14


def columns(n, W=2048, f32=False, seed=0):
    rng = np.random.default_rng(seed)
    edges = np.linspace(-PI, PI, W)
    step = (PI - (-PI)) / (W - 1)
    na = (W + 31) // 32
    A = (32.0 * np.arange(na)) * step + (-PI)                    # the device's two-level table: dir(-pi + 32 a step), dir(b step)
    B = np.arange(32) * step
    cA, sA, cB, sB = np.cos(A), np.sin(A), np.cos(B), np.sin(B)
    k = rng.integers(1, W - 1, n)
    off = 10.0 ** rng.uniform(-17, -5, n) * rng.choice([-1, 1], n)
    ang, r = edges[k] + off, rng.uniform(0.5, 100, n)
    x, y = r * np.cos(ang), r * np.sin(ang)
    if f32:
        x, y = x.astype(np.float32).astype(np.float64), y.astype(np.float32).astype(np.float64)
    phi = np.arctan2(y, x)
    ref = np.searchsorted(edges, phi, side="right")               # #{edges <= phi} = what digitize counts
    phi32 = phi + rng.uniform(-8.5e-7, 8.5e-7, n)                  # the prefilter's angle: true angle +- its error bound
    kk = np.rint((phi32 + PI) / step).astype(np.int64)
    a, b = kk // 32, kk % 32
    C, S = cA[a] * cB[b] - sA[a] * sB[b], sA[a] * cB[b] + cA[a] * sB[b]
    cross = y * C - x * S
    decided = (np.abs(cross) > BAND * (np.abs(x) + np.abs(y))) & (kk >= 1) & (kk <= W - 2)
    return int((decided & (kk + (cross >= 0) != ref)).sum()), int((~decided).sum())


def rows(n, H=64, f32=False, seed=1):
    rng = np.random.default_rng(seed)
    lo, hi = -0.4363323129985824 + 1e-3 * rng.standard_normal(), 0.03490658503988659 + 1e-3 * rng.standard_normal()
    edges = np.linspace(lo, hi, H)
    g = HALF_PI - edges
    Cg, Sg = np.cos(g), np.sin(g)
    j = rng.integers(1, H - 1, n)
    off = 10.0 ** rng.uniform(-17, -5, n) * rng.choice([-1, 1], n)
    th, az, r = edges[j] + off, rng.uniform(-PI, PI, n), rng.uniform(0.5, 100, n)
    x, y, z = r * np.cos(th) * np.cos(az), r * np.cos(th) * np.sin(az), r * np.sin(th)
    if f32:
        x, y, z = (v.astype(np.float32).astype(np.float64) for v in (x, y, z))
    rho = np.sqrt(x ** 2 + y ** 2)
    theta = -np.arctan2(rho, z) + HALF_PI
    ref = np.searchsorted(edges, theta, side="right")
    th32 = theta + rng.uniform(-8.5e-7, 8.5e-7, n)
    jj = np.clip(np.rint((th32 - lo) / ((hi - lo) / (H - 1))).astype(np.int64), 0, H - 1)
    cross = z * Sg[jj] - rho * Cg[jj]
    decided = (np.abs(cross) > BAND * (rho + np.abs(z))) & (jj >= 1) & (jj <= H - 2)
    return int((decided & (jj + (cross >= 0) != ref)).sum()), int((~decided).sum())


if __name__ == "__main__":
    n = int(next((a for a in sys.argv[1:] if a.isdigit()), 4_000_000))
    f32 = "--float32-inputs" in sys.argv
    for name, fn in (("columns", columns), ("rows", rows)):
        bad, band = fn(n, f32=f32)
        print(f"{name}: {n} points, {bad} disagreements among the decided ones, {band} inside the band (exact path)")
