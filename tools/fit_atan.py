#!/usr/bin/env python
"""Coefficients and error certificate of the fp32 arctangent used by the projection prefilter (csrc/slu_project.cu::fast_atan2).

atan(t) = t * P(t^2) on [0, 1], P of degree 6 (least squares on Chebyshev nodes, refined towards minimax by reweighting),
then the usual octant folding.  The prefilter only needs |error| well below ANGLE_MARGIN = 8e-6 rad; this prints the
maximum error of the fp32 evaluation (Horner in float32, division replaced by a 2-ulp perturbed quotient) against
float64 atan2 over a dense sweep.  CPU only."""
import numpy as np
from numpy.polynomial import chebyshev as Ch, Polynomial

DEG = 6
n = 4000
nodes = np.cos(np.pi * (np.arange(n) + 0.5) / n)          # on [-1,1] -> s in [0,1]
s = (nodes + 1) / 2
t = np.sqrt(s)
f = np.where(t > 0, np.arctan(t) / np.maximum(t, 1e-300), 1.0)
w = np.ones_like(s)
for _ in range(40):                                         # Lawson-style reweighting towards the minimax fit
    cf = Ch.chebfit(nodes, f, DEG, w=w)
    err = np.abs(Ch.chebval(nodes, cf) - f) * t
    w = w * (1 + 4 * err / err.max())
coef = Polynomial(Ch.cheb2poly(cf))(Polynomial([-1, 2.0])).coef   # in s, lowest order first
print("P(s) coefficients, lowest order first:")
print(", ".join("%.9ef" % c for c in coef))

f32 = np.float32
c32 = [f32(c) for c in coef]


def fast_atan2(y, x, perturb=0):
    y = y.astype(f32); x = x.astype(f32)
    ax, ay = np.abs(x), np.abs(y)
    mx, mn = np.maximum(ax, ay), np.minimum(ax, ay)
    q = (mn / mx).astype(f32)
    if perturb:
        q = np.nextafter(q, f32(perturb * 10), dtype=f32)
        q = np.nextafter(q, f32(perturb * 10), dtype=f32)
    ss = (q * q).astype(f32)
    acc = np.full_like(ss, c32[-1])
    for c in c32[-2::-1]:
        acc = (acc * ss + c).astype(f32)
    r = (acc * q).astype(f32)
    r = np.where(ay > ax, (f32(1.5707963267948966) - r).astype(f32), r)
    r = np.where(x < 0, (f32(3.14159265358979) - r).astype(f32), r)
    return np.where(y < 0, -r, r).astype(f32)


rng = np.random.default_rng(0)
worst = 0.0
for perturb in (0, 1, -1):
    ang = np.concatenate([np.linspace(-np.pi, np.pi, 4_000_001), rng.uniform(-np.pi, np.pi, 2_000_000)])
    rad = 10 ** rng.uniform(-2, 2.5, ang.size)
    x = (rad * np.cos(ang)).astype(f32); y = (rad * np.sin(ang)).astype(f32)
    e = np.abs(fast_atan2(y, x, perturb).astype(np.float64) - np.arctan2(y.astype(np.float64), x.astype(np.float64)))
    e = np.minimum(e, 2 * np.pi - e)                        # the branch cut at +-pi
    worst = max(worst, e.max())
    print("perturb %+d: max |fast_atan2 - atan2| = %.3e rad" % (perturb, e.max()))
print("certificate: max error %.3e rad  (ANGLE_MARGIN 8.0e-6: %.1fx room)" % (worst, 8e-6 / worst))
