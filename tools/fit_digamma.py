#!/usr/bin/env python
"""Coefficients of h(w) in csrc/slu_special.cuh::psi_g:  psi(x) = ln x + w (w h(w) - 1/2), w = 1/x in [0, 1].
Least-squares fit on 400 Chebyshev nodes in 40-digit arithmetic (mpmath), then the fp32 Horner evaluation is checked
against scipy's float64 digamma over x in [1, 1e7].  CPU only; prints the coefficients lowest order first."""
import mpmath as mp
import numpy as np
from numpy.polynomial import Polynomial
from numpy.polynomial import chebyshev as Ch
from scipy.special import digamma

mp.mp.dps = 40
DEG = 7


def h(w):
    w = mp.mpf(w)
    return mp.mpf(-1) / 12 if w == 0 else (mp.digamma(1 / w) + mp.log(w) + w / 2) / w ** 2


nodes = np.cos(np.pi * (np.arange(400) + 0.5) / 400)
hv = np.array([float(h(float(x))) for x in (nodes + 1) / 2])
coef = Polynomial(Ch.cheb2poly(Ch.chebfit(nodes, hv, DEG)))(Polynomial([-1, 2.0])).coef
print(", ".join("%.9ef" % c for c in coef))
f = np.float32
x = np.concatenate([np.linspace(1, 8, 40001), np.logspace(0, 7, 40001)]).astype(f)
w = (f(1) / x).astype(f)
acc = np.full_like(w, f(coef[-1]))
for c in coef[-2::-1]:
    acc = (acc * w + f(c)).astype(f)
g = ((acc * w - f(0.5)) * w).astype(np.float64)
err = np.abs(np.log(x.astype(np.float64)) + g - digamma(x.astype(np.float64)))
print("max |psi_fp32 - psi| = %.3e at x = %.4f" % (err.max(), x[err.argmax()]))

# ---- lgamma and trigamma corrections used by ldt_pos:  lgamma(x) = (x-1/2) ln x - x + ln(2 pi)/2 + w r(w),
#      psi'(x) = w + w^2/2 + w^3 v(w)
from scipy.special import gammaln, polygamma  # noqa: E402

HL2PI = mp.log(2 * mp.pi) / 2


def r_fn(w):
    w = mp.mpf(w)
    if w == 0:
        return mp.mpf(1) / 12
    xx = 1 / w
    return (mp.loggamma(xx) - ((xx - mp.mpf(1) / 2) * mp.log(xx) - xx + HL2PI)) / w


def v_fn(w):
    w = mp.mpf(w)
    if w == 0:
        return mp.mpf(1) / 6
    return (mp.polygamma(1, 1 / w) - w - w * w / 2) / w ** 3


for name, fn, deg in (("r", r_fn, 6), ("v", v_fn, 7)):
    vals = np.array([float(fn(float(t))) for t in (nodes + 1) / 2])
    cf = Polynomial(Ch.cheb2poly(Ch.chebfit(nodes, vals, deg)))(Polynomial([-1, 2.0])).coef
    print(name, ", ".join("%.9ef" % c for c in cf))
    acc = np.full_like(w, f(cf[-1]))
    for c in cf[-2::-1]:
        acc = (acc * w + f(c)).astype(f)
    xd = x.astype(np.float64)
    if name == "r":
        lg = ((xd - 0.5) * np.log(xd) - xd + float(HL2PI)) + (w * acc).astype(np.float64)
        print("  max |lgamma err| / max(1,|lgamma|) = %.3e" % (np.abs(lg - gammaln(xd)) / np.maximum(1, np.abs(gammaln(xd)))).max())
    else:
        tri = (w + f(0.5) * w * w + w * w * w * acc).astype(np.float64)
        print("  max rel |trigamma err| = %.3e" % (np.abs(tri - polygamma(1, xd)) / polygamma(1, xd)).max())


# ---- merged value polynomial of the KL regulariser (round 2):  f(a) = -lgamma(a) + (a - 1) psi(a)
#      f(a) = -ln(a)/2 + a - ln(2 pi)/2 - 1/2 + w u(w),  u(w) = 1/2 - r(w) + (1 - w) h(w),  u(0) = 1/3
def u_fn(w):
    w = mp.mpf(w)
    if w == 0:
        return mp.mpf(1) / 3
    a = 1 / w
    fa = -mp.loggamma(a) + (a - 1) * mp.digamma(a)
    return (fa + mp.log(a) / 2 - a + HL2PI + mp.mpf(1) / 2) / w


for deg in (6, 7, 8):
    vals = np.array([float(u_fn(float(t))) for t in (nodes + 1) / 2])
    cf = Polynomial(Ch.cheb2poly(Ch.chebfit(nodes, vals, deg)))(Polynomial([-1, 2.0])).coef
    print("u deg", deg, ", ".join("%.9ef" % c for c in cf))
    acc = np.full_like(w, f(cf[-1]))
    for c in cf[-2::-1]:
        acc = (acc * w + f(c)).astype(f)
    xd = x.astype(np.float64)
    fa = -0.5 * np.log(xd) + xd - float(HL2PI) - 0.5 + (w * acc).astype(np.float64)
    ref = -gammaln(xd) + (xd - 1) * digamma(xd)
    print("  max |f err| = %.3e   (max |w u| err %.3e)" % (np.abs(fa - ref).max(),
          np.abs((w * acc).astype(np.float64) - (ref + 0.5 * np.log(xd) - xd + float(HL2PI) + 0.5)).max()))
