// slu_special.cuh -- fp32 digamma / trigamma for positive arguments (evidential kernels).
//
// The reference calls torch.special.digamma / torch.digamma (src/models/probability_helper.py:128,
// src/losses/regularizers.py:335-337); ATen's float kernel is the cephes scheme: recurrence
// psi(x) = psi(x+1) - 1/x up to x >= 10, then the asymptotic series.  Same scheme here with the
// recurrence stopped at x >= 6 (the series' first dropped term is then 691/(32760 x^12) < 2e-11),
// and the whole evaluation kept in fp32.  Parity is by tolerance (1e-5 relative on the reduced maps).
#pragma once
#include <cuda_runtime.h>

namespace slu {

// psi(x), x > 0
__device__ __forceinline__ float digamma_pos(float x) {
    float r = 0.f;
#pragma unroll 1
    while (x < 2.f) {                 // rare in this code base: arguments are alpha + 1 >= 2
        r -= __frcp_rn(x);
        x += 1.f;
    }
    if (x < 6.f) {
        // four recurrence steps at once: 1/x + 1/(x+1) + 1/(x+2) + 1/(x+3) = (2x+3)(2u+2) / (u(u+2)), u = x(x+3)
        const float u = x * (x + 3.f);
        r -= __fdividef(fmaf(2.f, x, 3.f) * fmaf(2.f, u, 2.f), u * (u + 2.f));
        x += 4.f;
    }
    const float inv = __frcp_rn(x);
    const float z = inv * inv;
    // 1/12 z - 1/120 z^2 + 1/252 z^3 - 1/240 z^4 + 1/132 z^5
    float y = fmaf(z, 7.57575757575757576e-3f, -4.16666666666666667e-3f);
    y = fmaf(z, y, 3.96825396825396825e-3f);
    y = fmaf(z, y, -8.33333333333333333e-3f);
    y = fmaf(z, y, 8.33333333333333333e-2f);
    y *= z;
    return r + (logf(x) - 0.5f * inv - y);
}

// psi'(x), x > 0
__device__ __forceinline__ float trigamma_pos(float x) {
    float r = 0.f;
#pragma unroll 1
    while (x < 6.f) {
        const float i = __frcp_rn(x);
        r = fmaf(i, i, r);
        x += 1.f;
    }
    const float inv = __frcp_rn(x);
    const float z = inv * inv;
    // 1/x + 1/(2x^2) + 1/(6x^3) - 1/(30x^5) + 1/(42x^7) - 1/(30x^9) + 5/(66 x^11)
    float y = fmaf(z, 7.57575757575757576e-2f, -3.33333333333333333e-2f);
    y = fmaf(z, y, 2.38095238095238095e-2f);
    y = fmaf(z, y, -3.33333333333333333e-2f);
    y = fmaf(z, y, 1.66666666666666667e-1f);
    y = fmaf(y, z * inv, fmaf(0.5f, z, inv));
    return r + y;
}

// lgamma(a), psi(a), psi'(a) together, a > 0 (the KL regulariser needs all three per off-class), branch-free for
// a >= 1 (every Dirichlet concentration the heads produce): with w = 1/a and ln a from one lg2.approx,
//   lgamma(a) = (a - 1/2) ln a - a + ln(2 pi)/2 + w r(w)         r(0) = 1/12
//   psi(a)    = ln a + w (w h(w) - 1/2)                           h(0) = -1/12   (psi_g above)
//   psi'(a)   = w + w^2/2 + w^3 v(w)                              v(0) = 1/6
// r (degree 6), h and v (degree 7) are least-squares fits on Chebyshev nodes of w in [0,1] (tools/fit_digamma.py):
// lgamma to 1.2e-8 * max(1,|.|) before fp32 assembly, psi to 9e-8 absolute, psi' to 1.6e-7 relative.  Two MUFU
// (rcp, lg2) and ~30 FMA per call; a < 1 (clamped concentrations only) takes one recurrence step first.
// psi' (the only one gradients use) is good to 3e-7 relative; lgamma and psi enter the KL value only, next to the
// constant lgamma(C) ~ 39, and are held to 1e-5 * max(1,|.|) in the tests.
struct LDT { float lg, psi, tri; };

__device__ __forceinline__ float rcp_fast(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_fast_(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---- digamma as ln(x) + g(1/x), x >= 1 (the evaluation kernels: arguments are alpha + 1) --------------------
//   g(w) = psi(1/w) + ln(w) = w (w h(w) - 1/2),  h(0) = -1/12
// g is smooth on w in [0, 1]; h is a degree-7 least-squares polynomial on Chebyshev nodes (fit in 40-digit arithmetic,
// tools/fit_digamma.py), |g_fp32 - g| <= 9e-8 for every x in [1, 1e7] -- one MUFU.RCP and 9 FMA, no recurrence, no branch,
// so the C evaluations of a pixel interleave freely.  Callers that need psi(a) - psi(b) take ONE logarithm of a / b
// (lg2.approx: absolute error 2^-22 near 1), which is cheaper and, for the dominant class where the two arguments are
// close, more accurate than subtracting two rounded digammas.
struct PsiG { float w, g; };

__device__ __forceinline__ PsiG psi_g(float x) {
    PsiG r;
    r.w = rcp_fast(x);
    float h = 2.889277183e-04f;
    h = fmaf(h, r.w, -1.886666441e-03f);
    h = fmaf(h, r.w, 4.937323876e-03f);
    h = fmaf(h, r.w, -5.935221128e-03f);
    h = fmaf(h, r.w, 4.237475471e-04f);
    h = fmaf(h, r.w, 8.287647461e-03f);
    h = fmaf(h, r.w, 1.917667073e-06f);
    h = fmaf(h, r.w, -8.333334680e-02f);
    r.g = r.w * fmaf(h, r.w, -0.5f);
    return r;
}

// psi(xa) - psi(xb) with a = psi_g(xa), b = psi_g(xb)
__device__ __forceinline__ float psi_diff(float xa, const PsiG& a, const PsiG& b) {
    return fmaf(lg2_fast_(xa * b.w), 0.6931471805599453f, a.g - b.g);
}

__device__ __forceinline__ LDT ldt_pos(float a) {
    float x = a, lg_fix = 0.f, psi_fix = 0.f, tri_fix = 0.f;
    if (a < 1.0f) {                                          // rare: one recurrence step up
        const float ia = rcp_fast(a);
        lg_fix = -lg2_fast_(a) * 0.6931471805599453f;        // lgamma(a) = lgamma(a+1) - ln a
        psi_fix = -ia;                                       // psi(a)    = psi(a+1) - 1/a
        tri_fix = ia * ia;                                   // psi'(a)   = psi'(a+1) + 1/a^2
        x = a + 1.0f;
    }
    const float w = rcp_fast(x);
    const float lnx = lg2_fast_(x) * 0.6931471805599453f;
    float r = 1.383177419e-04f;
    r = fmaf(r, w, -6.336787229e-04f); r = fmaf(r, w, 1.049638091e-03f); r = fmaf(r, w, -5.343459393e-05f);
    r = fmaf(r, w, -2.772524576e-03f); r = fmaf(r, w, -1.822395578e-07f); r = fmaf(r, w, 8.333333419e-02f);
    float h = 2.889277183e-04f;
    h = fmaf(h, w, -1.886666441e-03f); h = fmaf(h, w, 4.937323876e-03f); h = fmaf(h, w, -5.935221128e-03f);
    h = fmaf(h, w, 4.237475471e-04f);  h = fmaf(h, w, 8.287647461e-03f); h = fmaf(h, w, 1.917667073e-06f);
    h = fmaf(h, w, -8.333334680e-02f);
    float v = -3.238177565e-03f;
    v = fmaf(v, w, 1.716627286e-02f);  v = fmaf(v, w, -3.719230805e-02f); v = fmaf(v, w, 3.725572734e-02f);
    v = fmaf(v, w, -2.641956099e-03f); v = fmaf(v, w, -3.307210709e-02f); v = fmaf(v, w, -1.011315425e-05f);
    v = fmaf(v, w, 1.666667325e-01f);
    LDT o;
    o.lg = (fmaf(x - 0.5f, lnx, -x) + 0.918938533204672742f) + fmaf(w, r, lg_fix);
    o.psi = lnx + fmaf(w, fmaf(h, w, -0.5f), psi_fix);
    const float w2 = w * w;
    o.tri = fmaf(w2 * w, v, fmaf(0.5f, w2, w)) + tri_fix;
    return o;
}

}  // namespace slu
