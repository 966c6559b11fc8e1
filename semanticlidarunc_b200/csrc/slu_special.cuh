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

// lgamma(a), psi(a), psi'(a) together, a > 0 (the KL regulariser needs all three per off-class).
//   1 <= a < 1.25 (a class without evidence: the common case): Taylor series about 1,
//       lgamma(1+d) = -g d + sum_{k>=2} (-1)^k zeta(k)/k d^k, and its two derivatives; truncation < 2e-8.
//   otherwise: one shared recurrence up to x >= 7 (product for lgamma, sum 1/x for psi, sum 1/x^2 for psi'),
//       one log of x shared by lgamma and psi, Stirling series in 1/x^2.
// lgamma and psi (value-only terms) are good to ~3e-6 * max(1, |.|); psi' (the gradient) to ~3e-7 relative.
// The KL term they feed carries the constant lgamma(C) ~ 39, and is held to 1e-5 relative in the tests.
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
    LDT r;
    if (a >= 1.0f && a < 1.25f) {
        const float d = a - 1.0f;
        float l = 0.083353840546109f;
        l = fmaf(l, d, -0.09095401714582904f); l = fmaf(l, d, 0.1000994575127818f);
        l = fmaf(l, d, -0.11133426586956469f); l = fmaf(l, d, 0.12550966952474304f);
        l = fmaf(l, d, -0.1440498967688461f);  l = fmaf(l, d, 0.1695571769974082f);
        l = fmaf(l, d, -0.20738555102867398f); l = fmaf(l, d, 0.27058080842778454f);
        l = fmaf(l, d, -0.40068563438653143f); l = fmaf(l, d, 0.8224670334241132f);
        r.lg = d * fmaf(l, d, -0.5772156649015329f);
        float q = 1.0000612481350588f;
        q = fmaf(q, d, -1.0001227133475785f); q = fmaf(q, d, 1.000246086553308f);
        q = fmaf(q, d, -1.0004941886041194f); q = fmaf(q, d, 1.000994575127818f);
        q = fmaf(q, d, -1.0020083928260821f); q = fmaf(q, d, 1.0040773561979444f);
        q = fmaf(q, d, -1.008349277381923f);  q = fmaf(q, d, 1.0173430619844492f);
        q = fmaf(q, d, -1.03692775514337f);   q = fmaf(q, d, 1.0823232337111381f);
        q = fmaf(q, d, -1.2020569031595942f); q = fmaf(q, d, 1.6449340668482264f);
        r.psi = fmaf(q, d, -0.5772156649015329f);
        float t = 15.00022923389113f;
        t = fmaf(t, d, -14.000428235308298f); t = fmaf(t, d, 13.000796225755764f);
        t = fmaf(t, d, -12.001472560170942f); t = fmaf(t, d, 11.002706952086388f);
        t = fmaf(t, d, -10.004941886041195f); t = fmaf(t, d, 9.008951176150363f);
        t = fmaf(t, d, -8.016067142608657f);  t = fmaf(t, d, 7.02854149338561f);
        t = fmaf(t, d, -6.050095664291537f);  t = fmaf(t, d, 5.086715309922246f);
        t = fmaf(t, d, -4.14771102057348f);   t = fmaf(t, d, 3.2469697011334144f);
        t = fmaf(t, d, -2.4041138063191885f); t = fmaf(t, d, 1.6449340668482264f);
        r.tri = t;
        return r;
    }
    // shared recurrence, at most six steps, predicated (no data-dependent trip count inside a warp)
    float x = a, prod = 1.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        if (x < 7.f) {
            const float i = rcp_fast(x);
            prod *= x;
            r1 += i;
            r2 = fmaf(i, i, r2);
            x += 1.f;
        }
    }
    // the two logs only enter the loss VALUE (gradients use psi' alone); lg2.approx (rel. error 2^-22) keeps
    // the KL value within 1e-6 relative, since the value carries the constant lgamma(C) ~ 39
    const float lnx = lg2_fast_(x) * 0.6931471805599453f;
    const float inv = rcp_fast(x);
    const float z = inv * inv;
    float sl = fmaf(z, -5.95238095238095238e-4f, 7.93650793650793651e-4f);      // 1/1260 - z/1680
    sl = fmaf(z, sl, -2.77777777777777778e-3f);                                   // -1/360
    sl = fmaf(z, sl, 8.33333333333333333e-2f);                                    // 1/12
    r.lg = (fmaf(x - 0.5f, lnx, -x) + 0.918938533204672742f + inv * sl) - lg2_fast_(prod) * 0.6931471805599453f;
    float sp = fmaf(z, 7.57575757575757576e-3f, -4.16666666666666667e-3f);
    sp = fmaf(z, sp, 3.96825396825396825e-3f);
    sp = fmaf(z, sp, -8.33333333333333333e-3f);
    sp = fmaf(z, sp, 8.33333333333333333e-2f);
    r.psi = (lnx - 0.5f * inv - z * sp) - r1;
    float st = fmaf(z, 7.57575757575757576e-2f, -3.33333333333333333e-2f);
    st = fmaf(z, st, 2.38095238095238095e-2f);
    st = fmaf(z, st, -3.33333333333333333e-2f);
    st = fmaf(z, st, 1.66666666666666667e-1f);
    r.tri = fmaf(st, z * inv, fmaf(0.5f, z, inv)) + r2;
    return r;
}

}  // namespace slu
