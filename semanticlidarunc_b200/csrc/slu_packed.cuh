// slu_packed.cuh -- packed fp32 (f32x2) arithmetic for the FP-issue-bound kernels (sm_100a: FFMA2 / FMUL2 / FADD2).
//
// The evidential reduction and the loss kernels are bound by instruction issue, not by HBM: 45-61 % of their
// instructions are FFMA / FADD / FMUL (profiles/kernel_report_r01.md).  sm_100 issues TWO fp32 operations per lane with one
// FFMA2 / FMUL2 / FADD2 (PTX fma.rn.f32x2 ...), and the instruction takes a scalar immediate that is broadcast to both halves,
// so Horner polynomials cost no constant registers.  A thread therefore owns TWO adjacent pixels, held as float2: loads and
// stores become 8-byte accesses (still fully coalesced), every polynomial / softmax / chain-rule operation is one packed
// instruction, and only the MUFU evaluations (ex2, lg2, rcp), comparisons and selects stay per pixel.
// Each half is an IEEE fp32 fma / mul / add with round-to-nearest, i.e. bit-identical to the scalar instruction.
#pragma once
#include <cuda_runtime.h>
#include "slu_special.cuh"

namespace slu {

struct f2 {
    float2 v;
    __device__ __forceinline__ f2() {}
    __device__ __forceinline__ f2(float2 a) : v(a) {}
    __device__ __forceinline__ explicit f2(float a) : v(make_float2(a, a)) {}
    __device__ __forceinline__ f2(float a, float b) : v(make_float2(a, b)) {}
    __device__ __forceinline__ float& operator[](int i) { return i ? v.y : v.x; }
    __device__ __forceinline__ float operator[](int i) const { return i ? v.y : v.x; }
};

__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return f2(__ffma2_rn(a.v, b.v, c.v)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, float c) { return f2(__ffma2_rn(a.v, b.v, make_float2(c, c))); }
__device__ __forceinline__ f2 fma2(f2 a, float b, f2 c) { return f2(__ffma2_rn(a.v, make_float2(b, b), c.v)); }
__device__ __forceinline__ f2 fma2(f2 a, float b, float c) { return f2(__ffma2_rn(a.v, make_float2(b, b), make_float2(c, c))); }
__device__ __forceinline__ f2 operator*(f2 a, f2 b) { return f2(__fmul2_rn(a.v, b.v)); }
__device__ __forceinline__ f2 operator*(f2 a, float b) { return f2(__fmul2_rn(a.v, make_float2(b, b))); }
__device__ __forceinline__ f2 operator+(f2 a, f2 b) { return f2(__fadd2_rn(a.v, b.v)); }
__device__ __forceinline__ f2 operator+(f2 a, float b) { return f2(__fadd2_rn(a.v, make_float2(b, b))); }
__device__ __forceinline__ f2 operator-(f2 a) { return f2(make_float2(-a.v.x, -a.v.y)); }
__device__ __forceinline__ f2 operator-(f2 a, f2 b) { return f2(__fadd2_rn(a.v, make_float2(-b.v.x, -b.v.y))); }
__device__ __forceinline__ f2 operator-(f2 a, float b) { return f2(__fadd2_rn(a.v, make_float2(-b, -b))); }
__device__ __forceinline__ f2 operator-(float a, f2 b) { return f2(__fadd2_rn(make_float2(a, a), make_float2(-b.v.x, -b.v.y))); }
__device__ __forceinline__ f2& operator+=(f2& a, f2 b) { a = a + b; return a; }
__device__ __forceinline__ f2 max2(f2 a, f2 b) { return f2(fmaxf(a.v.x, b.v.x), fmaxf(a.v.y, b.v.y)); }
__device__ __forceinline__ f2 max2(f2 a, float b) { return f2(fmaxf(a.v.x, b), fmaxf(a.v.y, b)); }
// MUFU evaluations: one per half
__device__ __forceinline__ f2 rcp2(f2 a) { return f2(rcp_fast(a.v.x), rcp_fast(a.v.y)); }
__device__ __forceinline__ f2 lg2_2(f2 a) { return f2(lg2_fast_(a.v.x), lg2_fast_(a.v.y)); }
__device__ __forceinline__ f2 rcp_rn2(f2 a) { return f2(__frcp_rn(a.v.x), __frcp_rn(a.v.y)); }
__device__ __forceinline__ f2 ex2_2(f2 a) {
    f2 r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r.v.x) : "f"(a.v.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r.v.y) : "f"(a.v.y));
    return r;
}
// 8-byte streaming load / plain store of two adjacent pixels
__device__ __forceinline__ f2 ldg_stream2(const float* p) {
    f2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.v.x), "=f"(r.v.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void st2(float* p, f2 a) { *reinterpret_cast<float2*>(p) = a.v; }

// alpha = 1 + scale * p + eps with SEPARATELY rounded multiply and adds, as torch evaluates it (probability_helper.py:104):
// the f32x2 intrinsics may be contracted into FFMA2 by the compiler, explicit .rn PTX is not
__device__ __forceinline__ f2 alpha_from_probs2(f2 scale, f2 pr, float eps) {
    unsigned long long a, b, c, d;
    const float2 one = make_float2(1.0f, 1.0f), e = make_float2(eps, eps);
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(scale.v.x), "f"(scale.v.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(pr.v.x), "f"(pr.v.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(one.x), "f"(one.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(e.x), "f"(e.y));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(a) : "l"(a), "l"(b));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(a) : "l"(c), "l"(a));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(a) : "l"(a), "l"(d));
    f2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.v.x), "=f"(r.v.y) : "l"(a));
    return r;
}

// softplus(x) = ln(1 + e^x) and its derivative sigmoid(x) with ATen's threshold (x > 20: softplus = x), without the
// library expf / log1pf / division (~85 instructions): t = 2^(x log2 e) from one ex2, log1p(t) from a degree-6 series for
// t <= 1/16 (truncation 5e-10 relative) and from lg2(1 + t) above (relative error <= 1e-6 at the switch point, falling
// with t), sigmoid = t / (1 + t) from one reciprocal.  ~16 instructions, 3 MUFU.
__device__ __forceinline__ float softplus_fast(float x, float& dsig) {
    float t;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x * 1.4426950408889634f));
    const float u = 1.0f + t;
    float ser = fmaf(t, -1.0f / 6.0f, 0.2f);
    ser = fmaf(ser, t, -0.25f);
    ser = fmaf(ser, t, 1.0f / 3.0f);
    ser = fmaf(ser, t, -0.5f);
    ser = fmaf(ser, t, 1.0f);
    const float big = lg2_fast_(u) * 0.6931471805599453f;
    const float sp = t <= 0.0625f ? ser * t : big;
    dsig = x > 20.f ? 1.f : t * rcp_fast(u);
    return x > 20.f ? x : sp;
}

// ---- special functions on two arguments at once (slu_special.cuh has the derivations and the accuracy figures) ---------
struct PsiG2 { f2 w, g; };

// psi(x) = ln x + g(1/x), x >= 1: returns w = 1/x and g
__device__ __forceinline__ PsiG2 psi_g2(f2 x) {
    PsiG2 r;
    r.w = rcp2(x);
    f2 h = fma2(f2(2.889277183e-04f), r.w, -1.886666441e-03f);
    h = fma2(h, r.w, 4.937323876e-03f);
    h = fma2(h, r.w, -5.935221128e-03f);
    h = fma2(h, r.w, 4.237475471e-04f);
    h = fma2(h, r.w, 8.287647461e-03f);
    h = fma2(h, r.w, 1.917667073e-06f);
    h = fma2(h, r.w, -8.333334680e-02f);
    r.g = r.w * fma2(h, r.w, -0.5f);
    return r;
}
// psi(xa) - psi(xb) with a = psi_g2(xa), b = psi_g2(xb): one logarithm of the ratio per half
__device__ __forceinline__ f2 psi_diff2(f2 xa, const PsiG2& a, const PsiG2& b) {
    return fma2(lg2_2(xa * b.w), 0.6931471805599453f, a.g - b.g);
}

struct LDT2 { f2 lg, psi, tri; };

// scalar routine per half, out of line: arguments below 1 are rare (clamped concentrations) and must not bloat the loop
static __device__ __noinline__ void ldt_pos2_slow(float ax, float ay, float* out6) {
    const LDT p = ldt_pos(ax), q = ldt_pos(ay);
    out6[0] = p.lg; out6[1] = q.lg; out6[2] = p.psi; out6[3] = q.psi; out6[4] = p.tri; out6[5] = q.tri;
}

// lgamma, psi, psi' of two positive arguments.  GE1: the caller guarantees both are >= 1 (concentrations formed as
// 1 + softplus * softmax + eps, and the KL regulariser's a~ = max(1 or alpha, eps) of them): no fallback is compiled in.
template <bool GE1>
__device__ __forceinline__ LDT2 ldt_pos2(f2 a) {
    LDT2 o;
    if (!GE1 && (a.v.x < 1.0f || a.v.y < 1.0f)) {
        float t[6];
        ldt_pos2_slow(a.v.x, a.v.y, t);
        o.lg = f2(t[0], t[1]); o.psi = f2(t[2], t[3]); o.tri = f2(t[4], t[5]);
        return o;
    }
    const f2 w = rcp2(a);
    const f2 lnx = lg2_2(a) * 0.6931471805599453f;
    f2 r = fma2(f2(1.383177419e-04f), w, -6.336787229e-04f);
    r = fma2(r, w, 1.049638091e-03f); r = fma2(r, w, -5.343459393e-05f);
    r = fma2(r, w, -2.772524576e-03f); r = fma2(r, w, -1.822395578e-07f); r = fma2(r, w, 8.333333419e-02f);
    f2 h = fma2(f2(2.889277183e-04f), w, -1.886666441e-03f);
    h = fma2(h, w, 4.937323876e-03f); h = fma2(h, w, -5.935221128e-03f);
    h = fma2(h, w, 4.237475471e-04f); h = fma2(h, w, 8.287647461e-03f); h = fma2(h, w, 1.917667073e-06f);
    h = fma2(h, w, -8.333334680e-02f);
    f2 v = fma2(f2(-3.238177565e-03f), w, 1.716627286e-02f);
    v = fma2(v, w, -3.719230805e-02f); v = fma2(v, w, 3.725572734e-02f);
    v = fma2(v, w, -2.641956099e-03f); v = fma2(v, w, -3.307210709e-02f); v = fma2(v, w, -1.011315425e-05f);
    v = fma2(v, w, 1.666667325e-01f);
    o.lg = (fma2(a - 0.5f, lnx, -a) + 0.918938533204672742f) + w * r;
    o.psi = lnx + w * fma2(h, w, -0.5f);
    const f2 w2 = w * w;
    o.tri = fma2(w2 * w, v, fma2(w2, 0.5f, w));
    return o;
}


// ---- merged forms for the KL regulariser (a >= 1, w = 1/a) -------------------------------------------------------------
// The KL value needs, per class, f(a) = -lgamma(a) + (a - 1) psi(a).  With the expansions above,
//   f(a) = -ln(a)/2 + a - (ln(2 pi) + 1)/2 + w u(w),   u(w) = 1/2 - r(w) + (1 - w) h(w),  u(0) = 1/3
// so ONE degree-6 polynomial (least squares on Chebyshev nodes, tools/fit_digamma.py: |w u| error 1.6e-7) replaces the
// separate lgamma and digamma evaluations; sum_c ln a_c and sum_c a_c are accumulated by the caller.
__device__ __forceinline__ f2 kl_value_poly(f2 w) {
    f2 u = fma2(f2(-1.514686388e-03f), w, 5.008932959e-03f);
    u = fma2(u, w, -3.317222958e-03f);
    u = fma2(u, w, -9.171346347e-03f);
    u = fma2(u, w, 1.127716751e-02f);
    u = fma2(u, w, 8.332211733e-02f);
    u = fma2(u, w, 3.333334523e-01f);
    return u;
}
constexpr float KL_VALUE_CONST = 1.418938533204672742f;      // (ln(2 pi) + 1) / 2

// The KL gradient needs (a - 1) psi'(a) = (1 - w)(1 + w/2 + w^2 v(w)), v as in ldt_pos.
__device__ __forceinline__ f2 kl_grad_term(f2 w) {
    f2 v = fma2(f2(-3.238177565e-03f), w, 1.716627286e-02f);
    v = fma2(v, w, -3.719230805e-02f); v = fma2(v, w, 3.725572734e-02f);
    v = fma2(v, w, -2.641956099e-03f); v = fma2(v, w, -3.307210709e-02f); v = fma2(v, w, -1.011315425e-05f);
    v = fma2(v, w, 1.666667325e-01f);
    const f2 t = fma2(w * w, v, fma2(w, 0.5f, 1.0f));
    return fma2(-w, t, t);
}

}  // namespace slu
