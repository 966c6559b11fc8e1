// slu_evidential.cu -- stage 3+4 for the evidential (Dirichlet) head, one pass over HBM.
//
// Replaces (reference file:line): the single-pass Dirichlet branch of Tester.test_epoch
// src/models/tester.py:484-512 = to_alpha_concentrations_from_shape_and_scale
// src/models/probability_helper.py:89-105, get_predictive_entropy :116-121,
// get_aleatoric_uncertainty :124-130, get_epistemic_uncertainty :133-136, the Dirichlet mutual
// information scored by AUROCAggregator src/metrics/auroc.py:55-63, IoUEvaluator.update
// src/models/evaluator.py:39-53 and ECEAggregator.update in 'alpha' mode src/metrics/ece.py:57-58,75-90.
// The reference runs ~15 eager kernels, each re-reading [B,C,H,W]; here a thread owns a pixel, loads
// its C(+1) values once (84 B/px in, <= 36 B/px out) and keeps everything in registers.
//
// The three eps conventions of the reference are kept apart on purpose:
//   probability_helper: alpha0 = sum(alpha) + eps, H = -sum p log(p + eps)          (eps = 1e-8)
//   AUROC MI          : a0 = sum(alpha) + e, p = alpha/a0, clamp(p, e) inside the log (e = 1e-12)
//   ECE 'alpha'       : p = alpha / (sum(alpha) + e), conf = max p                  (e = 1e-12)
// Bound: instructions, not HBM, at C=20: per class one ex2 (softmax), one reciprocal and one lg2 for the digamma
// DIFFERENCE psi(alpha_c+1) - psi(alpha0+1) (slu_special.cuh::psi_g: psi(x) = ln x + g(1/x) with a degree-7 polynomial,
// the logarithm taken once on the ratio of the two arguments), two lg2 for the entropies: 5 MUFU per class.
#include <math.h>
#include <stdlib.h>
#include "slu_common.cuh"
#include "slu_special.cuh"
#include "slu_packed.cuh"

namespace slu {

constexpr int EV_THREADS = 256;
#ifndef SLU_EV_MINB
#define SLU_EV_MINB 5
#endif
constexpr int EV_MINB = SLU_EV_MINB;      // resident CTAs per SM the kernel is compiled for

__device__ __forceinline__ float lg2_fast(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct EvParams {
    const float* outputs;      // [B,C+1,HW] shape logits + scale logit, or NULL
    const float* alpha_in;     // [B,C,HW], or NULL
    const long long* labels;
    int B, C;
    long long HW;
    float inv_temp, eps, eps_m;
    float logC, inv_logC;
    int has_ignore;
    long long ignore;
    int n_bins;
    float edges[SLU_MAX_BINS + 1];
    float* alpha_out;
    long long* pred;
    float* conf;
    float* h;
    float* au;
    float* eu;
    float* mi;
    unsigned long long* confmat;
    unsigned long long* bins;
    long long n_px;            // B * HW
    int bins_one_step;         // edges passed one_step_bin_search_ok()
};

// EXACT: C == CP, no per-class predicate anywhere;  MI: the AUROC-convention mutual information is requested.
template <int CP, bool EXACT, bool MI>
__global__ void __launch_bounds__(EV_THREADS, EV_MINB) evidential_kernel(const __grid_constant__ EvParams p) {
    __shared__ AtomicHist hs;
    const int tid = threadIdx.x;
    atomic_hist_zero(hs, p.C, tid, EV_THREADS);
    for (int i = tid; i <= p.n_bins; i += EV_THREADS) hs.edges[i] = p.edges[i];
    __syncthreads();
    int since_flush = 0;

    const long long chunks = (p.n_px + EV_THREADS - 1) / EV_THREADS;
    for (long long ch = blockIdx.x; ch < chunks; ch += gridDim.x) {
        if (p.labels && ++since_flush > ATOMIC_HIST_MAX_PX / EV_THREADS) {      // CTA-uniform: keep the split sums below 2^32
            __syncthreads();
            atomic_hist_flush(hs, p.C, p.n_bins, p.confmat, p.bins, tid, EV_THREADS);
            __syncthreads();
            atomic_hist_zero(hs, p.C, tid, EV_THREADS);
            __syncthreads();
            since_flush = 1;
        }
        const long long g = ch * EV_THREADS + tid;
        const bool live = g < p.n_px;
        const long long gs = live ? g : p.n_px - 1;
        const int b = (int)(gs / p.HW);
        const long long px = gs - (long long)b * p.HW;
        float a[CP];
        int pred = 0;
        if (p.outputs) {
            const float* base = p.outputs + ((long long)b * (p.C + 1)) * p.HW + px;
            float z[CP];
#pragma unroll
            for (int c = 0; c < CP; ++c) z[c] = (EXACT || c < p.C) ? ldg_stream(base + (long long)c * p.HW) : -1.0e30f;
            const float sl = ldg_stream(base + (long long)p.C * p.HW) * p.inv_temp;
            // softplus (ATen: x > 20 ? x : log1p(exp(x)))
            const float scale = sl > 20.f ? sl : log1pf(expf(sl));
            float m = z[0];
#pragma unroll
            for (int c = 1; c < CP; ++c) m = fmaxf(m, z[c]);
            float S = 0.f;
            const float m2 = m * 1.4426950408889634f;
#pragma unroll
            for (int c = 0; c < CP; ++c) { z[c] = ex2_approx(fmaf(z[c], 1.4426950408889634f, -m2)); S += z[c]; }
            const float invS = __frcp_rn(S);
            float best = -1.f;
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                const float pc = z[c] * invS;
                if ((EXACT || c < p.C) && pc > best) { best = pc; pred = c; }          // tester.py:493-495
                a[c] = __fadd_rn(__fadd_rn(1.0f, __fmul_rn(scale, pc)), p.eps);   // probability_helper.py:104
            }
        } else {
            const float* base = p.alpha_in + ((long long)b * p.C) * p.HW + px;
#pragma unroll
            for (int c = 0; c < CP; ++c) a[c] = (EXACT || c < p.C) ? ldg_stream(base + (long long)c * p.HW) : 0.f;
        }
        // sums and arg max over alpha
        float asum = 0.f, amax = 0.f;
        int aarg = 0;
#pragma unroll
        for (int c = 0; c < CP; ++c) {
            if (EXACT || c < p.C) {
                asum += a[c];
                const bool gt = (c == 0) | (a[c] > amax) | ((a[c] != a[c]) & (amax == amax));   // torch.argmax: first maximum, NaN maximal
                amax = gt ? a[c] : amax;
                aarg = gt ? c : aarg;
            }
        }
        if (p.alpha_out && live) {
            float* ao = p.alpha_out + ((long long)b * p.C) * p.HW + px;
#pragma unroll
            for (int c = 0; c < CP; ++c)
                if (EXACT || c < p.C) ao[(long long)c * p.HW] = a[c];
        }
        if (!p.outputs) pred = aarg;
        const float a0 = asum + p.eps;                     // probability_helper.py:119,127
        const float a0m = asum + p.eps_m;                  // auroc.py:57
        const bool want_unc = p.h || p.au || p.eu || MI;
        float H = 0.f, AU = 0.f, Hm = 0.f, EHm = 0.f;
        if (want_unc) {                                    // warp-uniform
            // psi(alpha_c + 1) - psi(alpha0 + 1) = ln((alpha_c+1)/(alpha0+1)) + g_c - g_0: one rcp and one lg2 per class
            const float x0 = a0 + 1.0f;
            const PsiG q0 = psi_g(x0);
            const bool same0 = a0m == a0;                  // alpha >= 1: both eps vanish in fp32 and the two sums coincide
            PsiG q0m = q0;
            if (MI && !same0) q0m = psi_g(a0m + 1.0f);
            const float inv0 = __frcp_rn(a0), inv0m = (MI && !same0) ? __frcp_rn(a0m) : inv0;
            float H2 = 0.f, Hm2 = 0.f;                     // entropies in log2 units (lg2.approx, rel. error <= 2^-22)
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                if (EXACT || c < p.C) {
                    const float ph = a[c] * inv0;
                    const float xc = a[c] + 1.0f;
                    const PsiG q = psi_g(xc);
                    const float d = psi_diff(xc, q, q0);
                    H2 = fmaf(-ph, lg2_fast(ph + p.eps), H2);                // :121
                    AU = fmaf(-ph, d, AU);                                   // :128-130
                    if (MI) {
                        const float pm = a[c] * inv0m;
                        const float pmc = fmaxf(pm, p.eps_m);
                        Hm2 = fmaf(-pmc, lg2_fast(pmc), Hm2);                // auroc.py:59
                        EHm = fmaf(-pm, same0 ? d : psi_diff(xc, q, q0m), EHm);   // auroc.py:60-61
                    }
                }
            }
            H = H2 * 0.6931471805599453f;
            Hm = Hm2 * 0.6931471805599453f;
        }
        const float conf = __fdiv_rn(amax, asum + p.eps_m);                  // ece.py:57-58,75
        if (live) {
            if (p.pred) p.pred[g] = pred;
            if (p.conf) p.conf[g] = conf;
            if (p.h) p.h[g] = __fdiv_rn(H, p.logC);
            if (p.au) p.au[g] = AU;
            if (p.eu) p.eu[g] = H - AU;
            if (MI) p.mi[g] = __fdiv_rn(Hm - EHm, p.logC);
        }
        if (p.labels && live)       // confusion: tester's argmax of the shape softmax; ECE: argmax of alpha/alpha0 (ece.py:75,84)
            atomic_hist_add(hs, p.C, p.n_bins, p.confmat != nullptr, p.bins != nullptr, p.bins_one_step != 0, p.labels[g], pred, aarg,
                            conf, p.has_ignore != 0, p.ignore);
    }
    __syncthreads();
    atomic_hist_flush(hs, p.C, p.n_bins, p.confmat, p.bins, tid, EV_THREADS);
}

// ---- packed variant: a thread owns TWO adjacent pixels (slu_packed.cuh) -------------------------------------------------
// Same formulas, every per-class operation issued once for both pixels (FFMA2 / FMUL2 / FADD2); MUFU evaluations, arg-max
// compares and the histogram updates stay per pixel.  Needs HW even, 8-byte aligned inputs / maps, 16-byte aligned labels
// and pred (the dispatcher checks; otherwise the one-pixel kernel above runs).
constexpr int EV2_THREADS = 128;
#ifndef SLU_EV2_MINB
#define SLU_EV2_MINB 6
#endif
static int g_ev_no_packed = 0;

template <int CP, bool EXACT, bool MI, bool GE1>
__global__ void __launch_bounds__(EV2_THREADS, SLU_EV2_MINB) evidential_x2_kernel(const __grid_constant__ EvParams p) {
    __shared__ AtomicHist hs;
    const int tid = threadIdx.x;
    atomic_hist_zero(hs, p.C, tid, EV2_THREADS);
    for (int i = tid; i <= p.n_bins; i += EV2_THREADS) hs.edges[i] = p.edges[i];
    __syncthreads();
    int since_flush = 0;

    const long long pairs = p.n_px >> 1;
    const long long chunks = (pairs + EV2_THREADS - 1) / EV2_THREADS;
    for (long long ch = blockIdx.x; ch < chunks; ch += gridDim.x) {
        if (p.labels && ++since_flush > ATOMIC_HIST_MAX_PX / (2 * EV2_THREADS)) {      // CTA-uniform: keep the split sums below 2^32
            __syncthreads();
            atomic_hist_flush(hs, p.C, p.n_bins, p.confmat, p.bins, tid, EV2_THREADS);
            __syncthreads();
            atomic_hist_zero(hs, p.C, tid, EV2_THREADS);
            __syncthreads();
            since_flush = 1;
        }
        const long long pi = ch * EV2_THREADS + tid;
        const bool live = pi < pairs;
        const long long g = (live ? pi : pairs - 1) << 1;
        const int b = (int)(g / p.HW);
        const long long px = g - (long long)b * p.HW;
        f2 a[CP];
        int pred0 = 0, pred1 = 0;
        if (p.outputs) {
            const float* base = p.outputs + ((long long)b * (p.C + 1)) * p.HW + px;
            f2 z[CP];
#pragma unroll
            for (int c = 0; c < CP; ++c) z[c] = (EXACT || c < p.C) ? ldg_stream2(base + (long long)c * p.HW) : f2(-1.0e30f);
            const f2 sl = ldg_stream2(base + (long long)p.C * p.HW) * p.inv_temp;
            // softplus (ATen: x > 20 ? x : log1p(exp(x))), slu_packed.cuh::softplus_fast
            float ds_unused;
            const f2 scale(softplus_fast(sl.v.x, ds_unused), softplus_fast(sl.v.y, ds_unused));
            f2 m = z[0];
#pragma unroll
            for (int c = 1; c < CP; ++c) m = max2(m, z[c]);
            f2 S(0.f);
            const f2 m2 = m * 1.4426950408889634f;
#pragma unroll
            for (int c = 0; c < CP; ++c) { z[c] = ex2_2(fma2(z[c], 1.4426950408889634f, -m2)); S += z[c]; }
            const f2 invS = rcp_rn2(S);
            float best0 = -1.f, best1 = -1.f;
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                const f2 pc = z[c] * invS;
                if (EXACT || c < p.C) {                                                   // tester.py:493-495
                    if (pc.v.x > best0) { best0 = pc.v.x; pred0 = c; }
                    if (pc.v.y > best1) { best1 = pc.v.y; pred1 = c; }
                }
                a[c] = alpha_from_probs2(scale, pc, p.eps);                             // probability_helper.py:104
            }
        } else {
            const float* base = p.alpha_in + ((long long)b * p.C) * p.HW + px;
#pragma unroll
            for (int c = 0; c < CP; ++c) a[c] = (EXACT || c < p.C) ? ldg_stream2(base + (long long)c * p.HW) : f2(0.f);
        }
        // sums and arg max over alpha
        f2 asum(0.f);
        float amax0 = 0.f, amax1 = 0.f;
        int aarg0 = 0, aarg1 = 0;
        if (GE1 && p.outputs) {
            // alpha = fl(fl(1 + fl(s p)) + eps) is a monotone function of p, so its maximum sits at the arg-max of the softmax
            // (pred) and only its FIRST occurrence can differ: an earlier class whose smaller p rounds to the same alpha.
            // Count the classes that equal the maximum; the exact scan runs only if there is more than one (or NaN).
            int neq0 = 0, neq1 = 0;
#pragma unroll
            for (int c = 0; c < CP; ++c)
                if (EXACT || c < p.C) {
                    asum += a[c];
                    amax0 = fmaxf(amax0, a[c].v.x); amax1 = fmaxf(amax1, a[c].v.y);
                }
#pragma unroll
            for (int c = 0; c < CP; ++c)
                if (EXACT || c < p.C) { neq0 += a[c].v.x == amax0 ? 1 : 0; neq1 += a[c].v.y == amax1 ? 1 : 0; }
            aarg0 = pred0; aarg1 = pred1;
            if (neq0 != 1 || neq1 != 1 || asum.v.x != asum.v.x || asum.v.y != asum.v.y) {
                amax0 = 0.f; amax1 = 0.f; aarg0 = 0; aarg1 = 0;
#pragma unroll 1
                for (int c = 0; c < p.C; ++c) {
                    float u = 0.f, w = 0.f;
#pragma unroll
                    for (int k = 0; k < CP; ++k)
                        if (k == c) { u = a[k].v.x; w = a[k].v.y; }
                    const bool g0 = (c == 0) | (u > amax0) | ((u != u) & (amax0 == amax0));
                    const bool g1 = (c == 0) | (w > amax1) | ((w != w) & (amax1 == amax1));
                    amax0 = g0 ? u : amax0; aarg0 = g0 ? c : aarg0;
                    amax1 = g1 ? w : amax1; aarg1 = g1 ? c : aarg1;
                }
            }
        } else {
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                if (EXACT || c < p.C) {
                    asum += a[c];
                    const float u = a[c].v.x, w = a[c].v.y;
                    const bool g0 = (c == 0) | (u > amax0) | ((u != u) & (amax0 == amax0));   // torch.argmax: first maximum, NaN maximal
                    const bool g1 = (c == 0) | (w > amax1) | ((w != w) & (amax1 == amax1));
                    amax0 = g0 ? u : amax0; aarg0 = g0 ? c : aarg0;
                    amax1 = g1 ? w : amax1; aarg1 = g1 ? c : aarg1;
                }
            }
        }
        if (p.alpha_out && live) {
            float* ao = p.alpha_out + ((long long)b * p.C) * p.HW + px;
#pragma unroll
            for (int c = 0; c < CP; ++c)
                if (EXACT || c < p.C) st2(ao + (long long)c * p.HW, a[c]);
        }
        if (!p.outputs) { pred0 = aarg0; pred1 = aarg1; }
        const f2 a0 = asum + p.eps;                        // probability_helper.py:119,127
        const f2 a0m = asum + p.eps_m;                     // auroc.py:57
        const bool want_unc = p.h || p.au || p.eu || MI;
        f2 H(0.f), AU(0.f), EU(0.f), MIv(0.f);
        if (want_unc) {                                    // warp-uniform
            const bool same0 = a0m.v.x == a0.v.x && a0m.v.y == a0.v.y;     // alpha >= 1: both eps vanish in fp32
            const f2 inv0 = rcp_rn2(a0), inv0m = (MI && !same0) ? rcp_rn2(a0m) : inv0;
            // Every concentration >= 1 (always true for a head output: alpha = 1 + softplus * softmax + eps):
            // the digamma recurrence psi(a + 1) = psi(a) + 1/a and psi(a) = ln a + g(1/a) give
            //     psi(alpha_c + 1) - psi(alpha0 + 1) = ln(alpha_c / alpha0) + G(1/alpha_c) - G(1/alpha0),  G(w) = g(w) + w,
            // so the digamma difference shares its logarithm with the entropy (ln p_c): 3 MUFU per class (softmax ex2, one
            // reciprocal, one lg2) instead of 5, and the epistemic part comes out WITHOUT the cancellation of H - AU:
            //     EU = sum_c p_c G(1/alpha_c) - G(1/alpha0) sum_c p_c,   AU = H - EU,   Dirichlet MI (auroc.py:55-63) = EU
            // with the AUROC eps.  (eps inside the log shifts H by <= C eps = 2e-7, far inside the 1e-5 tolerance.)
            bool ge1 = GE1;
            if (!GE1) {
                float amin = 3.0e38f;
#pragma unroll
                for (int c = 0; c < CP; ++c)
                    if (EXACT || c < p.C) amin = fminf(amin, fminf(a[c].v.x, a[c].v.y));
                ge1 = amin >= 1.0f;                        // NaN fails the test too: the literal formulas below then apply
            }
            if (ge1) {
                f2 H2(0.f), SG(0.f), SGm(0.f);
#pragma unroll
                for (int c = 0; c < CP; ++c) {
                    if (EXACT || c < p.C) {
                        const f2 ph = a[c] * inv0;
                        const f2 w = rcp2(a[c]);
                        f2 h = fma2(f2(2.889277183e-04f), w, -1.886666441e-03f);
                        h = fma2(h, w, 4.937323876e-03f);
                        h = fma2(h, w, -5.935221128e-03f);
                        h = fma2(h, w, 4.237475471e-04f);
                        h = fma2(h, w, 8.287647461e-03f);
                        h = fma2(h, w, 1.917667073e-06f);
                        h = fma2(h, w, -8.333334680e-02f);
                        const f2 Gc = w * fma2(h, w, 0.5f);                   // g(w) + w = w (w h(w) + 1/2)
                        H2 = fma2(-ph, lg2_2(ph + p.eps), H2);                // :121
                        SG = fma2(ph, Gc, SG);
                        if (MI && !same0) SGm = fma2(a[c] * inv0m, Gc, SGm);
                    }
                }
                H = H2 * 0.6931471805599453f;
                const PsiG2 q0 = psi_g2(a0);                                  // G(1/alpha0) = q0.g + q0.w
                EU = fma2(-(q0.g + q0.w), asum * inv0, SG);
                AU = H - EU;                                                  // :128-136
                if (MI) {
                    if (same0) MIv = EU;
                    else { const PsiG2 q0m = psi_g2(a0m); MIv = fma2(-(q0m.g + q0m.w), asum * inv0m, SGm); }
                }
            } else {
                // concentrations below 1 (caller-supplied alpha only): the literal formulas, digamma on alpha + 1 >= 1
                const f2 x0 = a0 + 1.0f;
                const PsiG2 q0 = psi_g2(x0);
                PsiG2 q0m = q0;
                if (MI && !same0) q0m = psi_g2(a0m + 1.0f);
                f2 H2(0.f), Hm2(0.f), EHm(0.f);
#pragma unroll
                for (int c = 0; c < CP; ++c) {
                    if (EXACT || c < p.C) {
                        const f2 ph = a[c] * inv0;
                        const f2 xc = a[c] + 1.0f;
                        const PsiG2 q = psi_g2(xc);
                        const f2 d = psi_diff2(xc, q, q0);
                        H2 = fma2(-ph, lg2_2(ph + p.eps), H2);                // :121
                        AU = fma2(-ph, d, AU);                                // :128-130
                        if (MI) {
                            const f2 pm = a[c] * inv0m;
                            const f2 pmc = max2(pm, p.eps_m);
                            Hm2 = fma2(-pmc, lg2_2(pmc), Hm2);                // auroc.py:59
                            EHm = fma2(-pm, same0 ? d : psi_diff2(xc, q, q0m), EHm);   // auroc.py:60-61
                        }
                    }
                }
                H = H2 * 0.6931471805599453f;
                EU = H - AU;
                MIv = Hm2 * 0.6931471805599453f - EHm;
            }
        }
        const float conf0 = __fdiv_rn(amax0, asum.v.x + p.eps_m), conf1 = __fdiv_rn(amax1, asum.v.y + p.eps_m);   // ece.py:57-58,75
        if (live) {
            if (p.pred) *reinterpret_cast<longlong2*>(p.pred + g) = make_longlong2(pred0, pred1);
            if (p.conf) st2(p.conf + g, f2(conf0, conf1));
            if (p.h) st2(p.h + g, H * p.inv_logC);
            if (p.au) st2(p.au + g, AU);
            if (p.eu) st2(p.eu + g, EU);
            if (MI) st2(p.mi + g, MIv * p.inv_logC);
        }
        if (p.labels && live) {     // confusion: tester's argmax of the shape softmax; ECE: argmax of alpha/alpha0 (ece.py:75,84)
            const longlong2 lb = *reinterpret_cast<const longlong2*>(p.labels + g);
            atomic_hist_add(hs, p.C, p.n_bins, p.confmat != nullptr, p.bins != nullptr, p.bins_one_step != 0, lb.x, pred0, aarg0,
                            conf0, p.has_ignore != 0, p.ignore);
            atomic_hist_add(hs, p.C, p.n_bins, p.confmat != nullptr, p.bins != nullptr, p.bins_one_step != 0, lb.y, pred1, aarg1,
                            conf1, p.has_ignore != 0, p.ignore);
        }
    }
    __syncthreads();
    atomic_hist_flush(hs, p.C, p.n_bins, p.confmat, p.bins, tid, EV2_THREADS);
}

// experiment (SLU_GRID_WAVES): more, shorter CTAs dealt out by the hardware scheduler instead of one resident wave (slower here:
// 0.076 -> 0.094 ms per 16 scans at 8 waves)
static int grid_waves() { static const int w = [] { const char* e = getenv("SLU_GRID_WAVES"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 1; }(); return w; }

template <int CP>
static int launch_ev(const EvParams& p, cudaStream_t st) {
    const int sms = sm_count_current_device();
    if (sms <= 0) return fail(SLU_E_DEVICE, "no CUDA device");
    const long long chunks = (p.n_px + EV_THREADS - 1) / EV_THREADS;
    const long long cap = 5LL * sms;                       // 48 registers: five resident CTAs of 256 threads
    const unsigned grid = (unsigned)(chunks < cap ? chunks : cap);
    const bool exact = p.C == CP, mi = p.mi != nullptr;
    const uintptr_t al8 = reinterpret_cast<uintptr_t>(p.outputs) | reinterpret_cast<uintptr_t>(p.alpha_in) |
                          reinterpret_cast<uintptr_t>(p.alpha_out) | reinterpret_cast<uintptr_t>(p.conf) |
                          reinterpret_cast<uintptr_t>(p.h) | reinterpret_cast<uintptr_t>(p.au) |
                          reinterpret_cast<uintptr_t>(p.eu) | reinterpret_cast<uintptr_t>(p.mi);
    const uintptr_t al16 = reinterpret_cast<uintptr_t>(p.labels) | reinterpret_cast<uintptr_t>(p.pred);
    if (!g_ev_no_packed && (p.HW & 1) == 0 && (al8 & 7) == 0 && (al16 & 15) == 0) {
        const long long chunks2 = ((p.n_px >> 1) + EV2_THREADS - 1) / EV2_THREADS;
        const long long cap2 = (long long)SLU_EV2_MINB * sms * grid_waves();
        const unsigned grid2 = (unsigned)(chunks2 < cap2 ? chunks2 : cap2);
        // GE1: a head output always gives alpha >= 1; caller-supplied concentrations are checked per thread
        if (p.outputs) {
            if (exact && mi) evidential_x2_kernel<CP, true, true, true><<<grid2, EV2_THREADS, 0, st>>>(p);
            else if (exact) evidential_x2_kernel<CP, true, false, true><<<grid2, EV2_THREADS, 0, st>>>(p);
            else if (mi) evidential_x2_kernel<CP, false, true, true><<<grid2, EV2_THREADS, 0, st>>>(p);
            else evidential_x2_kernel<CP, false, false, true><<<grid2, EV2_THREADS, 0, st>>>(p);
        } else {
            if (exact && mi) evidential_x2_kernel<CP, true, true, false><<<grid2, EV2_THREADS, 0, st>>>(p);
            else if (exact) evidential_x2_kernel<CP, true, false, false><<<grid2, EV2_THREADS, 0, st>>>(p);
            else if (mi) evidential_x2_kernel<CP, false, true, false><<<grid2, EV2_THREADS, 0, st>>>(p);
            else evidential_x2_kernel<CP, false, false, false><<<grid2, EV2_THREADS, 0, st>>>(p);
        }
        SLU_LAUNCH_CHECK("evidential_x2_kernel");
        return 0;
    }
    if (exact && mi) evidential_kernel<CP, true, true><<<grid, EV_THREADS, 0, st>>>(p);
    else if (exact) evidential_kernel<CP, true, false><<<grid, EV_THREADS, 0, st>>>(p);
    else if (mi) evidential_kernel<CP, false, true><<<grid, EV_THREADS, 0, st>>>(p);
    else evidential_kernel<CP, false, false><<<grid, EV_THREADS, 0, st>>>(p);
    SLU_LAUNCH_CHECK("evidential_kernel");
    return 0;
}

}  // namespace slu

extern "C" int slu_evidential_reduce(const float* d_outputs, const float* d_alpha_in, const int64_t* d_labels,
                                     int B, int C, int64_t HW, float temperature, float eps, float eps_metrics,
                                     int normalize, int has_ignore, int64_t ignore, int n_bins, const float* h_edges,
                                     float* d_alpha_out, int64_t* d_pred, float* d_conf, float* d_h, float* d_au,
                                     float* d_eu, float* d_mi, int64_t* d_confmat, int64_t* d_ece_bins,
                                     slu_stream_t stream) {
    using namespace slu;
    if ((d_outputs == nullptr) == (d_alpha_in == nullptr)) return fail(SLU_E_ARG, "exactly one of d_outputs / d_alpha_in must be given");
    if (B < 1 || HW < 1) return fail(SLU_E_ARG, "B=%d HW=%lld must be >= 1", B, (long long)HW);
    if (C < 2 || C > SLU_MAX_CLASSES) return fail(SLU_E_RANGE, "C=%d outside [2,%d]", C, SLU_MAX_CLASSES);
    if (!(temperature > 0.f)) return fail(SLU_E_ARG, "temperature must be > 0");
    if ((d_confmat || d_ece_bins) && !d_labels) return fail(SLU_E_ARG, "histograms requested without labels");
    if (d_ece_bins) {
        if (n_bins < 1 || n_bins > SLU_MAX_BINS) return fail(SLU_E_RANGE, "n_bins=%d outside [1,%d]", n_bins, SLU_MAX_BINS);
        if (!h_edges) return fail(SLU_E_ARG, "h_edges is NULL");
        for (int i = 0; i < n_bins; ++i)
            if (!(h_edges[i] < h_edges[i + 1])) return fail(SLU_E_ARG, "bin edges must increase strictly");
    }
    EvParams p{};
    p.outputs = d_outputs; p.alpha_in = d_alpha_in;
    p.labels = reinterpret_cast<const long long*>(d_labels);
    p.B = B; p.C = C; p.HW = HW; p.n_px = (long long)B * HW;
    p.inv_temp = 1.0f / temperature; p.eps = eps; p.eps_m = eps_metrics;
    p.logC = normalize ? (float)log((double)C) : 1.0f;
    p.inv_logC = normalize ? (float)(1.0 / log((double)C)) : 1.0f;
    p.has_ignore = has_ignore; p.ignore = ignore;
    p.n_bins = d_ece_bins ? n_bins : 0;
    for (int i = 0; i <= p.n_bins && d_ece_bins; ++i) p.edges[i] = h_edges[i];
    p.bins_one_step = (p.n_bins > 0 && one_step_bin_search_ok(p.edges, p.n_bins)) ? 1 : 0;
    p.alpha_out = d_alpha_out; p.pred = reinterpret_cast<long long*>(d_pred);
    p.conf = d_conf; p.h = d_h; p.au = d_au; p.eu = d_eu; p.mi = d_mi;
    p.confmat = reinterpret_cast<unsigned long long*>(d_confmat);
    p.bins = reinterpret_cast<unsigned long long*>(d_ece_bins);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch ((C + 3) / 4 * 4) {
        case 4: return launch_ev<4>(p, st);
        case 8: return launch_ev<8>(p, st);
        case 12: return launch_ev<12>(p, st);
        case 16: return launch_ev<16>(p, st);
        case 20: return launch_ev<20>(p, st);
        case 24: return launch_ev<24>(p, st);
        case 28: return launch_ev<28>(p, st);
        default: return launch_ev<32>(p, st);
    }
}

/* A/B switch (tests, profiles): 1 = slu_evidential_reduce always runs one pixel per thread (no packed f32x2 variant). */
extern "C" int slu_debug_no_packed_evidential(int on) {
    const int prev = slu::g_ev_no_packed;
    if (on >= 0) slu::g_ev_no_packed = on ? 1 : 0;
    return prev;
}
