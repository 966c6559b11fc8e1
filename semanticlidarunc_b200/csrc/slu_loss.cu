// slu_loss.cu -- evidential loss terms, forward + backward in one pass (config 5 of BASELINE.json).
//
// Replaces (reference file:line): DirichletMSELoss src/losses/dirichlet_losses.py:317-385 and
// KL_offClasses_to_uniform src/losses/regularizers.py:291-389 (with_conf_weighting=False), the two
// terms the shipped configs enable (src/configs/SemanticKitti_default.yaml:50-62), with the mask
// semantics of _valid_mask src/losses/dirichlet_losses.py:15-70.  The reference materialises a
// one-hot [B,C,H,W], several [B,C,H,W] temporaries and a permuted boolean gather, then autograd walks
// the graph backwards; here a thread owns a pixel, reads its C concentrations once and writes the
// analytic per-pixel gradients of each term, so each term stays an independently differentiable scalar
// (GradNorm calls autograd.grad per term, src/utils/grad_norm.py:52).
//
// Per pixel, with a0 = sum(alpha), D = a0 + eps, p = alpha / D, y = one-hot(target):
//   mse  = sum_c (y_c - p_c)^2 + (a0^2 - sum alpha_c^2) / G,          G = (a0^2 + eps)(a0 + 1)
//   dmse/dalpha_j = -2 (y_j - p_j)/D + 2 Q/D + (2 a0 - 2 alpha_j)/G - N G'/G^2
//                   Q = p_y - sum p_c^2,  N = a0^2 - sum alpha_c^2,  G' = 2 a0 (a0+1) + a0^2 + eps
//   kl   = lgamma(s) - sum lgamma(a_c) + sum (a_c - 1)(psi(a_c) - psi(s)),  a = max(y + (1-y) alpha, eps), s = sum a
//   dkl/dalpha_j = (a_j - 1) psi'(a_j) - (s - C) psi'(s)      for j != target and alpha_j > eps, else 0
// Loss = sum over valid pixels / max(n_valid, 1); the division and the upstream gradient are applied by
// the caller (one scalar), so the kernel needs no second pass.
// Bound: instruction-bound (lgamma + digamma + trigamma per class), 88 B/px read, 84 B/px written per term.
#include <math.h>
#include <stdlib.h>
#include "slu_common.cuh"
#include "slu_special.cuh"
#include "slu_packed.cuh"

namespace slu {

constexpr int LOSS_THREADS = 256;
#ifndef SLU_LOSS_MINB
#define SLU_LOSS_MINB 3
#endif
constexpr int LOSS_MINB = SLU_LOSS_MINB;       // resident CTAs per SM the class-loop kernels are compiled for
constexpr int MAX_IGNORE = 8;

struct LossParams {
    const float* alpha;
    const long long* target;
    const unsigned char* keep;     // optional [B*HW] bool mask, 1 = valid (overrides the ignore list)
    int B, C;
    long long HW, n_px;
    long long ignore[MAX_IGNORE];
    int n_ignore;
    float eps_mse, eps_kl;
    int want_mse, want_kl;
    double* sums;                  // [3] sum mse | sum kl | n_valid
    float* grad_mse;               // [B,C,HW] or NULL
    float* grad_kl;                // [B,C,HW] or NULL
};

// EXACT: C == CP (no per-class predicate).  The true class enters the KL sums as a~ = 1, for which lgamma = 0 and
// (a~ - 1) = 0, so the class loop needs no branch on c == y (y differs from thread to thread).
template <int CP, bool EXACT>
__global__ void __launch_bounds__(LOSS_THREADS, LOSS_MINB) dirichlet_loss_kernel(const __grid_constant__ LossParams p) {
    const int tid = threadIdx.x;
    double acc_mse = 0.0, acc_kl = 0.0;
    unsigned n_valid = 0;
    const long long chunks = (p.n_px + LOSS_THREADS - 1) / LOSS_THREADS;
    for (long long ch = blockIdx.x; ch < chunks; ch += gridDim.x) {
        const long long g = ch * LOSS_THREADS + tid;
        if (g >= p.n_px) continue;
        const int b = (int)(g / p.HW);
        const long long px = g - (long long)b * p.HW;
        const long long tgt = p.target[g];
        bool valid;
        if (p.keep) {
            valid = p.keep[g] != 0;
        } else {
            valid = true;
#pragma unroll
            for (int i = 0; i < MAX_IGNORE; ++i)
                if (i < p.n_ignore && tgt == p.ignore[i]) valid = false;
        }
        const float* base = p.alpha + ((long long)b * p.C) * p.HW + px;
        float* gm = p.grad_mse ? p.grad_mse + ((long long)b * p.C) * p.HW + px : nullptr;
        float* gk = p.grad_kl ? p.grad_kl + ((long long)b * p.C) * p.HW + px : nullptr;
        if (!valid) {                      // masked pixels contribute nothing and get zero gradient
            for (int c = 0; c < p.C; ++c) {
                if (gm) gm[(long long)c * p.HW] = 0.f;
                if (gk) gk[(long long)c * p.HW] = 0.f;
            }
            continue;
        }
        ++n_valid;
        float a[CP];
#pragma unroll
        for (int c = 0; c < CP; ++c) a[c] = (EXACT || c < p.C) ? ldg_stream(base + (long long)c * p.HW) : 0.f;
        const int y = (int)tgt;            // the reference's scatter_ requires 0 <= target < C on valid pixels

        if (p.want_mse) {
            float a0 = 0.f, s2 = 0.f, ay = 0.f;
#pragma unroll
            for (int c = 0; c < CP; ++c) { a0 += a[c]; s2 = fmaf(a[c], a[c], s2); if (c == y) ay = a[c]; }
            const float D = a0 + p.eps_mse, invD = 1.0f / D;
            const float G = (a0 * a0 + p.eps_mse) * (a0 + 1.0f), invG = 1.0f / G;
            float sq = 0.f, sp2 = 0.f;
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                if (EXACT || c < p.C) {
                    const float pc = a[c] * invD;
                    const float d = (c == y ? 1.0f : 0.0f) - pc;
                    sq = fmaf(d, d, sq);
                    sp2 = fmaf(pc, pc, sp2);
                }
            }
            float var = 0.f;
#pragma unroll
            for (int c = 0; c < CP; ++c)
                if (EXACT || c < p.C) var = fmaf(a[c] * (a0 - a[c]), invG, var);
            acc_mse += (double)(sq + var);
            if (gm) {
                const float N = a0 * a0 - s2;
                const float Gp = 2.0f * a0 * (a0 + 1.0f) + (a0 * a0 + p.eps_mse);
                const float Q = ay * invD - sp2;
                const float common = 2.0f * Q * invD + 2.0f * a0 * invG - N * Gp * invG * invG;
#pragma unroll
                for (int c = 0; c < CP; ++c) {
                    if (EXACT || c < p.C) {
                        const float pc = a[c] * invD;
                        const float yc = (c == y ? 1.0f : 0.0f);
                        gm[(long long)c * p.HW] = fmaf(-2.0f * (yc - pc), invD, fmaf(-2.0f * a[c], invG, common));
                    }
                }
            }
        }
        if (p.want_kl) {
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < CP; ++c)
                if (EXACT || c < p.C) s += fmaxf(c == y ? 1.0f : a[c], p.eps_kl);
            const LDT fs = ldt_pos(s);
            float kl = fs.lg;
            const float tail = (s - (float)p.C) * fs.tri;
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                if (EXACT || c < p.C) {
                    const float ac = fmaxf(c == y ? 1.0f : a[c], p.eps_kl);
                    const LDT f = ldt_pos(ac);
                    kl -= f.lg;
                    kl = fmaf(ac - 1.0f, f.psi - fs.psi, kl);
                    const float gj = (c != y && a[c] > p.eps_kl) ? fmaf(ac - 1.0f, f.tri, -tail) : 0.f;
                    if (gk) gk[(long long)c * p.HW] = gj;
                }
            }
            acc_kl += (double)kl;
        }
    }
    // block reduction -> one atomic triple per CTA
    __shared__ double s_m[LOSS_THREADS / 32], s_k[LOSS_THREADS / 32];
    __shared__ unsigned s_n[LOSS_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc_mse += __shfl_xor_sync(0xffffffffu, acc_mse, o);
        acc_kl += __shfl_xor_sync(0xffffffffu, acc_kl, o);
        n_valid += __shfl_xor_sync(0xffffffffu, n_valid, o);
    }
    if ((tid & 31) == 0) { s_m[tid >> 5] = acc_mse; s_k[tid >> 5] = acc_kl; s_n[tid >> 5] = n_valid; }
    __syncthreads();
    if (tid == 0) {
        double m = 0.0, k = 0.0, n = 0.0;
        for (int i = 0; i < LOSS_THREADS / 32; ++i) { m += s_m[i]; k += s_k[i]; n += (double)s_n[i]; }
        if (p.want_mse) atomicAdd(&p.sums[0], m);
        if (p.want_kl) atomicAdd(&p.sums[1], k);
        atomicAdd(&p.sums[2], n);
    }
}

// ---- packed variant of dirichlet_loss_kernel: two adjacent pixels per thread (slu_packed.cuh) -------------------------
constexpr int DL2_THREADS = 128;
#ifndef SLU_DL2_MINB
#define SLU_DL2_MINB 5
#endif
static int g_no_packed = 0;        // A/B switch (slu_debug_no_packed_loss): 1 = always run the one-pixel-per-thread kernels

template <int CP, bool EXACT>
__global__ void __launch_bounds__(DL2_THREADS, SLU_DL2_MINB) dirichlet_loss_x2_kernel(const __grid_constant__ LossParams p) {
    const int tid = threadIdx.x;
    double acc_mse = 0.0, acc_kl = 0.0;
    unsigned n_valid = 0;
    const long long pairs = p.n_px >> 1;
    const long long chunks = (pairs + DL2_THREADS - 1) / DL2_THREADS;
    for (long long ch = blockIdx.x; ch < chunks; ch += gridDim.x) {
        const long long pi = ch * DL2_THREADS + tid;
        if (pi >= pairs) continue;
        const long long g = pi << 1;
        const int b = (int)(g / p.HW);
        const long long px = g - (long long)b * p.HW;
        const longlong2 tg = *reinterpret_cast<const longlong2*>(p.target + g);
        bool v0, v1;
        if (p.keep) {
            v0 = p.keep[g] != 0; v1 = p.keep[g + 1] != 0;
        } else {
            v0 = true; v1 = true;
#pragma unroll
            for (int i = 0; i < MAX_IGNORE; ++i)
                if (i < p.n_ignore) { if (tg.x == p.ignore[i]) v0 = false; if (tg.y == p.ignore[i]) v1 = false; }
        }
        const float* base = p.alpha + ((long long)b * p.C) * p.HW + px;
        float* gm = p.grad_mse ? p.grad_mse + ((long long)b * p.C) * p.HW + px : nullptr;
        float* gk = p.grad_kl ? p.grad_kl + ((long long)b * p.C) * p.HW + px : nullptr;
        if (!v0 && !v1) {
            for (int c = 0; c < p.C; ++c) {
                if (gm) st2(gm + (long long)c * p.HW, f2(0.f));
                if (gk) st2(gk + (long long)c * p.HW, f2(0.f));
            }
            continue;
        }
        n_valid += (v0 ? 1u : 0u) + (v1 ? 1u : 0u);
        f2 a[CP];
#pragma unroll
        for (int c = 0; c < CP; ++c) a[c] = (EXACT || c < p.C) ? ldg_stream2(base + (long long)c * p.HW) : f2(0.f);
        const int y0 = (int)tg.x, y1 = (int)tg.y;
        if (p.want_mse) {
            f2 a0(0.f), s2(0.f), ay(0.f);
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                a0 += a[c];
                s2 = fma2(a[c], a[c], s2);
                ay = f2(c == y0 ? a[c].v.x : ay.v.x, c == y1 ? a[c].v.y : ay.v.y);
            }
            // sum p^2 = s2 / D^2 and the variance term (a0^2 - s2) / G in closed form; sum (y - p)^2 per class (it cancels
            // at confident pixels)
            const f2 D = a0 + p.eps_mse, invD(1.0f / D.v.x, 1.0f / D.v.y);
            const f2 G = fma2(a0, a0, p.eps_mse) * (a0 + 1.0f), invG(1.0f / G.v.x, 1.0f / G.v.y);
            const f2 N = fma2(a0, a0, -s2);
            const f2 sp2 = s2 * invD * invD;
            const f2 Gp = fma2(a0 * 2.0f, a0 + 1.0f, fma2(a0, a0, p.eps_mse));
            const f2 common = fma2((fma2(ay, invD, -sp2)) * 2.0f, invD, fma2(a0 * 2.0f, invG, -(N * Gp * invG * invG)));
            const f2 m2invD = invD * -2.0f, m2invG = invG * -2.0f;
            f2 sq(0.f);
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                if (EXACT || c < p.C) {
                    const f2 d = f2(c == y0 ? 1.0f : 0.0f, c == y1 ? 1.0f : 0.0f) - a[c] * invD;
                    sq = fma2(d, d, sq);
                    if (gm) {
                        const f2 o = fma2(d, m2invD, fma2(a[c], m2invG, common));
                        st2(gm + (long long)c * p.HW, f2(v0 ? o.v.x : 0.f, v1 ? o.v.y : 0.f));
                    }
                }
            }
            const f2 mse = fma2(N, invG, sq);
            if (v0) acc_mse += (double)mse.v.x;
            if (v1) acc_mse += (double)mse.v.y;
        }
        if (p.want_kl) {
            f2 s(0.f), amin(3.0e38f);
#pragma unroll
            for (int c = 0; c < CP; ++c)
                if (EXACT || c < p.C) {
                    const f2 ac = max2(f2(c == y0 ? 1.0f : a[c].v.x, c == y1 ? 1.0f : a[c].v.y), p.eps_kl);
                    s += ac;
                    amin = f2(fminf(amin.v.x, ac.v.x), fminf(amin.v.y, ac.v.y));
                }
            const LDT2 fs = ldt_pos2<false>(s);
            const f2 tail = (s - (float)p.C) * fs.tri;
            f2 kl;
            if (amin.v.x >= 1.0f && amin.v.y >= 1.0f) {
                // every a~ >= 1 (what an evidential head produces): merged polynomials, slu_packed.cuh
                f2 L(0.f), Qs(0.f);
#pragma unroll
                for (int c = 0; c < CP; ++c) {
                    if (EXACT || c < p.C) {
                        const bool t0 = c == y0, t1 = c == y1;
                        const f2 ac = max2(f2(t0 ? 1.0f : a[c].v.x, t1 ? 1.0f : a[c].v.y), p.eps_kl);
                        const f2 w = rcp2(ac);
                        L += lg2_2(ac);
                        Qs = fma2(w, kl_value_poly(w), Qs);
                        if (gk) {
                            const f2 gj = kl_grad_term(w) - tail;
                            st2(gk + (long long)c * p.HW, f2((v0 && !t0 && a[c].v.x > p.eps_kl) ? gj.v.x : 0.f,
                                                             (v1 && !t1 && a[c].v.y > p.eps_kl) ? gj.v.y : 0.f));
                        }
                    }
                }
                kl = (fma2(-fs.psi, s - (float)p.C, fs.lg) + fma2(L, -0.34657359027997264f, s - (float)p.C * KL_VALUE_CONST)) + Qs;
            } else {
                kl = fs.lg;
#pragma unroll
                for (int c = 0; c < CP; ++c) {
                    if (EXACT || c < p.C) {
                        const bool t0 = c == y0, t1 = c == y1;
                        const f2 ac = max2(f2(t0 ? 1.0f : a[c].v.x, t1 ? 1.0f : a[c].v.y), p.eps_kl);
                        const LDT2 f = ldt_pos2<false>(ac);
                        kl = kl - f.lg;
                        kl = fma2(ac - 1.0f, f.psi - fs.psi, kl);
                        if (gk) {
                            const f2 gj = fma2(ac - 1.0f, f.tri, -tail);
                            st2(gk + (long long)c * p.HW, f2((v0 && !t0 && a[c].v.x > p.eps_kl) ? gj.v.x : 0.f,
                                                             (v1 && !t1 && a[c].v.y > p.eps_kl) ? gj.v.y : 0.f));
                        }
                    }
                }
            }
            if (v0) acc_kl += (double)kl.v.x;
            if (v1) acc_kl += (double)kl.v.y;
        }
    }
    __shared__ double s_m[DL2_THREADS / 32], s_k[DL2_THREADS / 32];
    __shared__ unsigned s_n[DL2_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc_mse += __shfl_xor_sync(0xffffffffu, acc_mse, o);
        acc_kl += __shfl_xor_sync(0xffffffffu, acc_kl, o);
        n_valid += __shfl_xor_sync(0xffffffffu, n_valid, o);
    }
    if ((tid & 31) == 0) { s_m[tid >> 5] = acc_mse; s_k[tid >> 5] = acc_kl; s_n[tid >> 5] = n_valid; }
    __syncthreads();
    if (tid == 0) {
        double m = 0.0, k = 0.0, n = 0.0;
        for (int i = 0; i < DL2_THREADS / 32; ++i) { m += s_m[i]; k += s_k[i]; n += (double)s_n[i]; }
        if (p.want_mse) atomicAdd(&p.sums[0], m);
        if (p.want_kl) atomicAdd(&p.sums[1], k);
        atomicAdd(&p.sums[2], n);
    }
}

// SLU_GRID_WAVES (experiment): more, shorter CTAs dealt out by the hardware scheduler instead of one resident wave.  Measured on
// 16 scans (profiles/kernel_report_r02.md): the two-term kernel gains 8 % with 8 waves (0.125 -> 0.115 ms: its CTAs do one
// chunk each and the scheduler evens out slow SMs), the fused kernel and the evidential reduction lose 2-30 %.
static int grid_waves(int dflt) { static const int w = [] { const char* e = getenv("SLU_GRID_WAVES"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 0; }(); return w ? w : dflt; }

template <int CP>
static int launch_loss(const LossParams& p, cudaStream_t st) {
    const int sms = sm_count_current_device();
    if (sms <= 0) return fail(SLU_E_DEVICE, "no CUDA device");
    const long long chunks = (p.n_px + LOSS_THREADS - 1) / LOSS_THREADS;
    const long long cap = 6LL * sms;
    const bool packed = !g_no_packed && (p.HW & 1) == 0 && (reinterpret_cast<uintptr_t>(p.target) & 15) == 0 &&
                        ((reinterpret_cast<uintptr_t>(p.alpha) | reinterpret_cast<uintptr_t>(p.grad_mse) |
                          reinterpret_cast<uintptr_t>(p.grad_kl)) & 7) == 0;
    if (packed) {
        const long long chunks2 = ((p.n_px >> 1) + DL2_THREADS - 1) / DL2_THREADS;
        const long long cap2 = (long long)SLU_DL2_MINB * sms * grid_waves(8);
        const unsigned grid2 = (unsigned)(chunks2 < cap2 ? chunks2 : cap2);
        if (p.C == CP) dirichlet_loss_x2_kernel<CP, true><<<grid2, DL2_THREADS, 0, st>>>(p);
        else dirichlet_loss_x2_kernel<CP, false><<<grid2, DL2_THREADS, 0, st>>>(p);
        SLU_LAUNCH_CHECK("dirichlet_loss_x2_kernel");
        return 0;
    }
    if (p.C == CP) dirichlet_loss_kernel<CP, true><<<(unsigned)(chunks < cap ? chunks : cap), LOSS_THREADS, 0, st>>>(p);
    else dirichlet_loss_kernel<CP, false><<<(unsigned)(chunks < cap ? chunks : cap), LOSS_THREADS, 0, st>>>(p);
    SLU_LAUNCH_CHECK("dirichlet_loss_kernel");
    return 0;
}


// =====================================================================================================
// Fused training-step loss: head output -> alpha -> w_mse * MSE + w_kl * KL and d(loss)/d(head output),
// one read of [B,C+1,HW] + target, one write of the gradient (176 B/px, SURVEY.md 8d "L").
// Replaces the trainer's chain src/models/trainer.py:532-578 (split shape/scale logits,
// to_alpha_concentrations_from_shape_and_scale, the two loss modules, the weighted sum) and autograd's
// backward through softplus / softmax.  With p = softmax(z), s = softplus(l/T), alpha = 1 + s p + eps and
// g = d(loss)/d(alpha):   dL/dz_j = s p_j (g_j - sum_c g_c p_c),   dL/dl = (sum_c g_c p_c) sigmoid(l/T) / T.
// The mean over valid pixels needs n_valid first: a count kernel (8 B/px) runs before the main kernel,
// which reads the count from device memory, so the written gradient is final (already divided by n_valid).
// =====================================================================================================
struct FusedParams {
    const float* outputs;          // [B,C+1,HW]
    const long long* target;
    const unsigned char* keep;
    int B, C;
    long long HW, n_px;
    long long ignore[MAX_IGNORE];
    int n_ignore;
    float inv_temp, eps_alpha, eps_mse, eps_kl, w_mse, w_kl;
    double* sums;                  // [3] sum mse | sum kl | n_valid   (n_valid written by the count kernel)
    float* grad;                   // [B,C+1,HW] or NULL
    // step mode (slu_evidential_loss_step): the last CTA turns the sums into the final loss values and cleans up
    const double* count;           // [1] number of valid pixels the mean runs over (NULL: sums[2])
    float* loss4;                  // [4] total | mse | kl | n_valid, or NULL
    unsigned long long* ticket;    // CTA arrival counter inside the caller's state block, left at 0
    double* own_count;             // when the call owns the count buffer: reset to 0 by the last CTA
};

// Last CTA to arrive: loss values from the finished sums, then the state block is zeroed again, so the same buffers
// serve the next step (and every replay of a captured graph) without a memset.
__device__ __forceinline__ void fused_step_epilogue(const FusedParams& p, double n_valid) {
    __threadfence();
    const unsigned long long t = atomicAdd(p.ticket, 1ull);
    if (t != (unsigned long long)gridDim.x - 1ull) return;
    __threadfence();
    const double m = atomicAdd(&p.sums[0], 0.0), k = atomicAdd(&p.sums[1], 0.0);
    const double n = fmax(n_valid, 1.0);
    const double mse = m / n, kl = k / n;
    p.loss4[0] = (float)((double)p.w_mse * mse + (double)p.w_kl * kl);
    p.loss4[1] = (float)mse;
    p.loss4[2] = (float)kl;
    p.loss4[3] = (float)n_valid;
    p.sums[0] = 0.0; p.sums[1] = 0.0;
    *p.ticket = 0ull;
    if (p.own_count) *p.own_count = 0.0;
}

__device__ __forceinline__ bool px_valid(const FusedParams& p, long long g, long long tgt) {
    if (p.keep) return p.keep[g] != 0;
    bool v = true;
#pragma unroll
    for (int i = 0; i < MAX_IGNORE; ++i)
        if (i < p.n_ignore && tgt == p.ignore[i]) v = false;
    return v;
}

__global__ void __launch_bounds__(LOSS_THREADS) count_valid_kernel(const __grid_constant__ FusedParams p, double* __restrict__ out) {
    unsigned n = 0;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < p.n_px; g += (long long)gridDim.x * blockDim.x)
        n += px_valid(p, g, p.target[g]) ? 1u : 0u;
    n = __reduce_add_sync(0xffffffffu, n);
    __shared__ unsigned s_n[LOSS_THREADS / 32];
    if ((threadIdx.x & 31) == 0) s_n[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int i = 0; i < LOSS_THREADS / 32; ++i) t += s_n[i];
        if (t) atomicAdd(out, (double)t);
    }
}

template <int CP, bool EXACT>
__global__ void __launch_bounds__(LOSS_THREADS, LOSS_MINB) evidential_loss_fused_kernel(const __grid_constant__ FusedParams p) {
    const int tid = threadIdx.x;
    const double n_valid = p.count ? *p.count : p.sums[2];
    const float inv_n = (float)(1.0 / fmax(n_valid, 1.0));
    double acc_mse = 0.0, acc_kl = 0.0;
    const long long chunks = (p.n_px + LOSS_THREADS - 1) / LOSS_THREADS;
    for (long long ch = blockIdx.x; ch < chunks; ch += gridDim.x) {
        const long long g = ch * LOSS_THREADS + tid;
        if (g >= p.n_px) continue;
        const int b = (int)(g / p.HW);
        const long long px = g - (long long)b * p.HW;
        const long long tgt = p.target[g];
        const float* base = p.outputs + ((long long)b * (p.C + 1)) * p.HW + px;
        float* go = p.grad ? p.grad + ((long long)b * (p.C + 1)) * p.HW + px : nullptr;
        if (!px_valid(p, g, tgt)) {
            if (go) for (int c = 0; c <= p.C; ++c) go[(long long)c * p.HW] = 0.f;
            continue;
        }
        const int y = (int)tgt;
        float pr[CP], a[CP];
#pragma unroll
        for (int c = 0; c < CP; ++c) pr[c] = (EXACT || c < p.C) ? ldg_stream(base + (long long)c * p.HW) : -1.0e30f;
        const float sl = ldg_stream(base + (long long)p.C * p.HW) * p.inv_temp;
        const float scale = sl > 20.f ? sl : log1pf(expf(sl));
        const float dscale = sl > 20.f ? 1.f : __fdividef(1.f, 1.f + expf(-sl));       // d softplus
        float m = pr[0];
#pragma unroll
        for (int c = 1; c < CP; ++c) m = fmaxf(m, pr[c]);
        const float m2 = m * 1.4426950408889634f;
        float S = 0.f;
#pragma unroll
        for (int c = 0; c < CP; ++c) { pr[c] = ex2_approx(fmaf(pr[c], 1.4426950408889634f, -m2)); S += pr[c]; }
        const float invS = __frcp_rn(S);
        float a0 = 0.f, s2 = 0.f, ay = 0.f;
#pragma unroll
        for (int c = 0; c < CP; ++c) {
            pr[c] *= invS;
            a[c] = (EXACT || c < p.C) ? __fadd_rn(__fadd_rn(1.0f, __fmul_rn(scale, pr[c])), p.eps_alpha) : 0.f;
            a0 += a[c];
            s2 = fmaf(a[c], a[c], s2);
            if (c == y) ay = a[c];
        }
        // ---- MSE term (same algebra as dirichlet_loss_kernel)
        const float D = a0 + p.eps_mse, invD = 1.0f / D;
        const float G = (a0 * a0 + p.eps_mse) * (a0 + 1.0f), invG = 1.0f / G;
        float sq = 0.f, sp2 = 0.f, var = 0.f;
#pragma unroll
        for (int c = 0; c < CP; ++c) {
            if (EXACT || c < p.C) {
                const float pc = a[c] * invD;
                const float d = (c == y ? 1.0f : 0.0f) - pc;
                sq = fmaf(d, d, sq);
                sp2 = fmaf(pc, pc, sp2);
                var = fmaf(a[c] * (a0 - a[c]), invG, var);
            }
        }
        acc_mse += (double)(sq + var);
        const float N = a0 * a0 - s2;
        const float Gp = 2.0f * a0 * (a0 + 1.0f) + (a0 * a0 + p.eps_mse);
        const float common = 2.0f * (ay * invD - sp2) * invD + 2.0f * a0 * invG - N * Gp * invG * invG;
        // ---- KL term
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < CP; ++c)
            if (EXACT || c < p.C) s += fmaxf(c == y ? 1.0f : a[c], p.eps_kl);
        const LDT fs = ldt_pos(s);
        float kl = fs.lg;
        const float tail = (s - (float)p.C) * fs.tri;
        float gp_sum = 0.f;                      // sum_c g_c p_c
#pragma unroll
        for (int c = 0; c < CP; ++c) {
            if (EXACT || c < p.C) {
                const float pc = a[c] * invD;
                float gc = p.w_mse * fmaf(-2.0f * ((c == y ? 1.0f : 0.0f) - pc), invD, fmaf(-2.0f * a[c], invG, common));
                {
                    const float ac = fmaxf(c == y ? 1.0f : a[c], p.eps_kl);      // a~_y = 1: lgamma = 0, (a~ - 1) = 0
                    const LDT f = ldt_pos(ac);
                    kl -= f.lg;
                    kl = fmaf(ac - 1.0f, f.psi - fs.psi, kl);
                    if (c != y && a[c] > p.eps_kl) gc = fmaf(p.w_kl, fmaf(ac - 1.0f, f.tri, -tail), gc);
                }
                a[c] = gc * inv_n;               // a[] now holds d(loss)/d(alpha_c)
                gp_sum = fmaf(a[c], pr[c], gp_sum);
            }
        }
        acc_kl += (double)kl;
        if (go) {
#pragma unroll
            for (int c = 0; c < CP; ++c)
                if (EXACT || c < p.C) go[(long long)c * p.HW] = scale * pr[c] * (a[c] - gp_sum);
            go[(long long)p.C * p.HW] = gp_sum * dscale * p.inv_temp;
        }
    }
    __shared__ double s_m[LOSS_THREADS / 32], s_k[LOSS_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc_mse += __shfl_xor_sync(0xffffffffu, acc_mse, o);
        acc_kl += __shfl_xor_sync(0xffffffffu, acc_kl, o);
    }
    if ((tid & 31) == 0) { s_m[tid >> 5] = acc_mse; s_k[tid >> 5] = acc_kl; }
    __syncthreads();
    if (tid == 0) {
        double mm = 0.0, kk = 0.0;
        for (int i = 0; i < LOSS_THREADS / 32; ++i) { mm += s_m[i]; kk += s_k[i]; }
        atomicAdd(&p.sums[0], mm);
        atomicAdd(&p.sums[1], kk);
        if (p.loss4) fused_step_epilogue(p, n_valid);
    }
}

// ---- packed variant: a thread owns TWO adjacent pixels (slu_packed.cuh) -------------------------------------------------
// Same algebra as evidential_loss_fused_kernel, every per-class operation issued once for both pixels (FFMA2 / FMUL2 /
// FADD2); MUFU evaluations, the one-hot compares and the validity selects stay per pixel.  Needs HW even and 16-byte
// aligned target / 8-byte aligned outputs and gradient (the dispatcher checks; otherwise the scalar kernel runs).
constexpr int LOSS2_THREADS = 128;
#ifndef SLU_LOSS2_MINB
#define SLU_LOSS2_MINB 3
#endif

template <int CP, bool EXACT>
__global__ void __launch_bounds__(LOSS2_THREADS, SLU_LOSS2_MINB) evidential_loss_fused_x2_kernel(const __grid_constant__ FusedParams p) {
    const int tid = threadIdx.x;
    const double n_valid = p.count ? *p.count : p.sums[2];
    const float inv_n = (float)(1.0 / fmax(n_valid, 1.0));
    double acc_mse = 0.0, acc_kl = 0.0;
    const long long pairs = p.n_px >> 1;                       // HW is even: a pair never straddles two scans
    const long long chunks = (pairs + LOSS2_THREADS - 1) / LOSS2_THREADS;
    for (long long ch = blockIdx.x; ch < chunks; ch += gridDim.x) {
        const long long pi = ch * LOSS2_THREADS + tid;
        if (pi >= pairs) continue;
        const long long g = pi << 1;
        const int b = (int)(g / p.HW);
        const long long px = g - (long long)b * p.HW;
        const longlong2 tg = *reinterpret_cast<const longlong2*>(p.target + g);
        const bool v0 = px_valid(p, g, tg.x), v1 = px_valid(p, g + 1, tg.y);
        const float* base = p.outputs + ((long long)b * (p.C + 1)) * p.HW + px;
        float* go = p.grad ? p.grad + ((long long)b * (p.C + 1)) * p.HW + px : nullptr;
        if (!v0 && !v1) {
            if (go) for (int c = 0; c <= p.C; ++c) st2(go + (long long)c * p.HW, f2(0.f));
            continue;
        }
        const int y0 = (int)tg.x, y1 = (int)tg.y;
        f2 pr[CP], a[CP];
#pragma unroll
        for (int c = 0; c < CP; ++c) pr[c] = (EXACT || c < p.C) ? ldg_stream2(base + (long long)c * p.HW) : f2(-1.0e30f);
        const f2 sl = ldg_stream2(base + (long long)p.C * p.HW) * p.inv_temp;
        f2 scale, dscale;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float ds;
            scale[h] = softplus_fast(sl[h], ds);                                        // softplus and its derivative
            dscale[h] = ds;
        }
        f2 m = pr[0];
#pragma unroll
        for (int c = 1; c < CP; ++c) m = max2(m, pr[c]);
        const f2 m2 = m * 1.4426950408889634f;
        f2 S(0.f);
#pragma unroll
        for (int c = 0; c < CP; ++c) { pr[c] = ex2_2(fma2(pr[c], 1.4426950408889634f, -m2)); S += pr[c]; }
        const f2 invS = rcp_rn2(S);
        f2 a0(0.f), s2(0.f), ay(0.f), s(0.f);
#pragma unroll
        for (int c = 0; c < CP; ++c) {
            pr[c] = pr[c] * invS;
            a[c] = (EXACT || c < p.C) ? alpha_from_probs2(scale, pr[c], p.eps_alpha) : f2(0.f);
            a0 += a[c];
            s2 = fma2(a[c], a[c], s2);
            const bool t0 = c == y0, t1 = c == y1;
            ay = f2(t0 ? a[c].v.x : ay.v.x, t1 ? a[c].v.y : ay.v.y);
            // a~ = max(y + (1 - y) alpha, eps_kl); alpha >= 1 here and the launcher requires eps_kl <= 1, so the clamp is the identity
            if (EXACT || c < p.C) s += f2(t0 ? 1.0f : a[c].v.x, t1 ? 1.0f : a[c].v.y);
        }
        // ---- both terms in ONE class loop.  Closed forms carry everything that does not need the class index:
        //   sum p^2 = s2 / D^2,  variance term = (a0^2 - s2) / G;  only sum (y - p)^2 keeps its per-class form (it cancels
        //   at confident pixels).  KL value through the merged polynomial (slu_packed.cuh::kl_value_poly):
        //   kl = lgamma(s) - psi(s)(s - C) + sum_c f(a~_c),  sum_c f = -ln2/2 sum lg2 a~ + s - C (ln 2pi + 1)/2 + sum w u(w)
        const f2 D = a0 + p.eps_mse, invD = rcp2(D);                  // reciprocals to 2^-23 relative: inside the loss tolerance
        const f2 G = fma2(a0, a0, p.eps_mse) * (a0 + 1.0f), invG = rcp2(G);
        const f2 N = fma2(a0, a0, -s2);
        const f2 sp2 = s2 * invD * invD;
        const f2 var = N * invG;
        const f2 Gp = fma2(a0 * 2.0f, a0 + 1.0f, fma2(a0, a0, p.eps_mse));
        const f2 common = fma2((fma2(ay, invD, -sp2)) * 2.0f, invD, fma2(a0 * 2.0f, invG, -(N * Gp * invG * invG)));
        const f2 m2invD = invD * -2.0f, m2invG = invG * -2.0f;
        const float wm = p.w_mse * inv_n, wk = p.w_kl * inv_n;      // 1/n_valid folded into the term weights
        const LDT2 fs = ldt_pos2<true>(s);
        const f2 tail = (s - (float)p.C) * fs.tri;
        f2 sq(0.f), L(0.f), Qs(0.f), gp_sum(0.f);
#pragma unroll
        for (int c = 0; c < CP; ++c) {
            if (EXACT || c < p.C) {
                const bool t0 = c == y0, t1 = c == y1;
                const f2 d = f2(t0 ? 1.0f : 0.0f, t1 ? 1.0f : 0.0f) - a[c] * invD;
                sq = fma2(d, d, sq);
                const f2 gm = fma2(d, m2invD, fma2(a[c], m2invG, common));
                const f2 ac(t0 ? 1.0f : a[c].v.x, t1 ? 1.0f : a[c].v.y);              // a~_y = 1: f(1) = 0, no gradient
                const f2 w = rcp2(ac);
                L += lg2_2(ac);
                Qs = fma2(w, kl_value_poly(w), Qs);
                const f2 gA = gm * wm;
                const f2 gB = fma2(kl_grad_term(w) - tail, wk, gA);
                const f2 gc(t0 ? gA.v.x : gB.v.x, t1 ? gA.v.y : gB.v.y);
                a[c] = gc;                       // a[] now holds d(loss)/d(alpha_c), 1/n_valid included
                gp_sum = fma2(gc, pr[c], gp_sum);
            }
        }
        const f2 mse = sq + var;
        const f2 kl = (fma2(-fs.psi, s - (float)p.C, fs.lg) + fma2(L, -0.34657359027997264f, s - (float)p.C * KL_VALUE_CONST)) + Qs;
        if (v0) { acc_mse += (double)mse.v.x; acc_kl += (double)kl.v.x; }
        if (v1) { acc_mse += (double)mse.v.y; acc_kl += (double)kl.v.y; }
        if (go) {
            if (v0 && v1) {                                  // the common case: no masking of the stores
#pragma unroll
                for (int c = 0; c < CP; ++c)
                    if (EXACT || c < p.C) st2(go + (long long)c * p.HW, scale * pr[c] * (a[c] - gp_sum));
                st2(go + (long long)p.C * p.HW, gp_sum * dscale * p.inv_temp);
            } else {
#pragma unroll
                for (int c = 0; c < CP; ++c)
                    if (EXACT || c < p.C) {
                        const f2 o = scale * pr[c] * (a[c] - gp_sum);
                        st2(go + (long long)c * p.HW, f2(v0 ? o.v.x : 0.f, v1 ? o.v.y : 0.f));
                    }
                const f2 o = gp_sum * dscale * p.inv_temp;
                st2(go + (long long)p.C * p.HW, f2(v0 ? o.v.x : 0.f, v1 ? o.v.y : 0.f));
            }
        }
    }
    __shared__ double s_m[LOSS2_THREADS / 32], s_k[LOSS2_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc_mse += __shfl_xor_sync(0xffffffffu, acc_mse, o);
        acc_kl += __shfl_xor_sync(0xffffffffu, acc_kl, o);
    }
    if ((tid & 31) == 0) { s_m[tid >> 5] = acc_mse; s_k[tid >> 5] = acc_kl; }
    __syncthreads();
    if (tid == 0) {
        double mm = 0.0, kk = 0.0;
        for (int i = 0; i < LOSS2_THREADS / 32; ++i) { mm += s_m[i]; kk += s_k[i]; }
        atomicAdd(&p.sums[0], mm);
        atomicAdd(&p.sums[1], kk);
        if (p.loss4) fused_step_epilogue(p, n_valid);
    }
}

// ---- packed over CLASS pairs: one pixel per thread, the two halves of every f32x2 operation are classes 2k and 2k+1 ----
// The pixel-pair kernel above needs 80 registers for its two per-class arrays of two pixels (168 in all: 3 CTAs of 128
// threads per SM, 12 warps) and is bound by the latency of its dependent FFMA2 chains, not by issue (ncu: issue slots 51 %
// busy, 0.79 eligible warps per cycle; profiles/packed_kernels_r02_ncu_summary.txt).  Packing class pairs of ONE pixel
// keeps the packed instruction count per pixel and halves the per-thread arrays, so twice as many pixels are in flight per
// register; loads and stores stay 4-byte per thread (a warp covers 128 contiguous bytes of a class plane), which also
// lifts the even-HW / alignment conditions.  Horizontal sums over the two halves close each reduction.
constexpr int LOSSC_THREADS = 256;
#ifndef SLU_LOSSC_MINB
#define SLU_LOSSC_MINB 3
#endif

template <int CP, bool EXACT>
__global__ void __launch_bounds__(LOSSC_THREADS, SLU_LOSSC_MINB) evidential_loss_fused_cp_kernel(const __grid_constant__ FusedParams p) {
    static_assert(CP % 2 == 0, "class pairs");
    constexpr int C2 = CP / 2;
    const int tid = threadIdx.x;
    const double n_valid = p.count ? *p.count : p.sums[2];
    const float inv_n = (float)(1.0 / fmax(n_valid, 1.0));
    const float wm = p.w_mse * inv_n, wk = p.w_kl * inv_n;
    double acc_mse = 0.0, acc_kl = 0.0;
    const long long chunks = (p.n_px + LOSSC_THREADS - 1) / LOSSC_THREADS;
    for (long long ch = blockIdx.x; ch < chunks; ch += gridDim.x) {
        const long long g = ch * LOSSC_THREADS + tid;
        if (g >= p.n_px) continue;
        const int b = (int)(g / p.HW);
        const long long px = g - (long long)b * p.HW;
        const long long tgt = p.target[g];
        const float* base = p.outputs + ((long long)b * (p.C + 1)) * p.HW + px;
        float* go = p.grad ? p.grad + ((long long)b * (p.C + 1)) * p.HW + px : nullptr;
        if (!px_valid(p, g, tgt)) {
            if (go) for (int c = 0; c <= p.C; ++c) go[(long long)c * p.HW] = 0.f;
            continue;
        }
        const int y = (int)tgt;
        f2 pr[C2], a[C2];
#pragma unroll
        for (int k = 0; k < C2; ++k)
            pr[k] = f2((EXACT || 2 * k < p.C) ? ldg_stream(base + (long long)(2 * k) * p.HW) : -1.0e30f,
                       (EXACT || 2 * k + 1 < p.C) ? ldg_stream(base + (long long)(2 * k + 1) * p.HW) : -1.0e30f);
        const float sl = ldg_stream(base + (long long)p.C * p.HW) * p.inv_temp;
        float dscale;
        const float scale = softplus_fast(sl, dscale);
        f2 mm = pr[0];
#pragma unroll
        for (int k = 1; k < C2; ++k) mm = max2(mm, pr[k]);
        const float m2 = fmaxf(mm.v.x, mm.v.y) * 1.4426950408889634f;
        f2 S2(0.f);
#pragma unroll
        for (int k = 0; k < C2; ++k) { pr[k] = ex2_2(fma2(pr[k], 1.4426950408889634f, f2(-m2))); S2 += pr[k]; }
        const float invS = __frcp_rn(S2.v.x + S2.v.y);
        f2 a0v(0.f), s2v(0.f), sv(0.f);
        float ay = 0.f;
#pragma unroll
        for (int k = 0; k < C2; ++k) {
            pr[k] = pr[k] * invS;
            const bool in0 = EXACT || 2 * k < p.C, in1 = EXACT || 2 * k + 1 < p.C;
            const f2 al = alpha_from_probs2(f2(scale), pr[k], p.eps_alpha);
            a[k] = f2(in0 ? al.v.x : 0.f, in1 ? al.v.y : 0.f);
            a0v += a[k];
            s2v = fma2(a[k], a[k], s2v);
            const bool t0 = 2 * k == y, t1 = 2 * k + 1 == y;
            ay = t0 ? a[k].v.x : (t1 ? a[k].v.y : ay);
            const f2 ac(t0 ? 1.0f : a[k].v.x, t1 ? 1.0f : a[k].v.y);                // alpha >= 1 >= eps_kl: no clamp
            sv += f2(in0 ? ac.v.x : 0.f, in1 ? ac.v.y : 0.f);
        }
        const float a0 = a0v.v.x + a0v.v.y, s2 = s2v.v.x + s2v.v.y, s = sv.v.x + sv.v.y;
        // per-pixel scalars (formulas: evidential_loss_fused_x2_kernel)
        const float D = a0 + p.eps_mse, invD = rcp_fast(D);
        const float G = fmaf(a0, a0, p.eps_mse) * (a0 + 1.0f), invG = rcp_fast(G);
        const float N = fmaf(a0, a0, -s2);
        const float sp2 = s2 * invD * invD;
        const float var = N * invG;
        const float Gp = fmaf(2.0f * a0, a0 + 1.0f, fmaf(a0, a0, p.eps_mse));
        const float common = fmaf(2.0f * fmaf(ay, invD, -sp2), invD, fmaf(2.0f * a0, invG, -(N * Gp * invG * invG)));
        const float m2invD = -2.0f * invD, m2invG = -2.0f * invG;
        const LDT fs = ldt_pos(s);
        const float sC = s - (float)p.C;
        const float tail = sC * fs.tri;
        f2 sqv(0.f), Lv(0.f), Qv(0.f), gpv(0.f);
#pragma unroll
        for (int k = 0; k < C2; ++k) {
            const bool in0 = EXACT || 2 * k < p.C, in1 = EXACT || 2 * k + 1 < p.C;
            const bool t0 = 2 * k == y, t1 = 2 * k + 1 == y;
            const f2 d = f2(t0 ? 1.0f : 0.0f, t1 ? 1.0f : 0.0f) - a[k] * invD;
            const f2 gm = fma2(d, m2invD, fma2(a[k], m2invG, common));
            const f2 ac((t0 || !in0) ? 1.0f : a[k].v.x, (t1 || !in1) ? 1.0f : a[k].v.y);   // padded classes: a~ = 1 (f = 0)
            const f2 w = rcp2(ac);
            const f2 gA = gm * wm;
            const f2 gB = fma2(kl_grad_term(w) - tail, wk, gA);
            const f2 gc((in0 && !t0) ? gB.v.x : (in0 ? gA.v.x : 0.f), (in1 && !t1) ? gB.v.y : (in1 ? gA.v.y : 0.f));
            const f2 dm(in0 ? d.v.x : 0.f, in1 ? d.v.y : 0.f);
            sqv = fma2(dm, dm, sqv);
            Lv += lg2_2(ac);
            Qv = fma2(w, kl_value_poly(w), Qv);
            a[k] = gc;                           // a[] now holds d(loss)/d(alpha_c), 1/n_valid included
            gpv = fma2(gc, pr[k], gpv);
        }
        // padded classes entered Qv with w = 1 (a~ = 1): f(1) = 0 = -0 + 1 - KL_VALUE_CONST + u(1), i.e. their w u(w) term is
        // exactly what the constant part below subtracts for them when it runs over CP instead of C classes
        const float n_cls = EXACT ? (float)p.C : (float)CP;
        const float sq = sqv.v.x + sqv.v.y, L = Lv.v.x + Lv.v.y, Qs = Qv.v.x + Qv.v.y, gp_sum = gpv.v.x + gpv.v.y;
        const float s_all = EXACT ? s : s + (float)(CP - p.C);                    // padded a~ = 1 each
        const float mse = sq + var;
        const float kl = (fmaf(-fs.psi, sC, fs.lg) + fmaf(L, -0.34657359027997264f, s_all - n_cls * KL_VALUE_CONST)) + Qs;
        acc_mse += (double)mse;
        acc_kl += (double)kl;
        if (go) {
#pragma unroll
            for (int k = 0; k < C2; ++k) {
                const f2 o = pr[k] * scale * (a[k] - gp_sum);
                if (EXACT || 2 * k < p.C) go[(long long)(2 * k) * p.HW] = o.v.x;
                if (EXACT || 2 * k + 1 < p.C) go[(long long)(2 * k + 1) * p.HW] = o.v.y;
            }
            go[(long long)p.C * p.HW] = gp_sum * dscale * p.inv_temp;
        }
    }
    __shared__ double s_m[LOSSC_THREADS / 32], s_k[LOSSC_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc_mse += __shfl_xor_sync(0xffffffffu, acc_mse, o);
        acc_kl += __shfl_xor_sync(0xffffffffu, acc_kl, o);
    }
    if ((tid & 31) == 0) { s_m[tid >> 5] = acc_mse; s_k[tid >> 5] = acc_kl; }
    __syncthreads();
    if (tid == 0) {
        double mm2 = 0.0, kk = 0.0;
        for (int i = 0; i < LOSSC_THREADS / 32; ++i) { mm2 += s_m[i]; kk += s_k[i]; }
        atomicAdd(&p.sums[0], mm2);
        atomicAdd(&p.sums[1], kk);
        if (p.loss4) fused_step_epilogue(p, n_valid);
    }
}

static int g_loss_variant = [] { const char* e = getenv("SLU_LOSS_VARIANT"); return e ? atoi(e) : 0; }();   // 0 = automatic (pixel pairs, class pairs for odd / unaligned shapes), 1 = class pairs, 2 = one pixel per thread

template <int CP>
static int launch_fused(const FusedParams& p, bool precounted, cudaStream_t st) {
    const int sms = sm_count_current_device();
    if (sms <= 0) return fail(SLU_E_DEVICE, "no CUDA device");
    const long long chunks = (p.n_px + LOSS_THREADS - 1) / LOSS_THREADS;
    const long long cap = 6LL * sms;
    const unsigned grid = (unsigned)(chunks < cap ? chunks : cap);
    if (!precounted) {
        count_valid_kernel<<<grid, LOSS_THREADS, 0, st>>>(p, p.count ? const_cast<double*>(p.count) : p.sums + 2);
        SLU_LAUNCH_CHECK("count_valid_kernel");
    }
    // the packed kernels drop the a~ >= eps_kl clamp (alpha = 1 + softplus * softmax + eps >= 1): valid while eps_kl <= 1, eps_alpha >= 0
    const bool packed_math_ok = p.eps_kl <= 1.0f && p.eps_alpha >= 0.0f;
    const bool pair_ok = packed_math_ok && (p.HW & 1) == 0 && (reinterpret_cast<uintptr_t>(p.target) & 15) == 0 &&
                         (reinterpret_cast<uintptr_t>(p.outputs) & 7) == 0 && (reinterpret_cast<uintptr_t>(p.grad) & 7) == 0;
    // class-pair kernel: shapes the pixel-pair kernel cannot take (odd HW, unaligned views), or SLU_LOSS_VARIANT=1.  On
    // aligned even shapes the two measure the same (0.135 / 0.140 ms per 16 scans), the pixel-pair one is kept there.
    if (!g_no_packed && packed_math_ok && (g_loss_variant == 1 || (g_loss_variant == 0 && !pair_ok))) {
        const long long capc = (long long)SLU_LOSSC_MINB * sms;
        const long long chunksc = (p.n_px + LOSSC_THREADS - 1) / LOSSC_THREADS;
        const unsigned gridc = (unsigned)(chunksc < capc ? chunksc : capc);
        if (p.C == CP) evidential_loss_fused_cp_kernel<CP, true><<<gridc, LOSSC_THREADS, 0, st>>>(p);
        else evidential_loss_fused_cp_kernel<CP, false><<<gridc, LOSSC_THREADS, 0, st>>>(p);
        SLU_LAUNCH_CHECK("evidential_loss_fused_cp_kernel");
        return 0;
    }
    const bool packed = !g_no_packed && g_loss_variant != 2 && pair_ok;
    if (packed) {
        const long long chunks2 = ((p.n_px >> 1) + LOSS2_THREADS - 1) / LOSS2_THREADS;
        const long long cap2 = (long long)SLU_LOSS2_MINB * sms * grid_waves(1);
        const unsigned grid2 = (unsigned)(chunks2 < cap2 ? chunks2 : cap2);
        if (p.C == CP) evidential_loss_fused_x2_kernel<CP, true><<<grid2, LOSS2_THREADS, 0, st>>>(p);
        else evidential_loss_fused_x2_kernel<CP, false><<<grid2, LOSS2_THREADS, 0, st>>>(p);
        SLU_LAUNCH_CHECK("evidential_loss_fused_x2_kernel");
        return 0;
    }
    if (p.C == CP) evidential_loss_fused_kernel<CP, true><<<grid, LOSS_THREADS, 0, st>>>(p);
    else evidential_loss_fused_kernel<CP, false><<<grid, LOSS_THREADS, 0, st>>>(p);
    SLU_LAUNCH_CHECK("evidential_loss_fused_kernel");
    return 0;
}


// =====================================================================================================
// Single-term kernel for the alternative data-fit terms.  Per pixel, with a0 = sum(alpha),
// ay = alpha[target]:
//   SLU_TERM_NLL        NLLDirichletCategorical  src/losses/dirichlet_losses.py:73-119
//        L = log(a0 + e) - log(ay + e)                   dL/da_j = 1/(a0+e) - [j=y]/(ay+e)
//   SLU_TERM_DIGAMMA_CE DigammaDirichletCE       :122-167
//        L = psi(a0) - psi(ay)                           dL/da_j = psi'(a0) - [j=y] psi'(ay)
//   SLU_TERM_BRIER      BrierDirichlet           :174-220   (D = a0 + e, p = alpha/D, S2 = sum p^2,
//        s = a0 or the constant s_ref)  L = (s S2 + 1)/(s + 1) - 2 p_y + 1
//        dS2/da_j = 2 alpha_j/D^2 - 2 S2/D;  dp_y/da_j = [j=y]/D - alpha_y/D^2
// =====================================================================================================
struct TermParams {
    const float* alpha;
    const long long* target;
    const unsigned char* keep;
    int B, C;
    long long HW, n_px;
    long long ignore[MAX_IGNORE];
    int n_ignore;
    int term;
    float eps, s_ref;              // s_ref < 0: use a0
    double* sums;                  // [2] sum of per-pixel values | n_valid
    float* grad;                   // [B,C,HW] or NULL
};

template <int CP>
__global__ void __launch_bounds__(LOSS_THREADS) dirichlet_term_kernel(const __grid_constant__ TermParams p) {
    const int tid = threadIdx.x;
    double acc = 0.0;
    unsigned n_valid = 0;
    const long long chunks = (p.n_px + LOSS_THREADS - 1) / LOSS_THREADS;
    for (long long ch = blockIdx.x; ch < chunks; ch += gridDim.x) {
        const long long g = ch * LOSS_THREADS + tid;
        if (g >= p.n_px) continue;
        const int b = (int)(g / p.HW);
        const long long px = g - (long long)b * p.HW;
        const long long tgt = p.target[g];
        bool valid;
        if (p.keep) {
            valid = p.keep[g] != 0;
        } else {
            valid = true;
#pragma unroll
            for (int i = 0; i < MAX_IGNORE; ++i)
                if (i < p.n_ignore && tgt == p.ignore[i]) valid = false;
        }
        const float* base = p.alpha + ((long long)b * p.C) * p.HW + px;
        float* go = p.grad ? p.grad + ((long long)b * p.C) * p.HW + px : nullptr;
        if (!valid) {
            if (go) for (int c = 0; c < p.C; ++c) go[(long long)c * p.HW] = 0.f;
            continue;
        }
        ++n_valid;
        const int y = (int)tgt;
        float a[CP];
        float a0 = 0.f, ay = 0.f;
#pragma unroll
        for (int c = 0; c < CP; ++c) {
            a[c] = (c < p.C) ? ldg_stream(base + (long long)c * p.HW) : 0.f;
            a0 += a[c];
            if (c == y) ay = a[c];
        }
        float gA = 0.f, gB = 0.f, gC = 0.f, val = 0.f;      // dL/da_j = gA + gB [j=y] + gC a_j
        if (p.term == SLU_TERM_NLL) {
            val = logf(a0 + p.eps) - logf(ay + p.eps);
            gA = 1.0f / (a0 + p.eps);
            gB = -1.0f / (ay + p.eps);
        } else if (p.term == SLU_TERM_DIGAMMA_CE) {
            val = digamma_pos(a0) - digamma_pos(ay);
            gA = trigamma_pos(a0);
            gB = -trigamma_pos(ay);
        } else {   // SLU_TERM_BRIER
            const float D = a0 + p.eps, invD = 1.0f / D;
            float S2 = 0.f;
#pragma unroll
            for (int c = 0; c < CP; ++c) { const float pc = a[c] * invD; S2 = fmaf(pc, pc, S2); }
            const float py = ay * invD;
            const float dS2_A = -2.0f * S2 * invD, dS2_C = 2.0f * invD * invD;      // dS2/da_j = dS2_A + dS2_C a_j
            if (p.s_ref < 0.f) {
                const float q = 1.0f / (a0 + 1.0f);
                val = (a0 * S2 + 1.0f) * q - 2.0f * py + 1.0f;
                gA = (S2 - 1.0f) * q * q + a0 * q * dS2_A;
                gC = a0 * q * dS2_C;
            } else {
                const float q = p.s_ref / (p.s_ref + 1.0f);
                val = (p.s_ref * S2 + 1.0f) / (p.s_ref + 1.0f) - 2.0f * py + 1.0f;
                gA = q * dS2_A;
                gC = q * dS2_C;
            }
            gA += 2.0f * ay * invD * invD;
            gB = -2.0f * invD;
        }
        acc += (double)val;
        if (go) {
#pragma unroll
            for (int c = 0; c < CP; ++c)
                if (c < p.C) go[(long long)c * p.HW] = fmaf(gC, a[c], gA + (c == y ? gB : 0.f));
        }
    }
    __shared__ double s_v[LOSS_THREADS / 32];
    __shared__ unsigned s_n[LOSS_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        n_valid += __shfl_xor_sync(0xffffffffu, n_valid, o);
    }
    if ((tid & 31) == 0) { s_v[tid >> 5] = acc; s_n[tid >> 5] = n_valid; }
    __syncthreads();
    if (tid == 0) {
        double v = 0.0, n = 0.0;
        for (int i = 0; i < LOSS_THREADS / 32; ++i) { v += s_v[i]; n += (double)s_n[i]; }
        atomicAdd(&p.sums[0], v);
        atomicAdd(&p.sums[1], n);
    }
}

template <int CP>
static int launch_term(const TermParams& p, cudaStream_t st) {
    const int sms = sm_count_current_device();
    if (sms <= 0) return fail(SLU_E_DEVICE, "no CUDA device");
    const long long chunks = (p.n_px + LOSS_THREADS - 1) / LOSS_THREADS;
    const long long cap = 6LL * sms;
    dirichlet_term_kernel<CP><<<(unsigned)(chunks < cap ? chunks : cap), LOSS_THREADS, 0, st>>>(p);
    SLU_LAUNCH_CHECK("dirichlet_term_kernel");
    return 0;
}

// diagnostic: out[3i..3i+2] = lgamma, digamma, trigamma of in[i] through ldt_pos (accuracy tests)
__global__ void special_eval_kernel(const float* in, long long n, float* out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const LDT r = ldt_pos(in[i]);
        out[3 * i] = r.lg; out[3 * i + 1] = r.psi; out[3 * i + 2] = r.tri;
    }
}

}  // namespace slu
extern "C" int slu_dirichlet_term(const float* d_alpha, const int64_t* d_target, const uint8_t* d_keep_mask,
                                  int B, int C, int64_t HW, const int64_t* h_ignore, int n_ignore,
                                  int term, float eps, float s_ref, double* d_sums, float* d_grad, slu_stream_t stream) {
    using namespace slu;
    if (!d_alpha || !d_target || !d_sums) return fail(SLU_E_ARG, "d_alpha / d_target / d_sums is NULL");
    if (B < 1 || HW < 1) return fail(SLU_E_ARG, "B=%d HW=%lld must be >= 1", B, (long long)HW);
    if (C < 2 || C > SLU_MAX_CLASSES) return fail(SLU_E_RANGE, "C=%d outside [2,%d]", C, SLU_MAX_CLASSES);
    if (n_ignore < 0 || n_ignore > MAX_IGNORE || (n_ignore > 0 && !h_ignore)) return fail(SLU_E_RANGE, "n_ignore=%d outside [0,%d]", n_ignore, MAX_IGNORE);
    if (term != SLU_TERM_NLL && term != SLU_TERM_DIGAMMA_CE && term != SLU_TERM_BRIER) return fail(SLU_E_ARG, "unknown term %d", term);
    TermParams p{};
    p.alpha = d_alpha; p.target = reinterpret_cast<const long long*>(d_target); p.keep = d_keep_mask;
    p.B = B; p.C = C; p.HW = HW; p.n_px = (long long)B * HW;
    for (int i = 0; i < n_ignore; ++i) p.ignore[i] = h_ignore[i];
    p.n_ignore = n_ignore; p.term = term; p.eps = eps; p.s_ref = s_ref;
    p.sums = d_sums; p.grad = d_grad;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch ((C + 3) / 4 * 4) {
        case 4: return launch_term<4>(p, st);
        case 8: return launch_term<8>(p, st);
        case 12: return launch_term<12>(p, st);
        case 16: return launch_term<16>(p, st);
        case 20: return launch_term<20>(p, st);
        case 24: return launch_term<24>(p, st);
        case 28: return launch_term<28>(p, st);
        default: return launch_term<32>(p, st);
    }
}

extern "C" int slu_diag_special(const float* d_in, int64_t n, float* d_out, slu_stream_t stream) {
    using namespace slu;
    if (!d_in || !d_out || n < 1) return fail(SLU_E_ARG, "slu_diag_special: bad arguments");
    special_eval_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_in, n, d_out);
    SLU_LAUNCH_CHECK("special_eval_kernel");
    return 0;
}

extern "C" int slu_evidential_loss_fused(const float* d_outputs, const int64_t* d_target, const uint8_t* d_keep_mask,
                                         int B, int C, int64_t HW, const int64_t* h_ignore, int n_ignore,
                                         float temperature, float eps_alpha, float eps_mse, float eps_kl,
                                         float w_mse, float w_kl, int precounted, double* d_sums, float* d_grad_outputs,
                                         slu_stream_t stream) {
    using namespace slu;
    if (!d_outputs || !d_target || !d_sums) return fail(SLU_E_ARG, "d_outputs / d_target / d_sums is NULL");
    if (B < 1 || HW < 1) return fail(SLU_E_ARG, "B=%d HW=%lld must be >= 1", B, (long long)HW);
    if (C < 3 || C > SLU_MAX_CLASSES) return fail(SLU_E_RANGE, "C=%d outside [3,%d]", C, SLU_MAX_CLASSES);
    if (n_ignore < 0 || n_ignore > MAX_IGNORE || (n_ignore > 0 && !h_ignore)) return fail(SLU_E_RANGE, "n_ignore=%d outside [0,%d]", n_ignore, MAX_IGNORE);
    if (!(temperature > 0.f)) return fail(SLU_E_ARG, "temperature must be > 0");
    FusedParams p{};
    p.outputs = d_outputs; p.target = reinterpret_cast<const long long*>(d_target); p.keep = d_keep_mask;
    p.B = B; p.C = C; p.HW = HW; p.n_px = (long long)B * HW;
    for (int i = 0; i < n_ignore; ++i) p.ignore[i] = h_ignore[i];
    p.n_ignore = n_ignore;
    p.inv_temp = 1.0f / temperature; p.eps_alpha = eps_alpha; p.eps_mse = eps_mse; p.eps_kl = eps_kl;
    p.w_mse = w_mse; p.w_kl = w_kl;
    p.sums = d_sums; p.grad = d_grad_outputs;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch ((C + 3) / 4 * 4) {
        case 4: return launch_fused<4>(p, precounted != 0, st);
        case 8: return launch_fused<8>(p, precounted != 0, st);
        case 12: return launch_fused<12>(p, precounted != 0, st);
        case 16: return launch_fused<16>(p, precounted != 0, st);
        case 20: return launch_fused<20>(p, precounted != 0, st);
        case 24: return launch_fused<24>(p, precounted != 0, st);
        case 28: return launch_fused<28>(p, precounted != 0, st);
        default: return launch_fused<32>(p, precounted != 0, st);
    }
}

extern "C" int slu_dirichlet_loss(const float* d_alpha, const int64_t* d_target, const uint8_t* d_keep_mask,
                                  int B, int C, int64_t HW, const int64_t* h_ignore, int n_ignore,
                                  float eps_mse, float eps_kl, int want_mse, int want_kl,
                                  double* d_sums, float* d_grad_mse, float* d_grad_kl, slu_stream_t stream) {
    using namespace slu;
    if (!d_alpha || !d_target || !d_sums) return fail(SLU_E_ARG, "d_alpha / d_target / d_sums is NULL");
    if (B < 1 || HW < 1) return fail(SLU_E_ARG, "B=%d HW=%lld must be >= 1", B, (long long)HW);
    if (C < 2 || C > SLU_MAX_CLASSES) return fail(SLU_E_RANGE, "C=%d outside [2,%d]", C, SLU_MAX_CLASSES);
    if (n_ignore < 0 || n_ignore > MAX_IGNORE) return fail(SLU_E_RANGE, "n_ignore=%d outside [0,%d]", n_ignore, MAX_IGNORE);
    if (n_ignore > 0 && !h_ignore) return fail(SLU_E_ARG, "h_ignore is NULL");
    if (!want_mse && !want_kl) return fail(SLU_E_ARG, "no loss term requested");
    if ((d_grad_mse && !want_mse) || (d_grad_kl && !want_kl)) return fail(SLU_E_ARG, "gradient requested for a term that is off");
    LossParams p{};
    p.alpha = d_alpha; p.target = reinterpret_cast<const long long*>(d_target); p.keep = d_keep_mask;
    p.B = B; p.C = C; p.HW = HW; p.n_px = (long long)B * HW;
    for (int i = 0; i < n_ignore; ++i) p.ignore[i] = h_ignore[i];
    p.n_ignore = n_ignore;
    p.eps_mse = eps_mse; p.eps_kl = eps_kl; p.want_mse = want_mse; p.want_kl = want_kl;
    p.sums = d_sums; p.grad_mse = d_grad_mse; p.grad_kl = d_grad_kl;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch ((C + 3) / 4 * 4) {
        case 4: return launch_loss<4>(p, st);
        case 8: return launch_loss<8>(p, st);
        case 12: return launch_loss<12>(p, st);
        case 16: return launch_loss<16>(p, st);
        case 20: return launch_loss<20>(p, st);
        case 24: return launch_loss<24>(p, st);
        case 28: return launch_loss<28>(p, st);
        default: return launch_loss<32>(p, st);
    }
}

extern "C" int slu_count_valid(const int64_t* d_target, const uint8_t* d_keep_mask, int64_t n_px,
                               const int64_t* h_ignore, int n_ignore, double* d_count, slu_stream_t stream) {
    using namespace slu;
    if (!d_target || !d_count) return fail(SLU_E_ARG, "d_target / d_count is NULL");
    if (n_px < 1) return fail(SLU_E_ARG, "n_px=%lld must be >= 1", (long long)n_px);
    if (n_ignore < 0 || n_ignore > MAX_IGNORE || (n_ignore > 0 && !h_ignore)) return fail(SLU_E_RANGE, "n_ignore=%d outside [0,%d]", n_ignore, MAX_IGNORE);
    FusedParams p{};
    p.target = reinterpret_cast<const long long*>(d_target); p.keep = d_keep_mask;
    p.n_px = n_px;
    for (int i = 0; i < n_ignore; ++i) p.ignore[i] = h_ignore[i];
    p.n_ignore = n_ignore;
    const int sms = sm_count_current_device();
    if (sms <= 0) return fail(SLU_E_DEVICE, "no CUDA device");
    const long long chunks = (n_px + LOSS_THREADS - 1) / LOSS_THREADS;
    const long long cap = 6LL * sms;
    count_valid_kernel<<<(unsigned)(chunks < cap ? chunks : cap), LOSS_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p, d_count);
    SLU_LAUNCH_CHECK("count_valid_kernel");
    return 0;
}

/* The training-step form of slu_evidential_loss_fused: final loss values written by the kernel, self-cleaning state. */
extern "C" int slu_evidential_loss_step(const float* d_outputs, const int64_t* d_target, const uint8_t* d_keep_mask,
                                        int B, int C, int64_t HW, const int64_t* h_ignore, int n_ignore,
                                        float temperature, float eps_alpha, float eps_mse, float eps_kl,
                                        float w_mse, float w_kl, int precounted, double* d_count, double* d_state,
                                        float* d_loss4, float* d_grad_outputs, slu_stream_t stream) {
    using namespace slu;
    if (!d_outputs || !d_target || !d_count || !d_state || !d_loss4) return fail(SLU_E_ARG, "d_outputs / d_target / d_count / d_state / d_loss4 is NULL");
    if (B < 1 || HW < 1) return fail(SLU_E_ARG, "B=%d HW=%lld must be >= 1", B, (long long)HW);
    if (C < 3 || C > SLU_MAX_CLASSES) return fail(SLU_E_RANGE, "C=%d outside [3,%d]", C, SLU_MAX_CLASSES);
    if (n_ignore < 0 || n_ignore > MAX_IGNORE || (n_ignore > 0 && !h_ignore)) return fail(SLU_E_RANGE, "n_ignore=%d outside [0,%d]", n_ignore, MAX_IGNORE);
    if (!(temperature > 0.f)) return fail(SLU_E_ARG, "temperature must be > 0");
    if ((reinterpret_cast<uintptr_t>(d_state) & 7) != 0) return fail(SLU_E_ALIGN, "d_state not 8-byte aligned");
    FusedParams p{};
    p.outputs = d_outputs; p.target = reinterpret_cast<const long long*>(d_target); p.keep = d_keep_mask;
    p.B = B; p.C = C; p.HW = HW; p.n_px = (long long)B * HW;
    for (int i = 0; i < n_ignore; ++i) p.ignore[i] = h_ignore[i];
    p.n_ignore = n_ignore;
    p.inv_temp = 1.0f / temperature; p.eps_alpha = eps_alpha; p.eps_mse = eps_mse; p.eps_kl = eps_kl;
    p.w_mse = w_mse; p.w_kl = w_kl;
    p.sums = d_state; p.grad = d_grad_outputs;
    p.count = d_count; p.loss4 = d_loss4;
    p.ticket = reinterpret_cast<unsigned long long*>(d_state + 2);
    p.own_count = precounted ? nullptr : d_count;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch ((C + 3) / 4 * 4) {
        case 4: return launch_fused<4>(p, precounted != 0, st);
        case 8: return launch_fused<8>(p, precounted != 0, st);
        case 12: return launch_fused<12>(p, precounted != 0, st);
        case 16: return launch_fused<16>(p, precounted != 0, st);
        case 20: return launch_fused<20>(p, precounted != 0, st);
        case 24: return launch_fused<24>(p, precounted != 0, st);
        case 28: return launch_fused<28>(p, precounted != 0, st);
        default: return launch_fused<32>(p, precounted != 0, st);
    }
}

/* A/B switch (tests, profiles): 1 = the loss kernels always run one pixel per thread (no packed f32x2 variant). */
extern "C" int slu_debug_no_packed_loss(int on) {
    const int prev = slu::g_no_packed;
    if (on >= 0) slu::g_no_packed = on ? 1 : 0;
    return prev;
}
