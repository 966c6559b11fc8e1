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
#include "slu_common.cuh"
#include "slu_special.cuh"

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
    float logC;
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

template <int CP>
static int launch_ev(const EvParams& p, cudaStream_t st) {
    const int sms = sm_count_current_device();
    if (sms <= 0) return fail(SLU_E_DEVICE, "no CUDA device");
    const long long chunks = (p.n_px + EV_THREADS - 1) / EV_THREADS;
    const long long cap = 5LL * sms;                       // 48 registers: five resident CTAs of 256 threads
    const unsigned grid = (unsigned)(chunks < cap ? chunks : cap);
    const bool exact = p.C == CP, mi = p.mi != nullptr;
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
