// slu_hist.cu -- stage 4 standalone: confusion matrix + reliability bins from reduced maps.
//
// Replaces (reference file:line): IoUEvaluator.update src/models/evaluator.py:39-53
// (bincount(t*C+p) on the CPU after a D2H copy of preds/labels) and the binning half of
// ECEAggregator (src/metrics/ece.py:75-90 keeps every valid pixel on the host, :131-140 sorts them
// three times at compute()).  Here the state is the histogram itself: C*C + 3*n_bins int64.
//
// HBM-bound: 8 (pred) + 8 (label) + 4 (conf) = 20 B/pixel read, nothing written but the counts.
// Each thread loads 4 independent elements before touching shared memory; per-CTA histograms in
// shared memory take one atomic per DISTINCT key per warp (__match_any_sync), which is what makes
// spatially coherent label maps cheap; one global atomic per non-zero cell per CTA at the end.
#include "slu_common.cuh"

namespace slu {

constexpr int HIST_THREADS = 256;
constexpr int HIST_UNROLL = 4;
constexpr int HIST_SMEM_CELLS = 64 * 64;     // confusion matrices up to C=64 live in shared memory

struct HistParams {
    const long long* pred;
    const long long* labels;
    const float* conf;
    long long n;
    int C;
    int has_ignore;
    long long ignore;
    int n_bins;
    float edges[SLU_MAX_BINS + 1];
    unsigned long long* confmat;
    unsigned long long* bins;
};

template <bool SMEM_CM>
__global__ void __launch_bounds__(HIST_THREADS) confusion_ece_kernel(const __grid_constant__ HistParams p) {
    __shared__ unsigned cm[SMEM_CM ? HIST_SMEM_CELLS : 1];
    __shared__ unsigned bin_n[SLU_MAX_BINS], bin_c[SLU_MAX_BINS];
    __shared__ unsigned long long bin_s[SLU_MAX_BINS];
    __shared__ float edges[SLU_MAX_BINS + 1];
    const int tid = threadIdx.x;
    const int cells = p.C * p.C;
    if (SMEM_CM)
        for (int i = tid; i < cells; i += HIST_THREADS) cm[i] = 0;
    for (int i = tid; i < SLU_MAX_BINS; i += HIST_THREADS) { bin_n[i] = 0; bin_c[i] = 0; bin_s[i] = 0ull; }
    for (int i = tid; i <= p.n_bins; i += HIST_THREADS) edges[i] = p.edges[i];
    __syncthreads();

    const long long chunk = (long long)HIST_THREADS * HIST_UNROLL;
    // whole warps stay in the loop together: the warp-aggregated updates need all 32 lanes
    for (long long base = (long long)blockIdx.x * chunk; base < p.n; base += (long long)gridDim.x * chunk) {
        long long pr[HIST_UNROLL], lb[HIST_UNROLL];
        float cf[HIST_UNROLL];
#pragma unroll
        for (int u = 0; u < HIST_UNROLL; ++u) {
            const long long i = base + (long long)u * HIST_THREADS + tid;
            const bool in = i < p.n;
            pr[u] = in ? __ldg(p.pred + i) : -1;
            lb[u] = in ? __ldg(p.labels + i) : -1;
            cf[u] = (in && p.conf) ? __ldg(p.conf + i) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < HIST_UNROLL; ++u) {
            const bool in = base + (long long)u * HIST_THREADS + tid < p.n;
            if (p.confmat) {
                const bool ok = in && lb[u] >= 0 && lb[u] < p.C && pr[u] >= 0 && pr[u] < p.C;   // evaluator.py:49
                const int key = ok ? (int)lb[u] * p.C + (int)pr[u] : 0;
                if (SMEM_CM) {
                    warp_hist_add(cm, key, ok);
                } else {
                    const unsigned act = __ballot_sync(0xffffffffu, ok);
                    if (ok) {
                        const unsigned peers = __match_any_sync(act, key);
                        if ((tid & 31) == (__ffs(peers) - 1)) atomicAdd(&p.confmat[key], (unsigned long long)__popc(peers));
                    }
                }
            }
            if (p.bins) {
                const float c = fminf(fmaxf(cf[u], 0.f), 1.f);
                const int bin = (cf[u] == cf[u]) ? find_bin(edges, p.n_bins, c) : -1;
                const bool ok = in && bin >= 0 && !(p.has_ignore && lb[u] == p.ignore);
                warp_bins_add(bin_n, bin_c, bin_s, bin, pr[u] == lb[u], c, ok);
            }
        }
    }
    __syncthreads();
    if (SMEM_CM && p.confmat)
        for (int i = tid; i < cells; i += HIST_THREADS)
            if (cm[i]) atomicAdd(&p.confmat[i], (unsigned long long)cm[i]);
    if (p.bins)
        for (int i = tid; i < p.n_bins; i += HIST_THREADS) {
            if (bin_n[i]) atomicAdd(&p.bins[i], (unsigned long long)bin_n[i]);
            if (bin_c[i]) atomicAdd(&p.bins[p.n_bins + i], (unsigned long long)bin_c[i]);
            if (bin_s[i]) atomicAdd(&p.bins[2 * p.n_bins + i], bin_s[i]);
        }
}

// ---- error/score histogram: the sufficient statistic of AUROC, risk-coverage and accuracy-vs-uncertainty ----
// Replaces the per-pixel host arrays of AUROCAggregator (src/metrics/auroc.py:101-141), UncertaintyAccuracyAggregator
// (src/models/evaluator.py:660-701) and UncertaintyAggregator (src/metrics/aurc.py:273-304): every valid pixel adds 1 to
// hist[is_error][floor(clamp(score,0,1) * M)].  M is large (default 60000), so the histogram lives in global memory
// (2*M int64 ~ 1 MB, L2-resident) and is updated with fire-and-forget RED atomics; 20 B/px read.
__global__ void __launch_bounds__(HIST_THREADS) score_hist_kernel(const float* __restrict__ score, const long long* __restrict__ pred,
                                                                  const long long* __restrict__ labels, long long n, int M,
                                                                  int n_ignore, long long ig0, long long ig1, long long ig2, long long ig3,
                                                                  unsigned long long* __restrict__ hist) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long lb = __ldg(labels + i), pr = __ldg(pred + i);
        const float s = __ldg(score + i);
        bool valid = s == s;
        if (n_ignore > 0 && lb == ig0) valid = false;
        if (n_ignore > 1 && lb == ig1) valid = false;
        if (n_ignore > 2 && lb == ig2) valid = false;
        if (n_ignore > 3 && lb == ig3) valid = false;
        if (!valid) continue;
        const double c = (double)fminf(fmaxf(s, 0.f), 1.f) * (double)M;       // exact product: bin = #{k/M <= s} - 1
        int bin = (int)c;
        bin = bin > M - 1 ? M - 1 : bin;
        atomicAdd(&hist[(pr != lb ? (long long)M : 0ll) + bin], 1ull);
    }
}

// ---- per-class score histogram: UncertaintyPerClassAggregator (src/models/evaluator.py:191-262) --------------------
// The reference keeps, per class, every pixel's uncertainty on the host (for box / ridgeline plots and the per-class mean).
// Here: hist[label][floor(clamp(score,0,1) * M)] += 1 and an exact per-class sum in 2^-32 fixed point.
__global__ void __launch_bounds__(HIST_THREADS) class_score_hist_kernel(const float* __restrict__ score, const long long* __restrict__ labels,
                                                                        long long n, int C, int M, unsigned long long* __restrict__ hist,
                                                                        unsigned long long* __restrict__ sum_fx) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long lb = __ldg(labels + i);
        const float s = __ldg(score + i);
        if (lb < 0 || lb >= C || s != s) continue;
        int bin = (int)((double)fminf(fmaxf(s, 0.f), 1.f) * (double)M);
        bin = bin > M - 1 ? M - 1 : bin;
        atomicAdd(&hist[lb * M + bin], 1ull);
        atomicAdd(&sum_fx[lb], __double2ull_rn((double)s * 4294967296.0));      // unclamped value, as the reference stores it
    }
}

}  // namespace slu

extern "C" int slu_class_score_hist(const float* d_score, const int64_t* d_labels, int64_t n, int C, int n_score_bins,
                                    int64_t* d_hist, int64_t* d_sum_fx, slu_stream_t stream) {
    using namespace slu;
    if (n < 0) return fail(SLU_E_ARG, "n < 0");
    if (n == 0) return 0;
    if (!d_score || !d_labels || !d_hist || !d_sum_fx) return fail(SLU_E_ARG, "NULL pointer");
    if (C < 1 || C > 4096 || n_score_bins < 1 || n_score_bins > (1 << 20)) return fail(SLU_E_RANGE, "C=%d / n_score_bins=%d unsupported", C, n_score_bins);
    const int sms = sm_count_current_device();
    if (sms <= 0) return fail(SLU_E_DEVICE, "no CUDA device");
    const long long want = (n + HIST_THREADS - 1) / HIST_THREADS;
    const long long cap = 8LL * sms;
    class_score_hist_kernel<<<(unsigned)(want < cap ? want : cap), HIST_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        d_score, reinterpret_cast<const long long*>(d_labels), n, C, n_score_bins,
        reinterpret_cast<unsigned long long*>(d_hist), reinterpret_cast<unsigned long long*>(d_sum_fx));
    SLU_LAUNCH_CHECK("class_score_hist_kernel");
    return 0;
}

extern "C" int slu_score_hist(const float* d_score, const int64_t* d_pred, const int64_t* d_labels, int64_t n,
                              int n_score_bins, const int64_t* h_ignore, int n_ignore, int64_t* d_hist, slu_stream_t stream) {
    using namespace slu;
    if (n < 0) return fail(SLU_E_ARG, "n < 0");
    if (n == 0) return 0;
    if (!d_score || !d_pred || !d_labels || !d_hist) return fail(SLU_E_ARG, "NULL pointer");
    if (n_score_bins < 1 || n_score_bins > (1 << 24)) return fail(SLU_E_RANGE, "n_score_bins=%d outside [1,2^24]", n_score_bins);
    if (n_ignore < 0 || n_ignore > 4 || (n_ignore > 0 && !h_ignore)) return fail(SLU_E_RANGE, "n_ignore=%d outside [0,4]", n_ignore);
    long long ig[4] = {0, 0, 0, 0};
    for (int i = 0; i < n_ignore; ++i) ig[i] = h_ignore[i];
    const int sms = sm_count_current_device();
    if (sms <= 0) return fail(SLU_E_DEVICE, "no CUDA device");
    const long long want = (n + HIST_THREADS - 1) / HIST_THREADS;
    const long long cap = 8LL * sms;
    score_hist_kernel<<<(unsigned)(want < cap ? want : cap), HIST_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        d_score, reinterpret_cast<const long long*>(d_pred), reinterpret_cast<const long long*>(d_labels), n, n_score_bins,
        n_ignore, ig[0], ig[1], ig[2], ig[3], reinterpret_cast<unsigned long long*>(d_hist));
    SLU_LAUNCH_CHECK("score_hist_kernel");
    return 0;
}

extern "C" int slu_confusion_ece(const int64_t* d_pred, const int64_t* d_labels, const float* d_conf,
                                 int64_t n, int C, int has_ignore, int64_t ignore,
                                 int n_bins, const float* h_edges,
                                 int64_t* d_confmat, int64_t* d_ece_bins, slu_stream_t stream) {
    using namespace slu;
    if (n < 0) return fail(SLU_E_ARG, "n=%lld < 0", (long long)n);
    if (n == 0) return 0;
    if (!d_pred || !d_labels) return fail(SLU_E_ARG, "d_pred / d_labels is NULL");
    if (C < 1 || C > 1024) return fail(SLU_E_RANGE, "C=%d outside [1,1024]", C);
    if (d_ece_bins) {
        if (!d_conf) return fail(SLU_E_ARG, "reliability bins requested without d_conf");
        if (n_bins < 1 || n_bins > SLU_MAX_BINS) return fail(SLU_E_RANGE, "n_bins=%d outside [1,%d]", n_bins, SLU_MAX_BINS);
        if (!h_edges) return fail(SLU_E_ARG, "h_edges is NULL");
        for (int i = 0; i < n_bins; ++i)
            if (!(h_edges[i] < h_edges[i + 1])) return fail(SLU_E_ARG, "bin edges must increase strictly");
    }
    if (!d_confmat && !d_ece_bins) return 0;
    HistParams p{};
    p.pred = reinterpret_cast<const long long*>(d_pred);
    p.labels = reinterpret_cast<const long long*>(d_labels);
    p.conf = d_conf;
    p.n = n; p.C = C; p.has_ignore = has_ignore; p.ignore = ignore;
    p.n_bins = d_ece_bins ? n_bins : 0;
    for (int i = 0; i <= p.n_bins && d_ece_bins; ++i) p.edges[i] = h_edges[i];
    p.confmat = reinterpret_cast<unsigned long long*>(d_confmat);
    p.bins = reinterpret_cast<unsigned long long*>(d_ece_bins);
    const int sms = sm_count_current_device();
    if (sms <= 0) return fail(SLU_E_DEVICE, "no CUDA device");
    const long long chunk = (long long)HIST_THREADS * HIST_UNROLL;
    const long long want = (n + chunk - 1) / chunk;
    const long long cap = 8LL * sms;                       // 8 resident CTAs of 256 threads per SM
    const int grid = (int)(want < cap ? want : cap);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (C * C <= HIST_SMEM_CELLS)
        confusion_ece_kernel<true><<<grid, HIST_THREADS, 0, st>>>(p);
    else
        confusion_ece_kernel<false><<<grid, HIST_THREADS, 0, st>>>(p);
    SLU_LAUNCH_CHECK("confusion_ece_kernel");
    return 0;
}
