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
#include <math.h>
#include <type_traits>
#include "slu_common.cuh"

namespace slu {

constexpr int HIST_THREADS = 256;
constexpr int HIST_UNROLL = 4;
constexpr int HIST_SMEM_CELLS = 64 * 64;     // confusion matrices up to C=64 live in shared memory

// IT = index type of the prediction / label maps: long long (the reference's int64 tensors) or int (12 B/px instead of 20)
struct HistParams {
    const void* pred;
    const void* labels;
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

template <bool SMEM_CM, typename IT>
__global__ void __launch_bounds__(HIST_THREADS) confusion_ece_kernel(const __grid_constant__ HistParams p) {
    const IT* __restrict__ g_pred = reinterpret_cast<const IT*>(p.pred);
    const IT* __restrict__ g_lab = reinterpret_cast<const IT*>(p.labels);
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
            pr[u] = in ? (long long)__ldg(g_pred + i) : -1;
            lb[u] = in ? (long long)__ldg(g_lab + i) : -1;
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

// ---- streaming variant (the default whenever the inputs are 16-byte aligned and the private cells fit) ----------
// The kernel above spends its time in warp collectives (two __match_any_sync, two REDUX and three ballots per element)
// and in shared-memory atomics on 15 hot reliability cells; on random inputs it reaches 0.12 of the HBM roofline.
// Here the inner loop has no collective and no atomic on the reliability bins:
//   * a thread owns 4 CONSECUTIVE pixels per step (two 16-byte loads each of pred / labels, one of conf) and two steps
//     are in flight (160 B per thread);
//   * confusion cells: runs of equal keys inside the thread's 4 pixels are combined first (label maps are spatially
//     coherent), then one shared-memory atomic per run into the warp's PRIVATE copy of the matrix;
//   * reliability bins: every thread has private cells in shared memory, laid out [bin][thread] (bank = lane, so plain
//     conflict-free load/modify/store): a packed u32 (n << 16 | n_correct) and a u64 fixed-point confidence sum; runs
//     of equal bins are combined in registers first.  Packed counts hold 65535 pixels per thread and bin, which the host
//     guarantees by bounding the pixels per launch.
//   * one reduction per CTA at the end: warp shuffles over the private cells, one global atomic per non-empty cell.
// Counts are integers, so the result is bit-identical to the kernel above (tests run both).
constexpr int V2_THREADS = 256;
constexpr int V2_WARPS = V2_THREADS / 32;
constexpr int V2_PX = 4;                       // pixels per thread per step
constexpr int V2_STEPS = 2;                    // steps in flight
constexpr long long V2_MAX_PX_PER_THREAD = 32768;

struct V2Layout { int cm_copies; int cm_bytes; int nc_bytes; int sum_bytes; int total; };
static V2Layout v2_layout(int cells, int n_bins, bool want_cm, bool want_bins) {
    V2Layout L{};
    if (want_cm) {
        L.cm_copies = (cells * 4 * V2_WARPS <= 32 * 1024) ? V2_WARPS : 1;
        L.cm_bytes = (cells * 4 * L.cm_copies + 15) / 16 * 16;
    }
    if (want_bins) {
        L.nc_bytes = (n_bins + 1) * V2_THREADS * 4;          // + one spare row for pixels that do not count
        L.sum_bytes = (n_bins + 1) * V2_THREADS * 8;
    }
    L.total = L.cm_bytes + L.nc_bytes + L.sum_bytes + (SLU_MAX_BINS + 1) * 4 + 16;
    return L;
}

template <typename IT> struct Load4;
template <> struct Load4<long long> {
    static __device__ __forceinline__ void ld(const long long* base, long long g, long long* o) {
        const longlong2* q = reinterpret_cast<const longlong2*>(base);
        const longlong2 a = __ldcs(q + 2 * g), b = __ldcs(q + 2 * g + 1);
        o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
    }
};
template <> struct Load4<int> {
    static __device__ __forceinline__ void ld(const int* base, long long g, int* o) {
        const int4 a = __ldcs(reinterpret_cast<const int4*>(base) + g);
        o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
    }
};

template <typename IT>
__global__ void __launch_bounds__(V2_THREADS) confusion_ece_stream_kernel(const __grid_constant__ HistParams p, int cm_copies,
                                                                         int cm_bytes, long long first_px, long long n_px) {
    typedef typename std::conditional<sizeof(IT) == 8, unsigned long long, unsigned>::type UT;
    extern __shared__ __align__(16) unsigned char v2_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cells = p.C * p.C;
    unsigned* cm = reinterpret_cast<unsigned*>(v2_smem);
    unsigned long long* bsum = reinterpret_cast<unsigned long long*>(v2_smem + cm_bytes);
    unsigned* bnc = reinterpret_cast<unsigned*>(v2_smem + cm_bytes + (p.bins ? (p.n_bins + 1) * V2_THREADS * 8 : 0));
    float* edges = reinterpret_cast<float*>(v2_smem + cm_bytes + (p.bins ? (p.n_bins + 1) * V2_THREADS * 12 : 0));
    if (p.confmat)
        for (int i = tid; i < cells * cm_copies; i += V2_THREADS) cm[i] = 0;
    if (p.bins) {
        for (int i = tid; i < (p.n_bins + 1) * V2_THREADS; i += V2_THREADS) { bnc[i] = 0; bsum[i] = 0ull; }
        for (int i = tid; i <= p.n_bins; i += V2_THREADS) edges[i] = p.edges[i];
    }
    __syncthreads();
    unsigned* my_cm = cm + (cm_copies > 1 ? warp * cells : 0);
    const float nb_f = (float)p.n_bins;
    const float e_first = p.bins ? p.edges[0] : 0.f, e_last = p.bins ? p.edges[p.n_bins] : 0.f;

    const IT* pred_b = reinterpret_cast<const IT*>(p.pred) + first_px;
    const IT* lab_b = reinterpret_cast<const IT*>(p.labels) + first_px;
    const IT ignore = (IT)p.ignore;
    const float4* conf4 = reinterpret_cast<const float4*>(p.conf ? p.conf + first_px : nullptr);
    const long long n_groups = n_px / V2_PX;                    // the host passes n_px % 4 == 0
    const long long stride = (long long)gridDim.x * V2_THREADS;
    for (long long g0 = (long long)blockIdx.x * V2_THREADS + tid; g0 < n_groups; g0 += stride * V2_STEPS) {
        IT pr[V2_STEPS][V2_PX], lb[V2_STEPS][V2_PX];
        float cf[V2_STEPS][V2_PX];
#pragma unroll
        for (int s = 0; s < V2_STEPS; ++s) {
            const long long g = g0 + s * stride;
            if (g < n_groups) {
                Load4<IT>::ld(pred_b, g, pr[s]);
                Load4<IT>::ld(lab_b, g, lb[s]);
                if (conf4) { const float4 f = __ldcs(conf4 + g); cf[s][0] = f.x; cf[s][1] = f.y; cf[s][2] = f.z; cf[s][3] = f.w; }
                else { cf[s][0] = cf[s][1] = cf[s][2] = cf[s][3] = 0.f; }
            } else {
#pragma unroll
                for (int u = 0; u < V2_PX; ++u) { pr[s][u] = -1; lb[s][u] = -1; cf[s][u] = __int_as_float(0x7fc00000); }
            }
        }
#pragma unroll
        for (int s = 0; s < V2_STEPS; ++s) {
            if (p.confmat) {
                int key[V2_PX];
#pragma unroll
                for (int u = 0; u < V2_PX; ++u) {
                    // 0 <= label < C and 0 <= pred < C (evaluator.py:49) as two unsigned compares
                    const bool ok = (UT)lb[s][u] < (UT)p.C && (UT)pr[s][u] < (UT)p.C;
                    key[u] = ok ? (int)lb[s][u] * p.C + (int)pr[s][u] : -1;
                }
                if (key[0] == key[1] && key[1] == key[2] && key[2] == key[3]) {      // the common case on real label maps
                    if (key[0] >= 0) atomicAdd(&my_cm[key[0]], 4u);
                } else {
#pragma unroll
                    for (int u = 0; u < V2_PX; ++u)
                        if (key[u] >= 0) atomicAdd(&my_cm[key[u]], 1u);
                }
            }
            if (p.bins) {
                // branch-free: pixels that do not count (NaN, outside the edges, ignored label) go to a spare row
#pragma unroll
                for (int u = 0; u < V2_PX; ++u) {
                    const float raw = cf[s][u];
                    const float c = __saturatef(raw);                               // clamp to [0,1]; NaN -> 0, dropped below
                    int k = (int)(c * nb_f);
                    k = k > p.n_bins - 1 ? p.n_bins - 1 : k;
                    const float lo = edges[k], hi = edges[k + 1];
                    // the host has checked that floor(v * n_bins) is within one bin of the truth for these edges
                    k += (c >= hi && k < p.n_bins - 1) ? 1 : 0;
                    k -= (c < lo && k > 0) ? 1 : 0;
                    const bool ok = raw == raw && c >= e_first && c <= e_last && !(p.has_ignore && lb[s][u] == ignore);
                    const int cell = (ok ? k : p.n_bins) * V2_THREADS + tid;
                    bnc[cell] += 0x10000u + (pr[s][u] == lb[s][u] ? 1u : 0u);
                    bsum[cell] += __float2ull_rn(c * 4294967296.0f);
                }
            }
        }
    }
    __syncthreads();
    if (p.confmat) {
        for (int i = tid; i < cells; i += V2_THREADS) {
            unsigned v = 0;
            for (int k = 0; k < cm_copies; ++k) v += cm[k * cells + i];
            if (v) atomicAdd(&p.confmat[i], (unsigned long long)v);
        }
    }
    if (p.bins) {
        // warp w reduces bins w, w + 8, ...: 256 private cells each, 8 per lane
        for (int b = warp; b < p.n_bins; b += V2_WARPS) {
            unsigned n = 0, c = 0;
            unsigned long long sfx = 0ull;
#pragma unroll
            for (int k = 0; k < V2_THREADS / 32; ++k) {
                const unsigned v = bnc[b * V2_THREADS + k * 32 + lane];
                n += v >> 16; c += v & 0xffffu;
                sfx += bsum[b * V2_THREADS + k * 32 + lane];
            }
            n = __reduce_add_sync(0xffffffffu, n);
            c = __reduce_add_sync(0xffffffffu, c);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sfx += __shfl_xor_sync(0xffffffffu, sfx, o);
            if (lane == 0 && n) {
                atomicAdd(&p.bins[b], (unsigned long long)n);
                if (c) atomicAdd(&p.bins[p.n_bins + b], (unsigned long long)c);
                atomicAdd(&p.bins[2 * p.n_bins + b], sfx);
            }
        }
    }
}

static int g_hist_force_generic = 0;


// ---- error/score histogram: the sufficient statistic of AUROC, risk-coverage and accuracy-vs-uncertainty ----
// Replaces the per-pixel host arrays of AUROCAggregator (src/metrics/auroc.py:101-141), UncertaintyAccuracyAggregator
// (src/models/evaluator.py:660-701) and UncertaintyAggregator (src/metrics/aurc.py:273-304): every valid pixel adds 1 to
// hist[is_error][floor(clamp(score,0,1) * M)].  M is large (default 60000), so the histogram lives in global memory
// (2*M int64 ~ 1 MB, L2-resident) and is updated with fire-and-forget RED atomics; 20 B/px read.
// Hybrid bins (M = 2^20): scores pile up at BOTH ends of [0,1] (confident pixels near 0, near-uniform ones near 1), where a
// uniform grid ties thousands of pixels in one bin.  Monotone key: s >= 1/16 -> floor(s * 2^20) (step 9.5e-7 up to 1);
// s < 1/16 -> 2048 mantissa bins per binary octave over 32 octaves down to 2^-36 (relative step 4.9e-4), below that bin 0.
// 65536 + 983040 = 2^20 bins in all.
__device__ __forceinline__ int hybrid_score_bin(float s) {
    if (s >= 0.0625f) {
        const int k = (int)((double)s * 1048576.0);
        return k > 1048575 ? 1048575 : k;
    }
    const unsigned bits = __float_as_uint(s);
    const int e = (int)(bits >> 23) - 91;                 // biased exponent of 2^-36 is 91
    return e < 0 ? 0 : e * 2048 + (int)((bits >> 12) & 0x7ffu);
}

template <bool HYBRID>
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
        int bin;
        if (HYBRID) {
            bin = hybrid_score_bin(fminf(fmaxf(s, 0.f), 1.f));
        } else {
            const double c = (double)fminf(fmaxf(s, 0.f), 1.f) * (double)M;   // exact product: bin = #{k/M <= s} - 1
            bin = (int)c;
            bin = bin > M - 1 ? M - 1 : bin;
        }
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

namespace slu {
static int score_hist_impl(const float* d_score, const int64_t* d_pred, const int64_t* d_labels, int64_t n,
                           int n_score_bins, const int64_t* h_ignore, int n_ignore, int64_t* d_hist, slu_stream_t stream, bool hybrid);
}
extern "C" int slu_score_hist(const float* d_score, const int64_t* d_pred, const int64_t* d_labels, int64_t n,
                              int n_score_bins, const int64_t* h_ignore, int n_ignore, int64_t* d_hist, slu_stream_t stream) {
    return slu::score_hist_impl(d_score, d_pred, d_labels, n, n_score_bins, h_ignore, n_ignore, d_hist, stream, false);
}
/* Same accumulation into the 2^20 HYBRID bins (uniform above 1/16, 2048 bins per binary octave below): d_hist [2, 2^20]. */
extern "C" int slu_score_hist_hybrid(const float* d_score, const int64_t* d_pred, const int64_t* d_labels, int64_t n,
                                     const int64_t* h_ignore, int n_ignore, int64_t* d_hist, slu_stream_t stream) {
    return slu::score_hist_impl(d_score, d_pred, d_labels, n, 1 << 20, h_ignore, n_ignore, d_hist, stream, true);
}
static int slu::score_hist_impl(const float* d_score, const int64_t* d_pred, const int64_t* d_labels, int64_t n,
                                int n_score_bins, const int64_t* h_ignore, int n_ignore, int64_t* d_hist, slu_stream_t stream, bool hybrid) {
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
    if (hybrid)
        score_hist_kernel<true><<<(unsigned)(want < cap ? want : cap), HIST_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
            d_score, reinterpret_cast<const long long*>(d_pred), reinterpret_cast<const long long*>(d_labels), n, n_score_bins,
            n_ignore, ig[0], ig[1], ig[2], ig[3], reinterpret_cast<unsigned long long*>(d_hist));
    else
        score_hist_kernel<false><<<(unsigned)(want < cap ? want : cap), HIST_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
            d_score, reinterpret_cast<const long long*>(d_pred), reinterpret_cast<const long long*>(d_labels), n, n_score_bins,
            n_ignore, ig[0], ig[1], ig[2], ig[3], reinterpret_cast<unsigned long long*>(d_hist));
    SLU_LAUNCH_CHECK("score_hist_kernel");
    return 0;
}

namespace slu {
template <typename IT>
static int confusion_ece_impl(const IT* d_pred, const IT* d_labels, const float* d_conf,
                              int64_t n, int C, int has_ignore, int64_t ignore,
                              int n_bins, const float* h_edges,
                              int64_t* d_confmat, int64_t* d_ece_bins, slu_stream_t stream) {
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
    if (sizeof(IT) == 4 && has_ignore && (ignore < INT32_MIN || ignore > INT32_MAX)) has_ignore = 0;   // no int32 label can match
    HistParams p{};
    p.pred = d_pred;
    p.labels = d_labels;
    p.conf = d_conf;
    p.n = n; p.C = C; p.has_ignore = has_ignore; p.ignore = ignore;
    p.n_bins = d_ece_bins ? n_bins : 0;
    for (int i = 0; i <= p.n_bins && d_ece_bins; ++i) p.edges[i] = h_edges[i];
    p.confmat = reinterpret_cast<unsigned long long*>(d_confmat);
    p.bins = reinterpret_cast<unsigned long long*>(d_ece_bins);
    const int sms = sm_count_current_device();
    if (sms <= 0) return fail(SLU_E_DEVICE, "no CUDA device");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // streaming kernel: 16-byte aligned inputs, private cells within the shared-memory budget
    const V2Layout L = v2_layout(C * C, p.n_bins, p.confmat != nullptr, p.bins != nullptr);
    const bool aligned = ((reinterpret_cast<uintptr_t>(d_pred) | reinterpret_cast<uintptr_t>(d_labels) |
                           reinterpret_cast<uintptr_t>(d_conf)) & 15) == 0;
    long long done = 0;
    if (!g_hist_force_generic && aligned && L.total <= 100 * 1024 && n >= 4 * V2_PX * V2_THREADS &&
        (!p.bins || one_step_bin_search_ok(p.edges, p.n_bins))) {
        static bool attr_set[64] = {};
        int dev = 0;
        SLU_CUDA(cudaGetDevice(&dev));
        if (dev < 64 && !attr_set[dev]) {
            SLU_CUDA(cudaFuncSetAttribute(confusion_ece_stream_kernel<IT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            attr_set[dev] = true;
        }
        const int per_sm = L.total <= 56 * 1024 ? 4 : (L.total <= 75 * 1024 ? 3 : 2);
        const long long n4 = n / V2_PX * V2_PX;
        while (done < n4) {
            const long long groups_cap = (long long)per_sm * sms * V2_THREADS * (V2_MAX_PX_PER_THREAD / V2_PX);
            long long seg = n4 - done;
            if (seg / V2_PX > groups_cap) seg = groups_cap * V2_PX;
            const long long want = (seg / V2_PX + V2_THREADS * V2_STEPS - 1) / (V2_THREADS * V2_STEPS);
            const long long cap = (long long)per_sm * sms;
            confusion_ece_stream_kernel<IT><<<(unsigned)(want < cap ? want : cap), V2_THREADS, L.total, st>>>(p, L.cm_copies, L.cm_bytes, done, seg);
            SLU_LAUNCH_CHECK("confusion_ece_stream_kernel");
            done += seg;
        }
        if (done == n) return 0;
        p.pred = d_pred + done; p.labels = d_labels + done; if (p.conf) p.conf += done; p.n = n - done;      // < 4 trailing pixels
    }
    const long long chunk = (long long)HIST_THREADS * HIST_UNROLL;
    const long long want = (p.n + chunk - 1) / chunk;
    const long long cap = 8LL * sms;                       // 8 resident CTAs of 256 threads per SM
    const int grid = (int)(want < cap ? want : cap);
    if (C * C <= HIST_SMEM_CELLS)
        confusion_ece_kernel<true, IT><<<grid, HIST_THREADS, 0, st>>>(p);
    else
        confusion_ece_kernel<false, IT><<<grid, HIST_THREADS, 0, st>>>(p);
    SLU_LAUNCH_CHECK("confusion_ece_kernel");
    return 0;
}
}  // namespace slu

extern "C" int slu_confusion_ece(const int64_t* d_pred, const int64_t* d_labels, const float* d_conf,
                                 int64_t n, int C, int has_ignore, int64_t ignore,
                                 int n_bins, const float* h_edges,
                                 int64_t* d_confmat, int64_t* d_ece_bins, slu_stream_t stream) {
    return slu::confusion_ece_impl<long long>(reinterpret_cast<const long long*>(d_pred), reinterpret_cast<const long long*>(d_labels),
                                              d_conf, n, C, has_ignore, ignore, n_bins, h_edges, d_confmat, d_ece_bins, stream);
}

/* int32 prediction / label maps: 12 B/px instead of 20 and 32-bit validity tests; same counters, bit-identical results */
extern "C" int slu_confusion_ece_i32(const int32_t* d_pred, const int32_t* d_labels, const float* d_conf,
                                     int64_t n, int C, int has_ignore, int64_t ignore,
                                     int n_bins, const float* h_edges,
                                     int64_t* d_confmat, int64_t* d_ece_bins, slu_stream_t stream) {
    return slu::confusion_ece_impl<int>(d_pred, d_labels, d_conf, n, C, has_ignore, ignore, n_bins, h_edges, d_confmat, d_ece_bins, stream);
}

/* A/B switch for tests and profiles: 1 = always use the generic (warp-aggregated) histogram kernel. */
extern "C" int slu_debug_hist_generic(int on) {
    slu::g_hist_force_generic = on ? 1 : 0;
    return 0;
}
