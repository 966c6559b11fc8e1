// slu_common.cuh -- shared host/device helpers of libslu (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "slu.h"

namespace slu {

// ---- host: error plumbing (thread-local message behind slu_last_error) -------------------
int fail(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
int sm_count_current_device();
void note_launch();                 // counts kernel launches made by the library (slu_launch_count)

#define SLU_CUDA(call)                                            \
    do {                                                          \
        cudaError_t e__ = (call);                                 \
        if (e__ != cudaSuccess) return slu::cuda_fail(e__, #call); \
    } while (0)

#define SLU_LAUNCH_CHECK(what)                                      \
    do {                                                            \
        slu::note_launch();                                         \
        cudaError_t e__ = cudaGetLastError();                       \
        if (e__ != cudaSuccess) return slu::cuda_fail(e__, what);   \
    } while (0)

// floor(v * n_bins) is within one bin of the truth for these edges (true for (near-)uniform edges such as the
// reference's linspace(0,1,n_bins+1)): the kernels may then use find_bin_fast.  fl(v * n_bins) is monotone in v, so
// checking each bin's two ends (its lower edge and the float just below its upper edge) covers every v inside it.
bool one_step_bin_search_ok(const float* edges, int n_bins);

#ifdef __CUDACC__
// ---- device: PTX wrappers ---------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP); bytes % 16 == 0, both 16 B aligned.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// streaming 4-byte load: read-only path, do not keep in L1
__device__ __forceinline__ float ldg_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// ---- device: warp-aggregated shared-memory histogram updates -------------------------------
// One shared atomic per DISTINCT key in the warp instead of one per lane.
__device__ __forceinline__ void warp_hist_add(unsigned* hist, int key, bool active) {
    const unsigned act = __ballot_sync(0xffffffffu, active);
    if (!active) return;
    const unsigned peers = __match_any_sync(act, key);
    if ((threadIdx.x & 31) == (__ffs(peers) - 1)) atomicAdd(&hist[key], (unsigned)__popc(peers));
}

// Reliability bins: n, n_correct (u32 per CTA) and sum(conf) as 2^-32 fixed point (u64).
// Lanes are grouped by bin with one __match_any_sync; each group reduces its confidences with REDUX
// on its own lane mask (the hardware runs one REDUX per distinct group) and its leader issues the three
// shared-memory atomics.  All 32 lanes must call it.
__device__ __forceinline__ void warp_bins_add(unsigned* bin_n, unsigned* bin_c, unsigned long long* bin_s,
                                              int bin, bool correct, float conf, bool active) {
    const unsigned act = __ballot_sync(0xffffffffu, active);
    const unsigned okm = __ballot_sync(0xffffffffu, active && correct);
    if (!active) return;
    // conf in [0,1]: conf * 2^32 is exact for conf >= 2^-9 and fits 33 bits
    const unsigned long long fx = __float2ull_rn(conf * 4294967296.0f);
    const unsigned peers = __match_any_sync(act, bin);
    const unsigned lo = __reduce_add_sync(peers, (unsigned)(fx & 0xffffu));
    const unsigned hi = __reduce_add_sync(peers, (unsigned)(fx >> 16));
    if ((threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) {
        atomicAdd(&bin_n[bin], (unsigned)__popc(peers));
        const unsigned ok = peers & okm;
        if (ok) atomicAdd(&bin_c[bin], (unsigned)__popc(ok));
        atomicAdd(&bin_s[bin], ((unsigned long long)hi << 16) + lo);
    }
}

// Bin of a confidence under np.histogram's rule: edges[k] <= v < edges[k+1], last bin closed.
// Returns -1 when v is NaN or outside [edges[0], edges[n_bins]].
__device__ __forceinline__ int find_bin(const float* edges, int n_bins, float v) {
    if (!(v >= edges[0]) || !(v <= edges[n_bins])) return -1;
    int k = (int)(v * (float)n_bins);
    k = k < 0 ? 0 : (k > n_bins - 1 ? n_bins - 1 : k);
    while (k > 0 && v < edges[k]) --k;
    while (k < n_bins - 1 && v >= edges[k + 1]) ++k;
    return k;
}
// Same result as find_bin for edges that passed one_step_bin_search_ok(): no loop, two edge loads.  v in [0,1].
__device__ __forceinline__ int find_bin_fast(const float* edges, int n_bins, float v) {
    int k = (int)(v * (float)n_bins);
    k = k > n_bins - 1 ? n_bins - 1 : k;
    const float lo = edges[k], hi = edges[k + 1];
    k += (v >= hi && k < n_bins - 1) ? 1 : 0;
    k -= (v < lo && k > 0) ? 1 : 0;
    return (v >= edges[0] && v <= edges[n_bins]) ? k : -1;
}

// ---- per-CTA histogram scratch updated with plain shared-memory atomics ------------------------------------------
// (single-sample and evidential kernels: one epilogue per pixel, so the warp collectives of warp_bins_add would cost as
// much as the pixel's arithmetic.)  Counts use ATOMS.POPC.INC.  64-bit shared atomics are CAS loops on sm_100, so
// sum(conf * 2^32) is kept as two 32-bit halves (low 16 bits | the rest, < 2^17 per pixel): a CTA must flush before a
// bin has received ATOMIC_HIST_MAX_PX pixels.
constexpr int ATOMIC_HIST_MAX_PX = 16384;            // 2^17 * 2^14 < 2^32
struct AtomicHist {
    unsigned confmat[SLU_MAX_CLASSES * SLU_MAX_CLASSES];
    unsigned bin_n[SLU_MAX_BINS], bin_c[SLU_MAX_BINS], bin_lo[SLU_MAX_BINS], bin_hi[SLU_MAX_BINS];
    float edges[SLU_MAX_BINS + 1];
};
__device__ __forceinline__ void atomic_hist_zero(AtomicHist& hs, int C, int tid, int nthreads) {
    for (int i = tid; i < C * C; i += nthreads) hs.confmat[i] = 0;
    for (int i = tid; i < SLU_MAX_BINS; i += nthreads) { hs.bin_n[i] = 0; hs.bin_c[i] = 0; hs.bin_lo[i] = 0; hs.bin_hi[i] = 0; }
}
__device__ __forceinline__ void atomic_hist_flush(AtomicHist& hs, int C, int n_bins, unsigned long long* confmat,
                                                  unsigned long long* bins, int tid, int nthreads) {
    if (confmat)
        for (int i = tid; i < C * C; i += nthreads)
            if (hs.confmat[i]) atomicAdd(&confmat[i], (unsigned long long)hs.confmat[i]);
    if (bins)
        for (int i = tid; i < n_bins; i += nthreads)
            if (hs.bin_n[i]) {
                atomicAdd(&bins[i], (unsigned long long)hs.bin_n[i]);
                if (hs.bin_c[i]) atomicAdd(&bins[n_bins + i], (unsigned long long)hs.bin_c[i]);
                atomicAdd(&bins[2 * n_bins + i], ((unsigned long long)hs.bin_hi[i] << 16) + hs.bin_lo[i]);
            }
}
// One live pixel: confusion cell (label, pred_cm) and reliability bin of `conf` with correctness pred_ece == label.
__device__ __forceinline__ void atomic_hist_add(AtomicHist& hs, int C, int n_bins, bool want_cm, bool want_bins, bool one_step,
                                                long long lab, int pred_cm, int pred_ece, float conf, bool has_ignore, long long ignore) {
    if (want_cm && (unsigned long long)lab < (unsigned long long)C) atomicAdd(&hs.confmat[(int)lab * C + pred_cm], 1u);   // evaluator.py:49
    if (want_bins) {
        const float cf = __saturatef(conf);                                   // ece.py:83 clamp_(0,1)
        int bin = one_step ? find_bin_fast(hs.edges, n_bins, cf) : find_bin(hs.edges, n_bins, cf);
        if (conf != conf || (has_ignore && lab == ignore)) bin = -1;          // NaN stays out of every bin
        if (bin >= 0) {
            const unsigned long long fx = __float2ull_rn(cf * 4294967296.0f);
            atomicAdd(&hs.bin_n[bin], 1u);
            if ((long long)pred_ece == lab) atomicAdd(&hs.bin_c[bin], 1u);
            atomicAdd(&hs.bin_lo[bin], (unsigned)(fx & 0xffffu));
            atomicAdd(&hs.bin_hi[bin], (unsigned)(fx >> 16));
        }
    }
}
#endif  // __CUDACC__

}  // namespace slu
