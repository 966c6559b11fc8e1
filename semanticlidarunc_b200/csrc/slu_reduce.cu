// slu_reduce.cu -- stage 3+4: fused per-pixel reduction over T samples x C classes + evaluation
// histograms, one pass over HBM.
//
// Replaces (reference file:line): src/models/tester.py:412-471 (softmax, mean over T, argmax,
// predictive entropy, mutual information, IoUEvaluator.update, ECEAggregator.update),
// src/models/trainer.py:1170-1225 (T=1), src/utils/mc_dropout.py:121-133, src/metrics/ece.py:54-90,
// src/models/evaluator.py:39-53.
//
// Data layout.  d_in is [T,B,C,HW] fp32: for fixed (t,b,c) the HW pixels are contiguous, the two
// reduced axes (T, C) are the strided ones.  A tile is TILE consecutive pixels of one scan b; a
// tile's whole reduction (T*C values per pixel) happens in the registers of one thread per pixel, so
// every input byte is read from HBM exactly once and nothing but the 24 B/pixel of results is
// written.  HBM-bound: 4*T*C + 8 (label) bytes in, 24 bytes out per pixel; ~0.3 flop/B, no tensor
// cores (there is no contraction here).
//
// Staged kernel (default).  One producer warp streams [C x TILE] fp32 slabs (C bulk async copies
// of TILE*4 B, TMA engine, mbarrier complete_tx) through a STAGES-deep shared-memory ring (3 x 20 KB
// at C=20, three CTAs per SM), running ahead of the 8 consumer warps across (t, tile) boundaries; consumers copy their pixel's C values
// to registers, release the slot at once, and do the math.  CTAs are persistent (grid = resident
// CTAs), tiles are strided over CTAs.
//
// Per (pixel, t) with logits x_c (fast path, ~6.5 issue slots per value):
//   m = max_c x_c;  a_c = x_c*log2e - m*log2e;  e_c = 2^a_c;  S = sum e_c;  A = sum e_c*a_c
//   p_c = e_c / S;  H_t = ln S - ln2 * A / S   ( = -sum p_c ln p_c, no per-value log )
// and sum_t ln S_t is carried as a running product (one logf per 8 samples).
// The eps clamp of the reference (p.clamp_min(1e-12) inside H_t) changes H_t by at most
// C*eps*|ln eps| = 5.5e-10 for eps = 1e-12, below fp32 resolution of the result; when eps is large
// enough to matter, or the fast path's sum is not finite (-inf logits), the pixel's per-sample
// entropies are recomputed with the reference's formulas verbatim (literal_entropy_sum).  The clamp
// on p_bar is always applied literally.  The mean over T multiplies by fl(1/T) (the reference
// divides; <= 1 ulp apart) and ln(p_bar) uses lg2.approx (relative error <= 2^-22): both far inside
// the 1e-5 relative tolerance BASELINE.json states.
#include <math.h>
#include <stdlib.h>
#include "slu_common.cuh"

namespace slu {

#ifndef SLU_TILE
#define SLU_TILE 256
#endif
#ifndef SLU_STAGES
#define SLU_STAGES 3
#endif
#ifndef SLU_CTAS_PER_SM
#define SLU_CTAS_PER_SM 3
#endif
constexpr int TILE = SLU_TILE;          // pixels per tile == consumer threads per CTA
constexpr int STAGES = SLU_STAGES;      // ring depth (slabs of C*TILE*4 bytes)
// Resident CTAs per SM of the staged kernel.  Measured on B200 (profiles/reduce_variants_r01.md): the
// consumers are latency-bound chains (max -> ex2 -> sums), so 24 consumer warps/SM (3 CTAs, <= 72
// registers, no spills at C <= 20) reach 6.5 TB/s where 16 warps/SM (2 CTAs, 96 registers) stop at 5.8.
constexpr int CTAS_PER_SM = SLU_CTAS_PER_SM;
template <int CP> struct Occ { static constexpr int staged = CP <= 20 ? CTAS_PER_SM : 2; };
constexpr int NCONS_WARPS = TILE / 32;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

struct ReduceParams {
    const float* in;
    const long long* labels;
    int T, B, C;
    long long HW;
    int conf_mode;
    float eps;
    int has_ignore;
    long long ignore;
    int n_bins;
    float edges[SLU_MAX_BINS + 1];
    float* pbar;
    long long* pred;
    float* conf;
    float* hnorm;
    float* minorm;
    unsigned long long* confmat;
    unsigned long long* bins;
    int tiles_per_scan;
    long long n_tiles;
    float logC;
    float invT;
    int literal_clamp;   // apply the eps clamp per value inside H_t
    int need_ht;         // mutual information requested (per-sample entropies needed)
    int allow_single;    // T == 1 may take reduce_single_kernel (not the explicit *_direct entry point)
    int bins_one_step;   // edges passed one_step_bin_search_ok(): find_bin_fast is valid
};

// Per-CTA histogram scratch in static shared memory.
struct HistSmem {
    unsigned confmat[SLU_MAX_CLASSES * SLU_MAX_CLASSES];
    unsigned bin_n[SLU_MAX_BINS];
    unsigned bin_c[SLU_MAX_BINS];
    unsigned long long bin_s[SLU_MAX_BINS];
    float edges[SLU_MAX_BINS + 1];
};

// ---- per-pixel running state ---------------------------------------------------------------------------
// sum_t H[p_t] = sum_t ln S_t - ln2 * sum_t A_t/S_t.  The first sum is kept as a running PRODUCT of the
// S_t (S_t in [1, C], flushed through one accurate logf every 8 samples so it cannot overflow), which
// replaces T logf calls per pixel by ceil(T/8).
template <int CP>
struct Acc {
    float pbar[CP];
    float AS;      // sum_t A_t / S_t          (log2 units)
    float P;       // running product of S_t since the last flush
    float L;       // sum of ln(P) over flushes
    float EH;      // literal sum_t H[p_t]     (PROBS kind only)
};

// ---- per (pixel, t) step: x[] (one sample's C values) is folded into the accumulators --------------------
template <int CP, int KIND, bool EXACT>
__device__ __forceinline__ void sample_step(float (&x)[CP], Acc<CP>& a, const ReduceParams& p, int t) {
    if (KIND == SLU_IN_LOGITS) {
        float m = x[0];
#pragma unroll
        for (int c = 1; c < CP; ++c) m = fmaxf(m, x[c]);
        const float m2 = m * LOG2E;
        float S = 0.f, A = 0.f;
#pragma unroll
        for (int c = 0; c < CP; ++c) {
            const float arg = fmaf(x[c], LOG2E, -m2);
            const float e = ex2_approx(arg);
            x[c] = e;
            S += e;
            A = fmaf(e, arg, A);
        }
        const float inv = __frcp_rn(S);
#pragma unroll
        for (int c = 0; c < CP; ++c) a.pbar[c] = fmaf(x[c], inv, a.pbar[c]);
        a.AS = fmaf(A, inv, a.AS);
        a.P *= S;
        if ((t & 7) == 7) { a.L += logf(a.P); a.P = 1.f; }
    } else {
        if (KIND == SLU_IN_ALPHA) {       // p = alpha / (alpha0 + eps)   (src/metrics/ece.py:57-58)
            float a0 = 0.f;
#pragma unroll
            for (int c = 0; c < CP; ++c) a0 += x[c];
            const float d = a0 + p.eps;
#pragma unroll
            for (int c = 0; c < CP; ++c) x[c] = __fdiv_rn(x[c], d);
        }
        if (p.need_ht) {
            float Ht = 0.f;
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                if (EXACT || c < p.C) {
                    const float pc = fmaxf(x[c], p.eps);
                    Ht = fmaf(-pc, logf(pc), Ht);
                }
            }
            a.EH += Ht;
        }
#pragma unroll
        for (int c = 0; c < CP; ++c) a.pbar[c] += x[c];
    }
}

// The reference's formulas verbatim (softmax, clamp at eps, log per value), read straight from global
// memory: used for the rare pixel whose fast-path sum is not finite (-inf logits) and for every pixel
// when eps is large enough for the clamp inside H[p_t] to matter.
__device__ __noinline__ float literal_entropy_sum(const ReduceParams& p, int b, long long px) {
    float EH = 0.f;
    for (int t = 0; t < p.T; ++t) {
        const float* base = p.in + (((long long)t * p.B + b) * p.C) * p.HW + px;
        float m = -INFINITY;
        for (int c = 0; c < p.C; ++c) m = fmaxf(m, base[(long long)c * p.HW]);
        float S = 0.f;
        for (int c = 0; c < p.C; ++c) S += expf(base[(long long)c * p.HW] - m);
        float Ht = 0.f;
        for (int c = 0; c < p.C; ++c) {
            const float pc = fmaxf(__fdiv_rn(expf(base[(long long)c * p.HW] - m), S), p.eps);
            Ht = fmaf(-pc, logf(pc), Ht);
        }
        EH += Ht;
    }
    return EH;
}

__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---- per pixel epilogue: mean, argmax, entropies, confidence, outputs, histograms ------------------
template <int CP, int KIND, bool EXACT>
__device__ __forceinline__ void pixel_epilogue(Acc<CP>& a, bool live, int b, long long px,
                                               const ReduceParams& p, HistSmem& hs) {
    float pmax = 0.f, Hb2 = 0.f, sump = 0.f;
    int arg = 0;
    const long long obase = (long long)b * p.C * p.HW + px;
#pragma unroll
    for (int c = 0; c < CP; ++c) {
        if (EXACT || c < p.C) {
            const float pb = a.pbar[c] * p.invT;
            // torch.argmax: first maximal index, NaN counts as maximal
            if (c == 0 || pb > pmax || (pb != pb && pmax == pmax)) { pmax = pb; arg = c; }
            const float pc = fmaxf(pb, p.eps);
            Hb2 = fmaf(-pc, lg2_approx(pc), Hb2);          // log2 units; |rel err| of lg2.approx <= 2^-22
            sump += fmaxf(pb, 0.f);
        }
    }
    if (p.pbar && live) {                                  // uniform in p.pbar: no store slots issued when not requested
#pragma unroll
        for (int c = 0; c < CP; ++c)
            if (EXACT || c < p.C) p.pbar[obase + (long long)c * p.HW] = a.pbar[c] * p.invT;
    }
    const float Hb = Hb2 * LN2;
    float conf = pmax;
    if (p.conf_mode == SLU_CONF_RENORM) conf = __fdiv_rn(fmaxf(pmax, 0.f), fmaxf(sump, p.eps));
    const long long o = (long long)b * p.HW + px;
    if (live) {
        if (p.pred) p.pred[o] = arg;
        if (p.conf) p.conf[o] = conf;
        if (p.hnorm) p.hnorm[o] = __fdiv_rn(Hb, p.logC);
        if (p.minorm) {
            // T == 1: H[p_bar] and H[p_1] are the same number in the reference, so MI is exactly 0
            float mi = 0.f;
            if (p.need_ht) {
                float EH = (KIND == SLU_IN_LOGITS) ? fmaf(-LN2, a.AS, a.L + logf(a.P)) : a.EH;
                if (KIND == SLU_IN_LOGITS && (p.literal_clamp || !(fabsf(EH) <= 3.0e38f)))
                    EH = literal_entropy_sum(p, b, px);
                mi = fmaxf(__fdiv_rn(Hb - EH * p.invT, p.logC), 0.f);
            }
            p.minorm[o] = mi;
        }
    }
    if (p.labels) {   // warp-uniform
        const long long lab = live ? p.labels[o] : -1;
        if (p.confmat) {
            const bool ok = live && lab >= 0 && lab < p.C;
            warp_hist_add(hs.confmat, ok ? (int)lab * p.C + arg : 0, ok);
        }
        if (p.bins) {
            const float cf = fminf(fmaxf(conf, 0.f), 1.f);                 // ece.py:83 clamp_(0,1)
            const int bin = (conf == conf) ? find_bin(hs.edges, p.n_bins, cf) : -1;   // NaN stays out of every bin
            const bool ok = live && bin >= 0 && !(p.has_ignore && lab == p.ignore);
            warp_bins_add(hs.bin_n, hs.bin_c, hs.bin_s, bin, (long long)arg == lab, cf, ok);
        }
    }
}

template <int CP>
__device__ __forceinline__ void acc_reset(Acc<CP>& a) {
#pragma unroll
    for (int c = 0; c < CP; ++c) a.pbar[c] = 0.f;
    a.AS = 0.f; a.P = 1.f; a.L = 0.f; a.EH = 0.f;
}

__device__ __forceinline__ void hist_init(HistSmem& hs, const ReduceParams& p, int tid, int nthreads) {
    for (int i = tid; i < p.C * p.C; i += nthreads) hs.confmat[i] = 0;
    for (int i = tid; i < SLU_MAX_BINS; i += nthreads) { hs.bin_n[i] = 0; hs.bin_c[i] = 0; hs.bin_s[i] = 0ull; }
    for (int i = tid; i <= p.n_bins; i += nthreads) hs.edges[i] = p.edges[i];
}
__device__ __forceinline__ void hist_flush(HistSmem& hs, const ReduceParams& p, int tid, int nthreads) {
    if (p.confmat)
        for (int i = tid; i < p.C * p.C; i += nthreads)
            if (hs.confmat[i]) atomicAdd(&p.confmat[i], (unsigned long long)hs.confmat[i]);
    if (p.bins)
        for (int i = tid; i < p.n_bins; i += nthreads) {
            if (hs.bin_n[i]) atomicAdd(&p.bins[i], (unsigned long long)hs.bin_n[i]);
            if (hs.bin_c[i]) atomicAdd(&p.bins[p.n_bins + i], (unsigned long long)hs.bin_c[i]);
            if (hs.bin_s[i]) atomicAdd(&p.bins[2 * p.n_bins + i], hs.bin_s[i]);
        }
}

// =====================================================================================================
// Staged kernel: 8 consumer warps + 1 producer warp, STAGES-deep ring of [C x TILE] slabs.
// =====================================================================================================
// Consumer loop, compiled twice: EXACT (C == CP, no per-class predicates anywhere) and padded.
template <int CP, int KIND, bool EXACT>
__device__ __forceinline__ void consume_tiles(const ReduceParams& p, HistSmem& hs, const float* ring, uint32_t bar0) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int slab = (EXACT ? CP : p.C) * TILE;
    unsigned k = 0;
    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int b = (int)(tile / p.tiles_per_scan);
        const long long px0 = (tile % p.tiles_per_scan) * TILE;
        const bool live = px0 + tid < p.HW;
        Acc<CP> acc;
        acc_reset(acc);
        for (int t = 0; t < p.T; ++t, ++k) {
            const int s = k % STAGES;
            const uint32_t ph = (k / STAGES) & 1u;
            mbar_wait(bar0 + 8 * s, ph);
            const float* st = ring + (size_t)s * slab + tid;
            float x[CP];
#pragma unroll
            for (int c = 0; c < CP; ++c)
                x[c] = (EXACT || c < p.C) ? st[c * TILE] : (KIND == SLU_IN_LOGITS ? -1.0e30f : 0.f);
            __syncwarp();
            if (lane == 0) mbar_arrive(bar0 + 8 * (STAGES + s));   // values are in registers: free the slot
            sample_step<CP, KIND, EXACT>(x, acc, p, t);
        }
        pixel_epilogue<CP, KIND, EXACT>(acc, live, b, px0 + tid, p, hs);
    }
}

template <int CP, int KIND>
__global__ void __launch_bounds__(TILE + 32, Occ<CP>::staged) reduce_staged_kernel(const __grid_constant__ ReduceParams p) {
    extern __shared__ __align__(128) float ring[];          // STAGES * C * TILE floats
    __shared__ __align__(8) unsigned long long bars[2 * STAGES];   // full[0..S) | empty[0..S)
    __shared__ HistSmem hs;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const uint32_t bar0 = smem_u32(bars);

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar0 + 8 * s, 1);                      // full: one arrive (+tx bytes) by the producer
            mbar_init(bar0 + 8 * (STAGES + s), NCONS_WARPS); // empty: one arrive per consumer warp
        }
        fence_mbar_init();
    }
    hist_init(hs, p, tid, blockDim.x);
    __syncthreads();

    if (warp == NCONS_WARPS) {
        // ------------------------------- producer warp -------------------------------------------
        const uint64_t pol = policy_evict_first();           // input is read once: do not keep it in L2
        const int slab = p.C * TILE;                         // floats per stage
        unsigned k = 0;
        for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            const int b = (int)(tile / p.tiles_per_scan);
            const long long px0 = (tile % p.tiles_per_scan) * TILE;
            const int npx = (int)min((long long)TILE, p.HW - px0);
            const uint32_t row_bytes = (uint32_t)npx * 4u;
            for (int t = 0; t < p.T; ++t, ++k) {
                const int s = k % STAGES;
                const uint32_t ph = (k / STAGES) & 1u;
                const uint32_t full = bar0 + 8 * s, empty = bar0 + 8 * (STAGES + s);
                if (lane == 0) {
                    mbar_wait(empty, ph ^ 1u);               // slot drained by all consumer warps
                    mbar_arrive_expect_tx(full, row_bytes * (uint32_t)p.C);
                }
                __syncwarp();
                const float* src = p.in + (((long long)t * p.B + b) * p.C) * p.HW + px0;
                const uint32_t dst = smem_u32(ring + (size_t)s * slab);
                for (int c = lane; c < p.C; c += 32)
                    bulk_g2s(dst + (uint32_t)c * TILE * 4u, src + (long long)c * p.HW, row_bytes, full, pol);
            }
        }
    } else {
        // ------------------------------- consumer warps ------------------------------------------
        if (KIND == SLU_IN_LOGITS && p.C == CP) consume_tiles<CP, KIND, true>(p, hs, ring, bar0);
        else consume_tiles<CP, KIND, false>(p, hs, ring, bar0);
    }
    __syncthreads();
    hist_flush(hs, p, tid, blockDim.x);
}

// =====================================================================================================
// Direct kernel: every thread loads its pixel's values straight from global memory, the next
// sample's C values are prefetched into registers while the current one is reduced.
// =====================================================================================================
template <int CP, int KIND>
__global__ void __launch_bounds__(TILE, 2) reduce_direct_kernel(const __grid_constant__ ReduceParams p) {
    __shared__ HistSmem hs;
    const int tid = threadIdx.x;
    hist_init(hs, p, tid, blockDim.x);
    __syncthreads();
    const float pad = (KIND == SLU_IN_LOGITS) ? -1.0e30f : 0.f;
    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int b = (int)(tile / p.tiles_per_scan);
        const long long px = (tile % p.tiles_per_scan) * TILE + tid;
        const bool live = px < p.HW;
        const long long pxs = live ? px : p.HW - 1;          // clamp so dead lanes load valid memory
        const long long tstride = (long long)p.B * p.C * p.HW;
        const float* base = p.in + ((long long)b * p.C) * p.HW + pxs;
        Acc<CP> acc;
        acc_reset(acc);
        float x[CP], xn[CP];
#pragma unroll
        for (int c = 0; c < CP; ++c) xn[c] = (c < p.C) ? ldg_stream(base + (long long)c * p.HW) : pad;
        for (int t = 0; t < p.T; ++t) {
#pragma unroll
            for (int c = 0; c < CP; ++c) x[c] = xn[c];
            if (t + 1 < p.T) {
                const float* nb = base + (long long)(t + 1) * tstride;
#pragma unroll
                for (int c = 0; c < CP; ++c) xn[c] = (c < p.C) ? ldg_stream(nb + (long long)c * p.HW) : pad;
            }
            sample_step<CP, KIND, false>(x, acc, p, t);
        }
        pixel_epilogue<CP, KIND, false>(acc, live, b, px, p, hs);
    }
    __syncthreads();
    hist_flush(hs, p, tid, blockDim.x);
}

// =====================================================================================================
// Single-sample kernel (T == 1: the single-pass branches src/models/trainer.py:1170-1225, ECE/AUROC updates on
// probabilities or concentrations, config 1 of BASELINE.json).  With one sample per pixel there is nothing to
// stream through a ring: 108 B/pixel, and the per-pixel epilogue is as long as the sample step.  One thread per
// pixel, 128-thread CTAs at full occupancy (<= 64 registers), every thread's C loads issued at once (80 B in
// flight per thread, 160 KB per SM), grid-stride over tiles so the shared-memory histograms are flushed once per CTA.
// =====================================================================================================
constexpr int SINGLE_THREADS = 128;
constexpr int SINGLE_FLUSH_TILES = ATOMIC_HIST_MAX_PX / SINGLE_THREADS;   // the split confidence sums cannot overflow

template <int CP, int KIND, bool EXACT>
__global__ void __launch_bounds__(SINGLE_THREADS, 8) reduce_single_kernel(const __grid_constant__ ReduceParams p) {
    __shared__ AtomicHist hs;
    const int tid = threadIdx.x;
    atomic_hist_zero(hs, p.C, tid, SINGLE_THREADS);
    for (int i = tid; i <= p.n_bins; i += SINGLE_THREADS) hs.edges[i] = p.edges[i];
    __syncthreads();
    const float pad = (KIND == SLU_IN_LOGITS) ? -1.0e30f : 0.f;
    const long long tiles_per_scan = (p.HW + SINGLE_THREADS - 1) / SINGLE_THREADS;
    const long long n_tiles = tiles_per_scan * p.B;
    int since_flush = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (p.labels && ++since_flush > SINGLE_FLUSH_TILES) {        // CTA-uniform
            __syncthreads();
            atomic_hist_flush(hs, p.C, p.n_bins, p.confmat, p.bins, tid, SINGLE_THREADS);
            __syncthreads();
            atomic_hist_zero(hs, p.C, tid, SINGLE_THREADS);
            __syncthreads();
            since_flush = 1;
        }
        const int b = (int)(tile / tiles_per_scan);
        const long long px = (tile - (long long)b * tiles_per_scan) * SINGLE_THREADS + tid;
        const bool live = px < p.HW;
        const float* src = p.in + ((long long)b * p.C) * p.HW + (live ? px : p.HW - 1);
        float x[CP];
#pragma unroll
        for (int c = 0; c < CP; ++c) {
            x[c] = (EXACT || c < p.C) ? ldg_stream(src) : pad;
            src += p.HW;
        }
        // ---- distribution p (in x[]), entropy in log2 units
        float Hb2 = 0.f;
        bool literal = true;
        if (KIND == SLU_IN_LOGITS) {
            float m = x[0];
#pragma unroll
            for (int c = 1; c < CP; ++c) m = fmaxf(m, x[c]);
            const float m2 = m * LOG2E;
            float S = 0.f, A = 0.f;
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                const float arg = fmaf(x[c], LOG2E, -m2);
                const float e = ex2_approx(arg);
                x[c] = e;
                S += e;
                A = fmaf(e, arg, A);
            }
            const float inv = __frcp_rn(S);
#pragma unroll
            for (int c = 0; c < CP; ++c) x[c] *= inv;
            if (!p.literal_clamp) {
                // H[p] = ln S - ln2 * A/S: no per-class log.  The eps clamp it skips moves H by <= C*eps*|ln eps| (5.5e-10).
                Hb2 = fmaf(-A, inv, lg2_approx(S));
                literal = !(fabsf(Hb2) <= 3.0e38f);              // -inf logits / overflow
            }
        } else if (KIND == SLU_IN_ALPHA) {                       // p = alpha / (alpha0 + eps)   (src/metrics/ece.py:57-58)
            float a0 = 0.f;
#pragma unroll
            for (int c = 0; c < CP; ++c) a0 += x[c];
            const float d = a0 + p.eps;
#pragma unroll
            for (int c = 0; c < CP; ++c) x[c] = __fdiv_rn(x[c], d);
        }
        if (literal && (p.hnorm != nullptr)) {
            Hb2 = 0.f;
#pragma unroll
            for (int c = 0; c < CP; ++c)
                if (EXACT || c < p.C) { const float pc = fmaxf(x[c], p.eps); Hb2 = fmaf(-pc, lg2_approx(pc), Hb2); }
        }
        // ---- arg max (torch.argmax: first maximal index, NaN counts as maximal), renormalising sum.  The plain sum of
        //      the distribution is NaN exactly when some entry is: only then does the NaN-aware comparison run.
        float pmax = x[0], sump = fmaxf(x[0], 0.f), psum = x[0];
        int arg = 0;
#pragma unroll
        for (int c = 1; c < CP; ++c) {
            if (EXACT || c < p.C) {
                const bool gt = x[c] > pmax;
                pmax = gt ? x[c] : pmax;
                arg = gt ? c : arg;
                sump += fmaxf(x[c], 0.f);
                psum += x[c];
            }
        }
        if (psum != psum) {                                       // rare: NaN / inf inputs
            pmax = x[0]; arg = 0;
#pragma unroll
            for (int c = 1; c < CP; ++c) {
                if (EXACT || c < p.C) {
                    const bool gt = (x[c] > pmax) | ((x[c] != x[c]) & (pmax == pmax));
                    pmax = gt ? x[c] : pmax;
                    arg = gt ? c : arg;
                }
            }
        }
        float conf = pmax;
        if (p.conf_mode == SLU_CONF_RENORM) conf = __fdiv_rn(fmaxf(pmax, 0.f), fmaxf(sump, p.eps));
        const long long o = (long long)b * p.HW + px;
        if (live) {
            if (p.pbar) {                                         // CTA-uniform: skipped entirely when not requested
                float* dst = p.pbar + (long long)b * p.C * p.HW + px;
#pragma unroll
                for (int c = 0; c < CP; ++c) {
                    if (EXACT || c < p.C) *dst = x[c];
                    dst += p.HW;
                }
            }
            if (p.pred) p.pred[o] = arg;
            if (p.conf) p.conf[o] = conf;
            if (p.hnorm) p.hnorm[o] = __fdiv_rn(Hb2 * LN2, p.logC);
            if (p.minorm) p.minorm[o] = 0.f;                      // H[p_bar] and H[p_1] are the same number: MI is exactly 0
        }
        if (p.labels && live) {
            const long long lab = p.labels[o];
            atomic_hist_add(hs, p.C, p.n_bins, p.confmat != nullptr, p.bins != nullptr, p.bins_one_step != 0, lab, arg, arg, conf,
                            p.has_ignore != 0, p.ignore);
        }
    }
    __syncthreads();
    atomic_hist_flush(hs, p.C, p.n_bins, p.confmat, p.bins, tid, SINGLE_THREADS);
}

// ---- single-sample LOGITS with plain (not renormalised) confidence: config 1 of BASELINE.json ------------------------
// (the single-pass branch src/models/trainer.py:1170-1225 and ECEAggregator / IoUEvaluator updates in 'logits' mode).
// Against reduce_single_kernel above, two things change:
//   * less arithmetic per class: the arg-max is taken over the logits while the running maximum is formed (softmax is
//     monotone; first maximal index on ties, as torch.argmax of the probabilities gives), the confidence is
//     e_max / S (no normalised copy of the distribution unless p_bar is asked for), NaN / infinite inputs are detected
//     from S and A and sent to an out-of-line literal routine -- about 10 instructions per class instead of 17;
//   * the reliability bins are PRIVATE to a thread: cells laid out [bin][thread] in shared memory (bank = lane, plain
//     conflict-free read-modify-write, no atomics), packed n << 16 | n_correct plus a u64 fixed-point confidence sum, and
//     the confusion matrix has one copy per warp.  The shared-atomic epilogue of the kernel above costs it a quarter of
//     its bandwidth on 15 hot bins (0.74 -> 0.60 of the HBM peak, profiles/reduce_single_r01_ncu_summary.txt).
// Needs bins that pass one_step_bin_search_ok() and (n_bins + 1) * 128 * 12 B of shared memory; other cases take
// reduce_single_kernel.
constexpr int S2_THREADS = 128;
constexpr int S2_WARPS = S2_THREADS / 32;
constexpr int S2_MAX_BINS = 20;                       // 21 * 128 * 12 B = 32 KB of private cells
constexpr long long S2_MAX_PX_PER_THREAD = 32768;     // packed u16 counts cannot overflow below 65536 pixels

struct S2Odd { int arg; float conf; float hb2; };

// literal evaluation of one pixel (NaN / infinite logits, overflowing sums): the formulas of reduce_single_kernel
template <int CP>
__device__ __noinline__ S2Odd single_pixel_literal(const ReduceParams& p, const float* src) {
    float x[CP];
    for (int c = 0; c < CP; ++c) x[c] = c < p.C ? src[(long long)c * p.HW] : -1.0e30f;
    float m = x[0];
    for (int c = 1; c < CP; ++c) m = fmaxf(m, x[c]);
    const float m2 = m * LOG2E;
    float S = 0.f;
    for (int c = 0; c < CP; ++c) { x[c] = ex2_approx(fmaf(x[c], LOG2E, -m2)); S += x[c]; }
    const float inv = __frcp_rn(S);
    for (int c = 0; c < CP; ++c) x[c] *= inv;
    S2Odd o;
    o.hb2 = 0.f;
    for (int c = 0; c < p.C; ++c) { const float pc = fmaxf(x[c], p.eps); o.hb2 = fmaf(-pc, lg2_approx(pc), o.hb2); }
    float pmax = x[0];
    o.arg = 0;
    for (int c = 1; c < p.C; ++c) {
        const bool gt = (x[c] > pmax) | ((x[c] != x[c]) & (pmax == pmax));      // torch.argmax: first maximum, NaN maximal
        pmax = gt ? x[c] : pmax;
        o.arg = gt ? c : o.arg;
    }
    o.conf = pmax;
    return o;
}

template <int CP, bool EXACT>
__global__ void __launch_bounds__(S2_THREADS, 7) reduce_single_logits_kernel(const __grid_constant__ ReduceParams p) {
    extern __shared__ __align__(16) unsigned char s2_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cells = p.C * p.C;
    const int nb1 = p.n_bins + 1;                                     // + one spare row for pixels that do not count
    unsigned long long* bsum = reinterpret_cast<unsigned long long*>(s2_smem);
    unsigned* bnc = reinterpret_cast<unsigned*>(s2_smem + (p.bins ? nb1 * S2_THREADS * 8 : 0));
    unsigned* cm = reinterpret_cast<unsigned*>(s2_smem + (p.bins ? nb1 * S2_THREADS * 12 : 0));
    float* edges = reinterpret_cast<float*>(cm + (p.confmat ? cells * S2_WARPS : 0));
    if (p.bins)
        for (int i = tid; i < nb1 * S2_THREADS; i += S2_THREADS) { bnc[i] = 0; bsum[i] = 0ull; }
    if (p.confmat)
        for (int i = tid; i < cells * S2_WARPS; i += S2_THREADS) cm[i] = 0;
    for (int i = tid; i <= p.n_bins; i += S2_THREADS) edges[i] = p.edges[i];
    __syncthreads();
    unsigned* my_cm = cm + warp * cells;
    const float nb_f = (float)p.n_bins;
    const long long tiles_per_scan = (p.HW + S2_THREADS - 1) / S2_THREADS;
    const long long n_tiles = tiles_per_scan * p.B;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int b = (int)(tile / tiles_per_scan);
        const long long px = (tile - (long long)b * tiles_per_scan) * S2_THREADS + tid;
        const bool live = px < p.HW;
        const float* src = p.in + ((long long)b * p.C) * p.HW + (live ? px : p.HW - 1);
        const long long o = (long long)b * p.HW + px;
        const bool has_lab = p.labels && live;
        const long long lab = has_lab ? p.labels[o] : 0;
        float x[CP];
#pragma unroll
        for (int c = 0; c < CP; ++c) x[c] = (EXACT || c < p.C) ? ldg_stream(src + (long long)c * p.HW) : -1.0e30f;
        float m = x[0];
        int arg = 0;
#pragma unroll
        for (int c = 1; c < CP; ++c) {
            const bool gt = x[c] > m;
            m = gt ? x[c] : m;
            arg = gt ? c : arg;
        }
        const float m2 = m * LOG2E;
        float S = 0.f, A = 0.f;
#pragma unroll
        for (int c = 0; c < CP; ++c) {
            const float a = fmaf(x[c], LOG2E, -m2);
            const float e = ex2_approx(a);
            x[c] = e;
            S += e;
            A = fmaf(e, a, A);
        }
        const float inv = __frcp_rn(S);
        float conf = ex2_approx(fmaf(m, LOG2E, -m2)) * inv;             // p at the arg-max, the product reduce_single_kernel forms
        float Hb2 = fmaf(-A, inv, lg2_approx(S));                        // H[p] = ln S - ln2 A/S in log2 units
        if (!(fabsf(Hb2) <= 3.0e38f) || !(S <= 3.0e38f)) {               // NaN / infinite logits: literal routine, out of line
            const S2Odd odd = single_pixel_literal<CP>(p, src);
            arg = odd.arg; conf = odd.conf; Hb2 = odd.hb2;
        }
        if (live) {
            if (p.pbar) {                                                // CTA-uniform
                float* dst = p.pbar + (long long)b * p.C * p.HW + px;
#pragma unroll
                for (int c = 0; c < CP; ++c)
                    if (EXACT || c < p.C) dst[(long long)c * p.HW] = x[c] * inv;
            }
            if (p.pred) p.pred[o] = arg;
            if (p.conf) p.conf[o] = conf;
            if (p.hnorm) p.hnorm[o] = __fdiv_rn(Hb2 * LN2, p.logC);
            if (p.minorm) p.minorm[o] = 0.f;                             // H[p_bar] and H[p_1] are the same number
        }
        if (p.labels) {
            if (p.confmat && has_lab && (unsigned long long)lab < (unsigned long long)p.C) atomicAdd(&my_cm[(int)lab * p.C + arg], 1u);   // evaluator.py:49
            if (p.bins) {
                const float c = __saturatef(conf);                        // ece.py:83 clamp_(0,1); NaN -> 0, dropped below
                int k = (int)(c * nb_f);
                k = k > p.n_bins - 1 ? p.n_bins - 1 : k;
                const float lo = edges[k], hi = edges[k + 1];
                k += (c >= hi && k < p.n_bins - 1) ? 1 : 0;             // the host has checked one_step_bin_search_ok()
                k -= (c < lo && k > 0) ? 1 : 0;
                const bool ok = has_lab && conf == conf && c >= edges[0] && c <= edges[p.n_bins] &&
                                !(p.has_ignore && lab == p.ignore);
                const int cell = (ok ? k : p.n_bins) * S2_THREADS + tid;
                bnc[cell] += 0x10000u + ((long long)arg == lab ? 1u : 0u);
                bsum[cell] += __float2ull_rn(c * 4294967296.0f);
            }
        }
    }
    __syncthreads();
    if (p.confmat)
        for (int i = tid; i < cells; i += S2_THREADS) {
            unsigned v = 0;
#pragma unroll
            for (int k = 0; k < S2_WARPS; ++k) v += cm[k * cells + i];
            if (v) atomicAdd(&p.confmat[i], (unsigned long long)v);
        }
    if (p.bins)
        for (int bn = warp; bn < p.n_bins; bn += S2_WARPS) {
            unsigned n = 0, c = 0;
            unsigned long long sfx = 0ull;
#pragma unroll
            for (int k = 0; k < S2_THREADS / 32; ++k) {
                const unsigned v = bnc[bn * S2_THREADS + k * 32 + lane];
                n += v >> 16; c += v & 0xffffu;
                sfx += bsum[bn * S2_THREADS + k * 32 + lane];
            }
            n = __reduce_add_sync(0xffffffffu, n);
            c = __reduce_add_sync(0xffffffffu, c);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) sfx += __shfl_xor_sync(0xffffffffu, sfx, off);
            if (lane == 0 && n) {
                atomicAdd(&p.bins[bn], (unsigned long long)n);
                if (c) atomicAdd(&p.bins[p.n_bins + bn], (unsigned long long)c);
                atomicAdd(&p.bins[2 * p.n_bins + bn], sfx);
            }
        }
}

// ---- the same kernel with FOUR consecutive pixels per thread: 16-byte loads and stores ---------------------------------
// One-pixel threads issue 4-byte requests (a warp touches 128 B per class plane); at 108 B/pixel that tops out near
// 0.74 of the HBM peak even without histograms.  Here a thread owns 4 adjacent pixels: one LDG.128 per class, float4 /
// longlong2 stores, 320 B in flight per thread, 256-thread CTAs that live for many tiles (the private reliability cells
// are zeroed and reduced once per CTA).  Used when HW % 4 == 0, the buffers are 16-byte aligned and the launch has
// enough pixels to fill the machine; smaller or ragged inputs take the kernels above.
constexpr int S4_THREADS = 256;
constexpr int S4_WARPS = S4_THREADS / 32;
constexpr int S4_PX = 4;
#ifndef SLU_S4_MINB
#define SLU_S4_MINB 2
#endif

__device__ __forceinline__ float4 ldg_stream4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

template <int CP, bool EXACT>
__global__ void __launch_bounds__(S4_THREADS, SLU_S4_MINB) reduce_single_logits4_kernel(const __grid_constant__ ReduceParams p) {
    extern __shared__ __align__(16) unsigned char s2_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cells = p.C * p.C;
    const int nb1 = p.n_bins + 1;
    unsigned long long* bsum = reinterpret_cast<unsigned long long*>(s2_smem);
    unsigned* bnc = reinterpret_cast<unsigned*>(s2_smem + (p.bins ? nb1 * S4_THREADS * 8 : 0));
    unsigned* cm = reinterpret_cast<unsigned*>(s2_smem + (p.bins ? nb1 * S4_THREADS * 12 : 0));
    float* edges = reinterpret_cast<float*>(cm + (p.confmat ? cells * S4_WARPS : 0));
    if (p.bins)
        for (int i = tid; i < nb1 * S4_THREADS; i += S4_THREADS) { bnc[i] = 0; bsum[i] = 0ull; }
    if (p.confmat)
        for (int i = tid; i < cells * S4_WARPS; i += S4_THREADS) cm[i] = 0;
    for (int i = tid; i <= p.n_bins; i += S4_THREADS) edges[i] = p.edges[i];
    __syncthreads();
    unsigned* my_cm = cm + warp * cells;
    const float nb_f = (float)p.n_bins;
    constexpr int TILE_PX = S4_THREADS * S4_PX;
    const long long tiles_per_scan = (p.HW + TILE_PX - 1) / TILE_PX;
    const long long n_tiles = tiles_per_scan * p.B;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int b = (int)(tile / tiles_per_scan);
        const long long px = (tile - (long long)b * tiles_per_scan) * TILE_PX + (long long)tid * S4_PX;
        const bool live = px < p.HW;                                   // HW % 4 == 0: a thread's 4 pixels are all in or all out
        const float* src = p.in + ((long long)b * p.C) * p.HW + (live ? px : p.HW - S4_PX);
        const long long o = (long long)b * p.HW + px;
        const bool has_lab = p.labels && live;
        long long lab[S4_PX] = {0, 0, 0, 0};
        if (has_lab) {
            const longlong2 l0 = *reinterpret_cast<const longlong2*>(p.labels + o), l1 = *reinterpret_cast<const longlong2*>(p.labels + o + 2);
            lab[0] = l0.x; lab[1] = l0.y; lab[2] = l1.x; lab[3] = l1.y;
        }
        float x[CP][S4_PX];
#pragma unroll
        for (int c = 0; c < CP; ++c) {
            if (EXACT || c < p.C) {
                const float4 v = ldg_stream4(src + (long long)c * p.HW);
                x[c][0] = v.x; x[c][1] = v.y; x[c][2] = v.z; x[c][3] = v.w;
            } else {
                x[c][0] = x[c][1] = x[c][2] = x[c][3] = -1.0e30f;
            }
        }
        int arg[S4_PX];
        float conf[S4_PX], hn[S4_PX], inv[S4_PX];
#pragma unroll
        for (int k = 0; k < S4_PX; ++k) {
            float m = x[0][k];
            int a = 0;
#pragma unroll
            for (int c = 1; c < CP; ++c) {
                const bool gt = x[c][k] > m;
                m = gt ? x[c][k] : m;
                a = gt ? c : a;
            }
            const float m2 = m * LOG2E;
            float S = 0.f, A = 0.f;
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                const float t = fmaf(x[c][k], LOG2E, -m2);
                const float e = ex2_approx(t);
                x[c][k] = e;
                S += e;
                A = fmaf(e, t, A);
            }
            inv[k] = __frcp_rn(S);
            conf[k] = ex2_approx(fmaf(m, LOG2E, -m2)) * inv[k];
            float Hb2 = fmaf(-A, inv[k], lg2_approx(S));
            if (!(fabsf(Hb2) <= 3.0e38f) || !(S <= 3.0e38f)) {           // NaN / infinite logits: literal routine, out of line
                const S2Odd odd = single_pixel_literal<CP>(p, src + k);
                a = odd.arg; conf[k] = odd.conf; Hb2 = odd.hb2;
            }
            arg[k] = a;
            hn[k] = __fdiv_rn(Hb2 * LN2, p.logC);
        }
        if (live) {
            if (p.pbar) {                                                // CTA-uniform
                float* dst = p.pbar + (long long)b * p.C * p.HW + px;
#pragma unroll
                for (int c = 0; c < CP; ++c)
                    if (EXACT || c < p.C)
                        *reinterpret_cast<float4*>(dst + (long long)c * p.HW) =
                            make_float4(x[c][0] * inv[0], x[c][1] * inv[1], x[c][2] * inv[2], x[c][3] * inv[3]);
            }
            if (p.pred) {
                *reinterpret_cast<longlong2*>(p.pred + o) = make_longlong2(arg[0], arg[1]);
                *reinterpret_cast<longlong2*>(p.pred + o + 2) = make_longlong2(arg[2], arg[3]);
            }
            if (p.conf) *reinterpret_cast<float4*>(p.conf + o) = make_float4(conf[0], conf[1], conf[2], conf[3]);
            if (p.hnorm) *reinterpret_cast<float4*>(p.hnorm + o) = make_float4(hn[0], hn[1], hn[2], hn[3]);
            if (p.minorm) *reinterpret_cast<float4*>(p.minorm + o) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (p.labels) {
#pragma unroll
            for (int k = 0; k < S4_PX; ++k) {
                if (p.confmat && has_lab && (unsigned long long)lab[k] < (unsigned long long)p.C) atomicAdd(&my_cm[(int)lab[k] * p.C + arg[k]], 1u);
                if (p.bins) {
                    const float c = __saturatef(conf[k]);
                    int kk = (int)(c * nb_f);
                    kk = kk > p.n_bins - 1 ? p.n_bins - 1 : kk;
                    const float lo = edges[kk], hi = edges[kk + 1];
                    kk += (c >= hi && kk < p.n_bins - 1) ? 1 : 0;
                    kk -= (c < lo && kk > 0) ? 1 : 0;
                    const bool ok = has_lab && conf[k] == conf[k] && c >= edges[0] && c <= edges[p.n_bins] &&
                                    !(p.has_ignore && lab[k] == p.ignore);
                    const int cell = (ok ? kk : p.n_bins) * S4_THREADS + tid;
                    bnc[cell] += 0x10000u + ((long long)arg[k] == lab[k] ? 1u : 0u);
                    bsum[cell] += __float2ull_rn(c * 4294967296.0f);
                }
            }
        }
    }
    __syncthreads();
    if (p.confmat)
        for (int i = tid; i < cells; i += S4_THREADS) {
            unsigned v = 0;
#pragma unroll
            for (int k = 0; k < S4_WARPS; ++k) v += cm[k * cells + i];
            if (v) atomicAdd(&p.confmat[i], (unsigned long long)v);
        }
    if (p.bins)
        for (int bn = warp; bn < p.n_bins; bn += S4_WARPS) {
            unsigned n = 0, c = 0;
            unsigned long long sfx = 0ull;
#pragma unroll
            for (int k = 0; k < S4_THREADS / 32; ++k) {
                const unsigned v = bnc[bn * S4_THREADS + k * 32 + lane];
                n += v >> 16; c += v & 0xffffu;
                sfx += bsum[bn * S4_THREADS + k * 32 + lane];
            }
            n = __reduce_add_sync(0xffffffffu, n);
            c = __reduce_add_sync(0xffffffffu, c);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) sfx += __shfl_xor_sync(0xffffffffu, sfx, off);
            if (lane == 0 && n) {
                atomicAdd(&p.bins[bn], (unsigned long long)n);
                if (c) atomicAdd(&p.bins[p.n_bins + bn], (unsigned long long)c);
                atomicAdd(&p.bins[2 * p.n_bins + bn], sfx);
            }
        }
}

static int g_single_no_px4 = [] { const char* e = getenv("SLU_SINGLE_NO_PX4"); return (e && e[0] == '1') ? 1 : 0; }();   // A/B: one pixel per thread
static int g_single_no_private = 0;       // A/B switch (slu_debug_reduce_no_private): 1 = reduce_single_kernel for every T == 1 call

template <int CP>
static int launch_single_logits(const ReduceParams& p, cudaStream_t stream, bool& taken) {
    taken = false;
    const long long n_tiles = ((p.HW + S2_THREADS - 1) / S2_THREADS) * p.B;
    const int sms = sm_count_current_device();
    if (sms <= 0) return fail(SLU_E_DEVICE, "no CUDA device");
    if (g_single_no_private || p.conf_mode != SLU_CONF_RAW || p.literal_clamp || p.n_bins > S2_MAX_BINS ||
        (p.bins && !p.bins_one_step))
        return 0;
    // four pixels per thread when the launch is large enough to keep 2 CTAs of 256 threads per SM busy for several tiles
    const uintptr_t al = reinterpret_cast<uintptr_t>(p.in) | reinterpret_cast<uintptr_t>(p.labels) | reinterpret_cast<uintptr_t>(p.pbar) |
                         reinterpret_cast<uintptr_t>(p.pred) | reinterpret_cast<uintptr_t>(p.conf) | reinterpret_cast<uintptr_t>(p.hnorm) |
                         reinterpret_cast<uintptr_t>(p.minorm);
    const long long tiles4 = ((p.HW + S4_THREADS * S4_PX - 1) / (S4_THREADS * S4_PX)) * p.B;
    if (!g_single_no_px4 && (p.HW & 3) == 0 && (al & 15) == 0 && tiles4 >= 8LL * sms) {
        const long long grid4 = (long long)SLU_S4_MINB * sms;
        if ((tiles4 + grid4 - 1) / grid4 * S4_PX <= S2_MAX_PX_PER_THREAD) {
            const int smem4 = (p.bins ? (p.n_bins + 1) * S4_THREADS * 12 : 0) + (p.confmat ? p.C * p.C * S4_WARPS * 4 : 0) +
                              (SLU_MAX_BINS + 1) * 4 + 16;
            static bool attr4[64][2] = {};
            int dev4 = 0;
            SLU_CUDA(cudaGetDevice(&dev4));
            const bool ex4 = p.C == CP;
            if (dev4 < 64 && !attr4[dev4][ex4 ? 1 : 0]) {
                if (ex4) SLU_CUDA(cudaFuncSetAttribute(reduce_single_logits4_kernel<CP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
                else SLU_CUDA(cudaFuncSetAttribute(reduce_single_logits4_kernel<CP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
                attr4[dev4][ex4 ? 1 : 0] = true;
            }
            if (ex4) reduce_single_logits4_kernel<CP, true><<<(unsigned)grid4, S4_THREADS, smem4, stream>>>(p);
            else reduce_single_logits4_kernel<CP, false><<<(unsigned)grid4, S4_THREADS, smem4, stream>>>(p);
            SLU_LAUNCH_CHECK("reduce_single_logits4_kernel");
            taken = true;
            return 0;
        }
    }
    const long long max_ctas = 7LL * sms;
    const long long grid = n_tiles < max_ctas ? n_tiles : max_ctas;
    if ((n_tiles + grid - 1) / grid > S2_MAX_PX_PER_THREAD) return 0;     // packed counts could overflow: the other kernel flushes
    const int smem = (p.bins ? (p.n_bins + 1) * S2_THREADS * 12 : 0) + (p.confmat ? p.C * p.C * S2_WARPS * 4 : 0) +
                     (SLU_MAX_BINS + 1) * 4 + 16;
    static bool attr_set[64][2] = {};
    int dev = 0;
    SLU_CUDA(cudaGetDevice(&dev));
    const bool exact = p.C == CP;
    if (dev < 64 && !attr_set[dev][exact ? 1 : 0]) {
        if (exact) SLU_CUDA(cudaFuncSetAttribute(reduce_single_logits_kernel<CP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        else SLU_CUDA(cudaFuncSetAttribute(reduce_single_logits_kernel<CP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr_set[dev][exact ? 1 : 0] = true;
    }
    if (exact) reduce_single_logits_kernel<CP, true><<<(unsigned)grid, S2_THREADS, smem, stream>>>(p);
    else reduce_single_logits_kernel<CP, false><<<(unsigned)grid, S2_THREADS, smem, stream>>>(p);
    SLU_LAUNCH_CHECK("reduce_single_logits_kernel");
    taken = true;
    return 0;
}

template <int CP, int KIND>
static int launch_single(const ReduceParams& p, cudaStream_t stream) {
    const int sms = sm_count_current_device();
    if (sms <= 0) return fail(SLU_E_DEVICE, "no CUDA device");
    if (KIND == SLU_IN_LOGITS) {
        bool taken = false;
        const int rc = launch_single_logits<CP>(p, stream, taken);
        if (rc || taken) return rc;
    }
    const long long n_tiles = ((p.HW + SINGLE_THREADS - 1) / SINGLE_THREADS) * p.B;
    const long long max_ctas = 8LL * sms;
    const int grid = (int)(n_tiles < max_ctas ? n_tiles : max_ctas);
    if (p.C == CP) reduce_single_kernel<CP, KIND, true><<<grid, SINGLE_THREADS, 0, stream>>>(p);
    else reduce_single_kernel<CP, KIND, false><<<grid, SINGLE_THREADS, 0, stream>>>(p);
    SLU_LAUNCH_CHECK("reduce_single_kernel");
    return 0;
}

// ---- host side --------------------------------------------------------------------------------------
static int g_reduce_no_single = 0;

template <int CP, int KIND>
static int launch(const ReduceParams& p, bool staged, cudaStream_t stream) {
    if (p.T == 1 && p.allow_single && !g_reduce_no_single) return launch_single<CP, KIND>(p, stream);
    const int sms = sm_count_current_device();
    if (sms <= 0) return fail(SLU_E_DEVICE, "no CUDA device");
    const long long max_ctas = (long long)(staged ? Occ<CP>::staged : 2) * sms;
    // balanced persistent grid: every CTA gets the same number of tiles (+-1).  With 512 tiles (one HDL-64 scan) on 444
    // resident CTAs, 68 CTAs would run a second tile alone at a fraction of the HBM bandwidth; 256 CTAs x 2 tiles finish
    // together.  Large batches keep every resident slot busy (more bytes in flight; the last partial wave is < 1/19 there).
    const long long waves = (p.n_tiles + max_ctas - 1) / max_ctas;
    const long long full = p.n_tiles < max_ctas ? p.n_tiles : max_ctas;
    const int grid = (int)(waves <= 4 ? (p.n_tiles + waves - 1) / waves : full);     // many waves: the imbalance is < 1/waves
    if (staged) {
        const size_t dyn = (size_t)STAGES * p.C * TILE * sizeof(float);
        SLU_CUDA(cudaFuncSetAttribute(reduce_staged_kernel<CP, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        reduce_staged_kernel<CP, KIND><<<grid, TILE + 32, dyn, stream>>>(p);
        SLU_LAUNCH_CHECK("reduce_staged_kernel");
    } else {
        reduce_direct_kernel<CP, KIND><<<grid, TILE, 0, stream>>>(p);
        SLU_LAUNCH_CHECK("reduce_direct_kernel");
    }
    return 0;
}

template <int KIND>
static int dispatch_cp(const ReduceParams& p, bool staged, cudaStream_t stream) {
    const int cp = (p.C + 3) / 4 * 4;
#ifdef SLU_FAST_BUILD            // experiments only: one instantiation
    if (cp == 20) return launch<20, KIND>(p, staged, stream);
    return fail(SLU_E_RANGE, "fast build supports C in 17..20 only");
#else
    switch (cp) {
        case 4: return launch<4, KIND>(p, staged, stream);
        case 8: return launch<8, KIND>(p, staged, stream);
        case 12: return launch<12, KIND>(p, staged, stream);
        case 16: return launch<16, KIND>(p, staged, stream);
        case 20: return launch<20, KIND>(p, staged, stream);
        case 24: return launch<24, KIND>(p, staged, stream);
        case 28: return launch<28, KIND>(p, staged, stream);
        case 32: return launch<32, KIND>(p, staged, stream);
    }
    return fail(SLU_E_RANGE, "C=%d outside [2,%d]", p.C, SLU_MAX_CLASSES);
#endif
}

static int reduce_entry(const float* d_in, const int64_t* d_labels, int T, int B, int C, int64_t HW,
                        int in_kind, int conf_mode, float eps, int normalize, int has_ignore, int64_t ignore,
                        int n_bins, const float* h_edges,
                        float* d_pbar, int64_t* d_pred, float* d_conf, float* d_hnorm, float* d_minorm,
                        int64_t* d_confmat, int64_t* d_ece_bins, slu_stream_t stream, bool allow_staged) {
    if (!d_in) return fail(SLU_E_ARG, "d_in is NULL");
    if (T < 1 || B < 1 || HW < 1) return fail(SLU_E_ARG, "T=%d B=%d HW=%lld must be >= 1", T, B, (long long)HW);
    if (C < 2 || C > SLU_MAX_CLASSES) return fail(SLU_E_RANGE, "C=%d outside [2,%d]", C, SLU_MAX_CLASSES);
    if (in_kind < SLU_IN_LOGITS || in_kind > SLU_IN_ALPHA) return fail(SLU_E_ARG, "in_kind=%d", in_kind);
    if (in_kind == SLU_IN_ALPHA && T != 1) return fail(SLU_E_ARG, "SLU_IN_ALPHA needs T=1, got %d", T);
    if (conf_mode != SLU_CONF_RAW && conf_mode != SLU_CONF_RENORM) return fail(SLU_E_ARG, "conf_mode=%d", conf_mode);
    if (!(eps >= 0.f)) return fail(SLU_E_ARG, "eps must be >= 0");
    if ((d_confmat || d_ece_bins) && !d_labels) return fail(SLU_E_ARG, "histograms requested without labels");
    if (d_ece_bins) {
        if (n_bins < 1 || n_bins > SLU_MAX_BINS) return fail(SLU_E_RANGE, "n_bins=%d outside [1,%d]", n_bins, SLU_MAX_BINS);
        if (!h_edges) return fail(SLU_E_ARG, "h_edges is NULL");
        for (int i = 0; i < n_bins; ++i)
            if (!(h_edges[i] < h_edges[i + 1])) return fail(SLU_E_ARG, "bin edges must increase strictly");
    }
    if ((reinterpret_cast<uintptr_t>(d_in) & 3) != 0) return fail(SLU_E_ALIGN, "d_in not 4-byte aligned");

    ReduceParams p{};
    p.in = d_in;
    p.labels = reinterpret_cast<const long long*>(d_labels);
    p.T = T; p.B = B; p.C = C; p.HW = HW;
    p.conf_mode = conf_mode;
    p.eps = eps;
    p.has_ignore = has_ignore; p.ignore = ignore;
    p.n_bins = d_ece_bins ? n_bins : 0;
    for (int i = 0; i <= p.n_bins && d_ece_bins; ++i) p.edges[i] = h_edges[i];
    p.pbar = d_pbar; p.pred = reinterpret_cast<long long*>(d_pred);
    p.conf = d_conf; p.hnorm = d_hnorm; p.minorm = d_minorm;
    p.confmat = reinterpret_cast<unsigned long long*>(d_confmat);
    p.bins = reinterpret_cast<unsigned long long*>(d_ece_bins);
    p.tiles_per_scan = (int)((HW + TILE - 1) / TILE);
    p.n_tiles = (long long)p.tiles_per_scan * B;
    p.logC = normalize ? (float)log((double)C) : 1.0f;    // x / 1.0f is exact: entropies stay in nats
    // the clamp matters once C*eps*|ln eps| reaches fp32 resolution of an O(1) entropy
    p.literal_clamp = (eps > 0.f && (double)C * eps * fabs(log((double)eps)) > 1e-8) ? 1 : 0;
    p.need_ht = (d_minorm != nullptr && T > 1) ? 1 : 0;
    p.invT = 1.0f / (float)T;
    p.allow_single = allow_staged ? 1 : 0;
    p.bins_one_step = (p.n_bins > 0 && one_step_bin_search_ok(p.edges, p.n_bins)) ? 1 : 0;

    const bool staged = allow_staged && (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_in) & 15) == 0);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (in_kind) {
        case SLU_IN_LOGITS: return dispatch_cp<SLU_IN_LOGITS>(p, staged, st);
        case SLU_IN_PROBS: return dispatch_cp<SLU_IN_PROBS>(p, staged, st);
        default: return dispatch_cp<SLU_IN_ALPHA>(p, staged, st);
    }
}

}  // namespace slu

extern "C" int slu_reduce_metrics(const float* d_in, const int64_t* d_labels, int T, int B, int C, int64_t HW,
                                  int in_kind, int conf_mode, float eps, int normalize, int has_ignore, int64_t ignore,
                                  int n_bins, const float* h_edges,
                                  float* d_pbar, int64_t* d_pred, float* d_conf, float* d_hnorm, float* d_minorm,
                                  int64_t* d_confmat, int64_t* d_ece_bins, slu_stream_t stream) {
    return slu::reduce_entry(d_in, d_labels, T, B, C, HW, in_kind, conf_mode, eps, normalize, has_ignore, ignore, n_bins, h_edges,
                             d_pbar, d_pred, d_conf, d_hnorm, d_minorm, d_confmat, d_ece_bins, stream, true);
}

extern "C" int slu_reduce_metrics_direct(const float* d_in, const int64_t* d_labels, int T, int B, int C, int64_t HW,
                                         int in_kind, int conf_mode, float eps, int normalize, int has_ignore, int64_t ignore,
                                         int n_bins, const float* h_edges,
                                         float* d_pbar, int64_t* d_pred, float* d_conf, float* d_hnorm, float* d_minorm,
                                         int64_t* d_confmat, int64_t* d_ece_bins, slu_stream_t stream) {
    return slu::reduce_entry(d_in, d_labels, T, B, C, HW, in_kind, conf_mode, eps, normalize, has_ignore, ignore, n_bins, h_edges,
                             d_pbar, d_pred, d_conf, d_hnorm, d_minorm, d_confmat, d_ece_bins, stream, false);
}

/* A/B switch for tests and profiles: 1 = T == 1 inputs also go through the multi-sample (staged) kernel. */
/* A/B switch (tests, profiles): 1 = T == 1 logits take reduce_single_kernel (shared-atomic histograms) instead of the
 * private-cell kernel. */
extern "C" int slu_debug_reduce_no_private(int on) {
    slu::g_single_no_private = on ? 1 : 0;
    return 0;
}

extern "C" int slu_debug_reduce_no_single(int on) {
    slu::g_reduce_no_single = on ? 1 : 0;
    return 0;
}
