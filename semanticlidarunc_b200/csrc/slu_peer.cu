// slu_peer.cu -- the ONE exchange of the batch-sharded training step (BASELINE.json configs[4]): every rank
// needs the GLOBAL number of valid pixels before its loss kernel can write a final gradient.  As an NCCL all-reduce of one
// float64 that exchange is 30-60 us of latency between the count kernel and the loss kernel at 8 ranks -- as long as the
// rest of a 2-scan step.  Here the count kernel does the exchange itself over NVLink / NVSwitch peer memory:
//
//   peer_allreduce_i64_kernel  the same mailboxes carry the one all-reduce of an evaluation sweep (confusion matrix +
//                           reliability bins, C*C + 3*n_bins int64 <= 512): payload stores, fence, release flag, acquire.
//   count_exchange_kernel   every CTA counts valid pixels of its part of the label map; the LAST CTA to arrive (ticket
//                           counter) packs (step number << 32 | local count) into one 64-bit word, stores that word into
//                           slot [parity][my rank] of EVERY rank's mailbox (plain 8-byte peer stores: atomic on the wire,
//                           so the word is its own ready flag), then spins on its own mailbox until all `world` words of
//                           this step have arrived and writes their sum -- an integer, identical on every rank -- as the
//                           float64 count the loss kernel reads.
//
// Mailboxes are cudaMalloc'ed by the library and shared between the ranks' processes with CUDA IPC handles (the Python
// side passes the 64 handle bytes through torch.distributed once).  The step number lives in device memory and is
// advanced by the kernel, so the launch is capturable in a CUDA graph; two parity halves of the mailbox cover the one
// step by which a fast rank can run ahead (it cannot finish step s+1 before every peer has published step s+1, which a
// peer does only after it has read step s).  A bounded spin turns a missing peer into a NaN count instead of a hang.
// The reference has no counterpart (single process; its dormant all_reduce is src/utils/agg.py:75-83).
#include <stdlib.h>
#include <string.h>
#include "slu_common.cuh"

namespace slu {

constexpr int PEER_MAX_WORLD = 16;
constexpr int PEER_THREADS = 256;
constexpr int PEER_MAX_IGNORE = 8;

constexpr int PEER_VEC_MAX = 512;                   // longest int64 vector of slu_peer_allreduce_i64 (C*C + 3*n_bins = 445 at C=20, 15 bins)

// one rank's mailbox (device memory of that rank, mapped into every peer)
struct PeerMailbox {
    unsigned long long slot[2][PEER_MAX_WORLD];     // [parity][source rank]: step << 32 | count
    unsigned long long step;                        // this rank's step counter (local use only)
    unsigned long long acc;                         // local count accumulator of the running launch
    unsigned int ticket;                            // CTA arrival counter of the running launch
    unsigned int timeouts;                          // number of exchanges that gave up waiting
    // small-vector all-reduce (slu_peer_allreduce_i64): payload first, then the flag that publishes it
    unsigned long long vstep;
    unsigned long long vflag[2][PEER_MAX_WORLD];    // [parity][source rank]: step number of the payload in vec[parity][rank]
    long long vec[2][PEER_MAX_WORLD][PEER_VEC_MAX];
};
constexpr size_t PEER_MAILBOX_BYTES = (sizeof(PeerMailbox) + 4095) / 4096 * 4096;

struct PeerParams {
    const long long* target;
    const unsigned char* keep;
    long long n_px;
    long long ignore[PEER_MAX_IGNORE];
    int n_ignore;
    PeerMailbox* box[PEER_MAX_WORLD];               // box[r] = rank r's mailbox as mapped HERE; box[rank] is the local one
    int rank, world;
    double* count_out;                              // [1] global count
    unsigned long long spin_ns;                     // give up after this long
};

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(PEER_THREADS) count_exchange_kernel(const __grid_constant__ PeerParams p) {
    PeerMailbox* mine = p.box[p.rank];
    unsigned n = 0;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < p.n_px; g += (long long)gridDim.x * blockDim.x) {
        bool v = true;
        if (p.keep) {
            v = p.keep[g] != 0;
        } else {
            const long long t = p.target[g];
#pragma unroll
            for (int i = 0; i < PEER_MAX_IGNORE; ++i)
                if (i < p.n_ignore && t == p.ignore[i]) v = false;
        }
        n += v ? 1u : 0u;
    }
    n = __reduce_add_sync(0xffffffffu, n);
    __shared__ unsigned s_n[PEER_THREADS / 32];
    __shared__ bool s_last;
    if ((threadIdx.x & 31) == 0) s_n[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int i = 0; i < PEER_THREADS / 32; ++i) t += s_n[i];
        if (t) atomicAdd(&mine->acc, (unsigned long long)t);
        __threadfence();
        s_last = atomicAdd(&mine->ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last || threadIdx.x >= 32) return;
    // ---- last CTA, warp 0: publish, collect, clean up
    __threadfence();
    const int lane = threadIdx.x;
    unsigned long long local = 0, step = 0;
    if (lane == 0) {
        local = atomicAdd(&mine->acc, 0ull);
        step = mine->step + 1;
        mine->step = step;
        mine->acc = 0ull;
        mine->ticket = 0u;
    }
    local = __shfl_sync(0xffffffffu, local, 0);
    step = __shfl_sync(0xffffffffu, step, 0);
    const unsigned tag = (unsigned)(step & 0xffffffffull);
    const int par = (int)(step & 1ull);
    const unsigned long long word = ((unsigned long long)tag << 32) | (local & 0xffffffffull);
    if (lane < p.world) st_sys_u64(&p.box[lane]->slot[par][p.rank], word);      // one 8-byte store per peer, own mailbox included
    unsigned long long got = 0;
    bool ok = true;
    if (lane < p.world) {
        const unsigned long long t0 = globaltimer_ns();
        for (;;) {
            got = ld_volatile_u64(&mine->slot[par][lane]);
            if ((unsigned)(got >> 32) == tag) break;
            if (globaltimer_ns() - t0 > p.spin_ns) { ok = false; break; }
            __nanosleep(64);
        }
    }
    const bool all_ok = __all_sync(0xffffffffu, ok);
    unsigned long long sum = lane < p.world ? (got & 0xffffffffull) : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);       // integers: the same value on every rank
    if (lane == 0) {
        *p.count_out = all_ok ? (double)sum : __longlong_as_double(0x7ff8000000000000ll);
        if (!all_ok) atomicAdd(&mine->timeouts, 1u);
    }
}

struct PeerVecParams {
    const long long* a;                             // first part of the vector (e.g. the confusion matrix)
    const long long* b;                             // second part (e.g. the reliability bins) or NULL
    int n_a, n_b;
    long long* out;                                 // [n_a + n_b] sums over the ranks
    PeerMailbox* box[PEER_MAX_WORLD];
    int rank, world;
    unsigned long long spin_ns;
};

__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// One CTA: every rank stores its vector into slot [rank] of every rank's mailbox (8-byte peer stores, one element per
// thread and peer), fences, publishes the step number with a release store, waits (acquire) for the `world` flags of this
// step in its own mailbox and sums the payloads in rank order -- integers, so every rank gets the same bits.
__global__ void __launch_bounds__(PEER_VEC_MAX) peer_allreduce_i64_kernel(const __grid_constant__ PeerVecParams p) {
    PeerMailbox* mine = p.box[p.rank];
    __shared__ unsigned long long s_step;
    __shared__ int s_ok;
    if (threadIdx.x == 0) { s_step = mine->vstep + 1; mine->vstep = s_step; s_ok = 1; }
    __syncthreads();
    const unsigned long long step = s_step;
    const int par = (int)(step & 1ull), n = p.n_a + p.n_b, i = threadIdx.x;
    if (i < n) {
        const long long v = i < p.n_a ? p.a[i] : p.b[i - p.n_a];
        for (int r = 0; r < p.world; ++r) st_sys_u64(reinterpret_cast<unsigned long long*>(&p.box[r]->vec[par][p.rank][i]), (unsigned long long)v);
    }
    __threadfence_system();
    __syncthreads();
    if (i < p.world) {
        st_release_sys_u64(&p.box[i]->vflag[par][p.rank], step);
        const unsigned long long t0 = globaltimer_ns();
        while (ld_acquire_sys_u64(&mine->vflag[par][i]) != step) {
            if (globaltimer_ns() - t0 > p.spin_ns) { atomicExch(&s_ok, 0); break; }
            __nanosleep(64);
        }
        __threadfence_system();
    }
    __syncthreads();
    if (i < n) {
        long long sum = 0;
        for (int r = 0; r < p.world; ++r) sum += (long long)ld_volatile_u64(reinterpret_cast<const unsigned long long*>(&mine->vec[par][r][i]));
        p.out[i] = s_ok ? sum : (long long)0x8000000000000000ull;        // a missing peer poisons the result instead of hanging
    }
    if (i == 0 && !s_ok) atomicAdd(&mine->timeouts, 1u);
}

}  // namespace slu

extern "C" int slu_peer_mailbox_create(void** d_box_out, uint8_t* handle64_out) {
    using namespace slu;
    if (!d_box_out || !handle64_out) return fail(SLU_E_ARG, "slu_peer_mailbox_create: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    SLU_CUDA(cudaMalloc(&p, PEER_MAILBOX_BYTES));
    cudaError_t e = cudaMemset(p, 0, PEER_MAILBOX_BYTES);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "slu_peer_mailbox_create"); }
    memcpy(handle64_out, &h, 64);
    *d_box_out = p;
    return 0;
}

extern "C" int slu_peer_mailbox_open(const uint8_t* handle64, void** d_box_out) {
    using namespace slu;
    if (!handle64 || !d_box_out) return fail(SLU_E_ARG, "slu_peer_mailbox_open: NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    SLU_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *d_box_out = p;
    return 0;
}

extern "C" int slu_peer_mailbox_close(void* d_box) {
    using namespace slu;
    if (!d_box) return 0;
    SLU_CUDA(cudaIpcCloseMemHandle(d_box));
    return 0;
}

extern "C" int slu_peer_mailbox_destroy(void* d_box) {
    using namespace slu;
    if (!d_box) return 0;
    SLU_CUDA(cudaFree(d_box));
    return 0;
}

extern "C" int slu_count_valid_exchange(const int64_t* d_target, const uint8_t* d_keep_mask, int64_t n_px,
                                        const int64_t* h_ignore, int n_ignore,
                                        void* const* h_boxes, int rank, int world, double timeout_s,
                                        double* d_count, slu_stream_t stream) {
    using namespace slu;
    if (!d_target || !d_count || !h_boxes) return fail(SLU_E_ARG, "slu_count_valid_exchange: NULL argument");
    if (n_px < 1 || n_px > 0xffffffffLL) return fail(SLU_E_RANGE, "n_px=%lld outside [1, 2^32)", (long long)n_px);
    if (world < 1 || world > PEER_MAX_WORLD || rank < 0 || rank >= world) return fail(SLU_E_RANGE, "rank %d / world %d unsupported (world <= %d)", rank, world, PEER_MAX_WORLD);
    if (n_ignore < 0 || n_ignore > PEER_MAX_IGNORE || (n_ignore > 0 && !h_ignore)) return fail(SLU_E_RANGE, "n_ignore=%d outside [0,%d]", n_ignore, PEER_MAX_IGNORE);
    PeerParams p{};
    p.target = reinterpret_cast<const long long*>(d_target); p.keep = d_keep_mask; p.n_px = n_px;
    for (int i = 0; i < n_ignore; ++i) p.ignore[i] = h_ignore[i];
    p.n_ignore = n_ignore;
    for (int r = 0; r < world; ++r) {
        if (!h_boxes[r]) return fail(SLU_E_ARG, "mailbox of rank %d is NULL", r);
        p.box[r] = static_cast<PeerMailbox*>(h_boxes[r]);
    }
    p.rank = rank; p.world = world; p.count_out = d_count;
    p.spin_ns = (unsigned long long)((timeout_s > 0.0 ? timeout_s : 2.0) * 1e9);
    const int sms = sm_count_current_device();
    if (sms <= 0) return fail(SLU_E_DEVICE, "no CUDA device");
    const long long chunks = (n_px + PEER_THREADS - 1) / PEER_THREADS;
    const long long cap = 2LL * sms;
    count_exchange_kernel<<<(unsigned)(chunks < cap ? chunks : cap), PEER_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    SLU_LAUNCH_CHECK("count_exchange_kernel");
    return 0;
}

extern "C" int slu_peer_mailbox_timeouts(const void* d_box, uint32_t* h_out) {
    using namespace slu;
    if (!d_box || !h_out) return fail(SLU_E_ARG, "slu_peer_mailbox_timeouts: NULL argument");
    SLU_CUDA(cudaMemcpy(h_out, &static_cast<const PeerMailbox*>(d_box)->timeouts, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int slu_peer_allreduce_i64(const int64_t* d_a, int n_a, const int64_t* d_b, int n_b,
                                      void* const* h_boxes, int rank, int world, double timeout_s,
                                      int64_t* d_out, slu_stream_t stream) {
    using namespace slu;
    if (!d_a || !d_out || !h_boxes) return fail(SLU_E_ARG, "slu_peer_allreduce_i64: NULL argument");
    if (n_a < 1 || n_b < 0 || (n_b > 0 && !d_b) || n_a + n_b > PEER_VEC_MAX) return fail(SLU_E_RANGE, "vector of %d + %d elements unsupported (<= %d)", n_a, n_b, PEER_VEC_MAX);
    if (world < 1 || world > PEER_MAX_WORLD || rank < 0 || rank >= world) return fail(SLU_E_RANGE, "rank %d / world %d unsupported (world <= %d)", rank, world, PEER_MAX_WORLD);
    PeerVecParams p{};
    p.a = reinterpret_cast<const long long*>(d_a); p.b = reinterpret_cast<const long long*>(d_b);
    p.n_a = n_a; p.n_b = n_b; p.out = reinterpret_cast<long long*>(d_out);
    for (int r = 0; r < world; ++r) {
        if (!h_boxes[r]) return fail(SLU_E_ARG, "mailbox of rank %d is NULL", r);
        p.box[r] = static_cast<PeerMailbox*>(h_boxes[r]);
    }
    p.rank = rank; p.world = world;
    p.spin_ns = (unsigned long long)((timeout_s > 0.0 ? timeout_s : 2.0) * 1e9);
    peer_allreduce_i64_kernel<<<1, PEER_VEC_MAX, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    SLU_LAUNCH_CHECK("peer_allreduce_i64_kernel");
    return 0;
}
