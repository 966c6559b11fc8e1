// slu_project.cu -- stage 1 (spherical range-image projection with a depth test) and stage 2
// (label back-projection).
//
// Replaces (reference file:line): spherical_projection src/dataset/utils.py:288-349,
// to_deflection_coordinates :61-67, and the KITTI loader glue around them
// (src/dataset/dataloader_semantic_KITTI.py:40-49 label remap + float64 concat, :83 fp32 range).
// The reference sorts the cloud far->near (argsort + gather), digitizes against two linspace edge
// arrays and lets a fancy-index scatter keep the last write; here every point is handled in place:
//
// (exact kernels, used by the generic fp64 entry point; the batched entry point runs the
//  fp32-prefiltered variants further down, which give bit-identical results)
//   K0 init      key[b,px] = ~0, winner[b,px] = INT_MAX, theta min/max cells, diagnostics
//   K1 angles    per point (fp64, same operation order as numpy, no FMA contraction):
//                r, phi, theta -> column index from the fixed azimuth edges, range key, and a
//                block-reduced theta min/max per scan (ordered-integer atomicMin/Max)
//   K2 rows      per point: row index from the scan's theta edges, pix = row*W+col,
//                atomicMin(key[b,pix], range bits)                     -- 64-bit, exact fp64 order
//   K3 ties      per point: if my range bits == key[b,pix]: atomicMin(winner[b,pix], my index)
//                -- deterministic "lowest index" among exactly equal ranges
//   K4 resolve   per pixel: gather the winner, write the channel planes (0 where empty)
//
// The depth test compares the FULL float64 range (a positive double's bit pattern is monotone),
// so it reproduces the reference's sort order exactly; packing range and index into one 64-bit
// word would have to truncate the range to ~44 bits and would create ties the reference does not
// have.  A scan (<= 4 MB) and its image stay in the 126 MB L2, so only K1's read of the points and
// K4's write of the image are HBM traffic: 20 B/point + 24 B/pixel.
//
// Bin edges are numpy's linspace restated bit for bit: step = (stop-start)/(num-1),
// edge(i) = fl(fl(i*step) + start), edge(num-1) = stop.  cnt = #{edges <= v} is found by a guess
// from the step and an exact fix-up against those edges, then row = (H-1-cnt) mod H (the
// reference's digitize(...)-1 with negative indices wrapping).
#include <math.h>
#include <stdlib.h>
#include "slu_common.cuh"

namespace slu {

constexpr int PT_THREADS = 256;
#ifndef SLU_PT_BATCH
#define SLU_PT_BATCH 4
#endif
constexpr int PT_BATCH = SLU_PT_BATCH;     // points (pixels) a thread handles per loop iteration in the gather / scatter kernels
constexpr int MAX_SCANS = 256;
constexpr double HALF_PI = 1.5707963267948966;   // np.pi / 2
constexpr double PI = 3.141592653589793;         // np.pi

struct Edges {            // numpy.linspace(start, stop, num) without materialising it
    double start, stop, step, delta;
    int num;
};

__host__ __device__ inline Edges make_edges(double start, double stop, int num) {
    Edges e;
    e.start = start; e.stop = stop; e.num = num;
    e.delta = stop - start;
    e.step = num > 1 ? e.delta / (double)(num - 1) : 0.0;
    return e;
}

#ifdef __CUDACC__
__device__ __forceinline__ double edge_at(const Edges& e, int i) {
    if (i == e.num - 1 && e.num > 1) return e.stop;
    if (e.num == 1) return e.start;
    if (e.step != 0.0) return __dadd_rn(__dmul_rn((double)i, e.step), e.start);
    // numpy's denormal/zero-step branch: y = (i / div) * delta + start
    return __dadd_rn(__dmul_rn(__ddiv_rn((double)i, (double)(e.num - 1)), e.delta), e.start);
}

// #{i : edge(i) <= v}; NaN counts as beyond every edge (numpy sorts NaN last).  `near` reports
// whether v lies within 4 ulp of an edge, where a 1-2 ulp difference between CUDA's and numpy's
// atan2 could move the point to the neighbouring bin.
__device__ __forceinline__ int count_le(const Edges& e, double v, bool& near) {
    near = false;
    if (v != v) return e.num;
    int i;                                          // candidate: last edge <= v
    if (e.step > 0.0) {
        const double g = floor((v - e.start) / e.step);
        i = g < -1.0 ? -1 : (g > (double)(e.num - 1) ? e.num - 1 : (int)g);
    } else {
        i = -1;
    }
    while (i >= 0 && edge_at(e, i) > v) --i;
    while (i + 1 < e.num && edge_at(e, i + 1) <= v) ++i;
    const double tol = 8.9e-16 * fmax(fabs(v), 2.3e-308);
    if (i >= 0 && fabs(v - edge_at(e, i)) <= tol) near = true;
    if (i + 1 < e.num && fabs(edge_at(e, i + 1) - v) <= tol) near = true;
    return i + 1;
}

__device__ __forceinline__ unsigned long long order_bits(double d) {     // monotone double -> u64
    const unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double unorder_bits(unsigned long long u) {
    const unsigned long long b = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
    return __longlong_as_double((long long)b);
}
#endif

struct ProjParams;
struct Pt { double x, y, z; };
struct ProjParams {
    // input, one of the two forms
    const float4* xyzi;            // [n_total] x,y,z,intensity
    const unsigned* raw_label;     // [n_total] or NULL
    const int* lut;                // [65536] or NULL
    const double* pc;              // generic form: [N,Cin] float64
    const double* yaw;             // [B,2] (cos, sin) of a per-scan yaw applied before projection, or NULL
    int cin;
    long long offsets[MAX_SCANS + 1];
    int B, H, W;
    long long HW;
    int use_range;
    double theta_lo, theta_hi;
    int farthest;
    int key_sq;                    // depth keys hold the bits of r^2 (fast path: no DSQRT per point) instead of r
    const double* row_edges;       // caller-supplied row edges (bins_h of the reference), ASCENDING, [H], or NULL
    int row_edges_increasing;      // the caller's array was increasing: idx = cnt - 1 instead of H - 1 - cnt
    // workspace
    double* theta;                 // [n_total]
    unsigned long long* rkey;      // [n_total]
    int* col;                      // [n_total]
    unsigned long long* key;       // [B*HW]
    unsigned long long* tminmax;   // [B*2] ordered bits: min | max
    // fast (fp32-prefiltered) path
    float* theta32;                // [n_total]
    float* part_min;               // [B*gx] per-block fp32 theta min
    float* part_max;               // [B*gx]
    int* part_diag;                // [B*gx*2] per-block (missing ids, near-edge points)
    double* part_exact;            // [B*gx*2] per-block exact fp64 theta (min, max) over the block's candidates
    int gx;                        // blocks per scan of the point kernels
    int dbg_atom;                  // timing experiments only (SLU_P3_ATOM): 1 = 64-bit RED.MIN on the key, 2 = no atomic
    unsigned long long* dbg_times; // timing experiments only (SLU_P3_TIMES=1): per-block globaltimer stamps, [3][4096][8], or NULL
    struct Cell* cell;             // [B*HW] 128-bit (range key, point index) cells of the single-pass depth test
    int want_pix, want_winner;     // the caller asked for d_pix / d_winner (the cell pipeline skips the stores otherwise)
    int use_cells;                 // the depth test runs on the 128-bit cells (no key / winner arrays, no tie pass)
    int use_cross;                 // near-edge points are decided by the cross-product test where it can tell
    // outputs
    int* pix;                      // [n_total]
    int* winner;                   // [B*HW]
    float* img;                    // [B,6,HW] planes, or [H,W,Cin] generic
    long long* label_img;          // [B,HW] int64 train ids (planes form only)
    double* theta_out;             // [B,2]
    int* diag;                     // [B,2]
};

// A point of the batched form as the loaders see it: float32 file values widened to float64, then the
// optional yaw augmentation rotate_z (src/dataset/utils.py:4-18: xyz @ [[c,-s,0],[s,c,0],[0,0,1]]).
__device__ __forceinline__ Pt load_pt(const ProjParams& p, int b, const float4 v) {
    Pt q;
    q.x = (double)v.x; q.y = (double)v.y; q.z = (double)v.z;
    if (p.yaw) {
        const double c = p.yaw[2 * b], s = p.yaw[2 * b + 1];
        const double x = __dadd_rn(__dmul_rn(q.x, c), __dmul_rn(q.y, s));
        const double y = __dadd_rn(__dmul_rn(q.x, -s), __dmul_rn(q.y, c));
        q.x = x; q.y = y;
    }
    return q;
}

// ---- the per-pixel depth-test cell: (range key, point index) as ONE 128-bit word, updated by compare-and-swap ----
// A point is a winner candidate of its pixel when its range equals the pixel's best range.  With r^2 keys (key_sq) two
// DIFFERENT keys can stand for the same float64 range (sqrt rounds at most a few neighbouring r^2 values together); the
// reference, which sorts by r, sees a tie there, so keys within 8 ulp of the best are compared after the square root.
__device__ __forceinline__ bool same_range(const ProjParams& p, unsigned long long rk, unsigned long long best) {
    if (rk == best) return true;
    if (!p.key_sq || rk - best > 8ull) return false;           // best is the minimum: rk >= best
    const unsigned long long a = p.farthest ? 0x7fffffffffffffffull - rk : rk;
    const unsigned long long c = p.farthest ? 0x7fffffffffffffffull - best : best;
    return __dsqrt_rn(__longlong_as_double((long long)a)) == __dsqrt_rn(__longlong_as_double((long long)c));
}

struct alignas(16) Cell { unsigned long long idx, key; };     // one little-endian 128-bit word: (key << 64) | idx
constexpr unsigned long long CELL_EMPTY = ~0ull;

__device__ __forceinline__ void cas128(Cell* addr, unsigned long long exp_lo, unsigned long long exp_hi,
                                       unsigned long long new_lo, unsigned long long new_hi,
                                       unsigned long long& old_lo, unsigned long long& old_hi) {
    asm volatile(
        "{\n\t.reg .b128 e, n, o;\n\t"
        "mov.b128 e, {%3, %4};\n\t"
        "mov.b128 n, {%5, %6};\n\t"
        "atom.global.relaxed.gpu.cas.b128 o, [%2], e, n;\n\t"
        "mov.b128 {%0, %1}, o;\n\t}"
        : "=l"(old_lo), "=l"(old_hi)
        : "l"(addr), "l"(exp_lo), "l"(exp_hi), "l"(new_lo), "l"(new_hi)
        : "memory");
}

typedef unsigned __int128 u128;
__device__ __forceinline__ u128 cas128q(Cell* addr, u128 expect, u128 desired) {      // result stays ONE 128-bit register
    u128 old;
    asm volatile("atom.global.relaxed.gpu.cas.b128 %0, [%1], %2, %3;" : "=q"(old) : "l"(addr), "q"(expect), "q"(desired) : "memory");
    return old;
}

// strict total order of the depth test: smaller float64 range first (same_range: r^2 keys whose square roots coincide are
// the ties they are in the reference), then the lower point index
__device__ __forceinline__ bool cell_before(const ProjParams& p, unsigned long long ka, unsigned long long ia,
                                            unsigned long long kb, unsigned long long ib) {
    if (ka == kb) return ia < ib;
    const bool a_small = ka < kb;
    if (same_range(p, a_small ? kb : ka, a_small ? ka : kb)) return ia < ib;
    return a_small;
}

// retire or retry after a failed first CAS (old = what the cell held)
__device__ __forceinline__ void depth_test_finish(const ProjParams& p, Cell* a, unsigned long long key, unsigned long long idx,
                                                  unsigned long long ol, unsigned long long oh) {
    unsigned long long el = CELL_EMPTY, eh = CELL_EMPTY;
    while (!(ol == el && oh == eh)) {
        if (!cell_before(p, key, idx, oh, ol)) return;          // the resident point stays
        el = ol; eh = oh;
        cas128(a, el, eh, idx, key, ol, oh);
    }
}
__device__ __forceinline__ void depth_test(const ProjParams& p, Cell* a, unsigned long long key, unsigned long long idx) {
    unsigned long long ol, oh;
    cas128(a, CELL_EMPTY, CELL_EMPTY, idx, key, ol, oh);
    depth_test_finish(p, a, key, idx, ol, oh);
}

// the rare half of the depth test, a real call: the compare-and-swap retry loop with its float64 square roots would
// otherwise be inlined at every site and set the register budget of the point loop
__device__ __noinline__ void depth_test_retry(Cell* a, unsigned long long key, unsigned long long idx,
                                              unsigned long long ol, unsigned long long oh, int key_sq, int farthest) {
    ProjParams q;                                   // only the two fields same_range reads
    q.key_sq = key_sq; q.farthest = farthest;
    depth_test_finish(q, a, key, idx, ol, oh);
}
__device__ __forceinline__ void depth_test_settle(const ProjParams& p, Cell* a, unsigned long long key, unsigned long long idx,
                                                  unsigned long long ol, unsigned long long oh) {
    if ((ol & oh) != CELL_EMPTY) depth_test_retry(a, key, idx, ol, oh, p.key_sq, p.farthest);
}

// issue / settle pair used by the point kernels: several first attempts of a thread are put in flight before any is examined
__device__ __forceinline__ u128 depth_test_issue(Cell* a, unsigned long long key, unsigned long long idx) {
    return cas128q(a, ~(u128)0, ((u128)key << 64) | idx);
}
__device__ __forceinline__ void depth_test_settle(const ProjParams& p, Cell* a, unsigned long long key, unsigned long long idx, u128 old) {
    if (old != ~(u128)0) depth_test_retry(a, key, idx, (unsigned long long)old, (unsigned long long)(old >> 64), p.key_sq, p.farthest);
}

__global__ void __launch_bounds__(PT_THREADS) proj_init_kernel(const __grid_constant__ ProjParams p) {
    const long long total = (long long)p.B * p.HW;
    if (p.use_cells) {
        ulonglong2* cell = reinterpret_cast<ulonglong2*>(p.cell);
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
            cell[i] = make_ulonglong2(CELL_EMPTY, CELL_EMPTY);
    } else {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
            p.key[i] = ~0ull;
            p.winner[i] = 0x7fffffff;
        }
    }
    if (blockIdx.x == 0) {
        for (int b = threadIdx.x; b < p.B; b += blockDim.x) {
            p.tminmax[2 * b] = ~0ull;
            p.tminmax[2 * b + 1] = 0ull;
            p.diag[2 * b] = 0;
            p.diag[2 * b + 1] = 0;
        }
    }
}

template <bool GENERIC>
__global__ void __launch_bounds__(PT_THREADS) proj_angles_kernel(const __grid_constant__ ProjParams p) {
    const int b = blockIdx.y;
    const long long n0 = p.offsets[b], n1 = p.offsets[b + 1];
    const Edges ew = make_edges(-PI, PI, p.W);
    double tmin = INFINITY, tmax = -INFINITY;
    int missing = 0, near_cnt = 0;
    for (long long n = n0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; n < n1; n += (long long)gridDim.x * blockDim.x) {
        double x, y, z;
        if (GENERIC) {
            const double* q = p.pc + n * p.cin;
            x = q[0]; y = q[1]; z = q[2];
        } else {
            const Pt q = load_pt(p, b, __ldg(p.xyzi + n));
            x = q.x; y = q.y; z = q.z;
            if (p.raw_label && p.lut && __ldg(p.lut + (__ldg(p.raw_label + n) & 0xffffu)) < 0) ++missing;
        }
        // r = sqrt(x**2 + y**2 + z**2), p = sqrt(x**2 + y**2)   (utils.py:299, :63), each op rounded
        const double xx = __dmul_rn(x, x), yy = __dmul_rn(y, y), zz = __dmul_rn(z, z);
        const double sxy = __dadd_rn(xx, yy);
        const double r = __dsqrt_rn(__dadd_rn(sxy, zz));
        const double rho = __dsqrt_rn(sxy);
        const double phi = atan2(y, x);                                     // :64
        const double theta = __dadd_rn(-atan2(rho, z), HALF_PI);            // :66
        bool near;
        const int cnt_w = count_le(ew, phi, near);
        if (near) ++near_cnt;
        int c = p.W - 1 - cnt_w;                    // cnt_w in [0, W]: only -1 wraps (numpy's negative index)
        if (c < 0) c += p.W;
        p.col[n] = c;
        p.theta[n] = theta;
        const unsigned long long rb = (unsigned long long)__double_as_longlong(r) & 0x7fffffffffffffffull;
        p.rkey[n] = p.farthest ? (0x7fffffffffffffffull - rb) : rb;
        if (theta == theta) { tmin = fmin(tmin, theta); tmax = fmax(tmax, theta); }
    }
    // block reduction of theta min/max and the diagnostics
    __shared__ double s_min[PT_THREADS / 32], s_max[PT_THREADS / 32];
    __shared__ int s_miss[PT_THREADS / 32], s_near[PT_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        tmin = fmin(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
        tmax = fmax(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
        missing += __shfl_xor_sync(0xffffffffu, missing, o);
        near_cnt += __shfl_xor_sync(0xffffffffu, near_cnt, o);
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s_min[w] = tmin; s_max[w] = tmax; s_miss[w] = missing; s_near[w] = near_cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < PT_THREADS / 32; ++i) {
            tmin = fmin(tmin, s_min[i]); tmax = fmax(tmax, s_max[i]);
            missing += s_miss[i]; near_cnt += s_near[i];
        }
        if (tmin <= tmax) {
            atomicMin(&p.tminmax[2 * b], order_bits(tmin));
            atomicMax(&p.tminmax[2 * b + 1], order_bits(tmax));
        }
        if (missing) atomicAdd(&p.diag[2 * b], missing);
        if (near_cnt) atomicAdd(&p.diag[2 * b + 1], near_cnt);
    }
}

__device__ __forceinline__ void scan_theta_range(const ProjParams& p, int b, double& lo, double& hi) {
    if (p.use_range) { lo = p.theta_lo; hi = p.theta_hi; return; }
    const unsigned long long ulo = p.tminmax[2 * b], uhi = p.tminmax[2 * b + 1];
    if (ulo == ~0ull) { lo = 0.0; hi = 0.0; return; }          // empty scan
    lo = unorder_bits(ulo);
    hi = unorder_bits(uhi);
}

__global__ void __launch_bounds__(PT_THREADS) proj_rows_kernel(const __grid_constant__ ProjParams p) {
    const int b = blockIdx.y;
    const long long n0 = p.offsets[b], n1 = p.offsets[b + 1];
    double lo, hi;
    scan_theta_range(p, b, lo, hi);
    const Edges eh = make_edges(lo, hi, p.H);
    if (blockIdx.x == 0 && threadIdx.x == 0 && p.theta_out) { p.theta_out[2 * b] = lo; p.theta_out[2 * b + 1] = hi; }
    int near_cnt = 0;
    for (long long n = n0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; n < n1; n += (long long)gridDim.x * blockDim.x) {
        bool near;
        int cnt_h, r;
        if (p.row_edges) {
            // np.digitize(theta, bins_h) - 1 with the caller's edges (utils.py:330-338): cnt = #{edges <= theta} by bisection
            const double th = p.theta[n];
            int lo_i = 0, hi_i = p.H;                       // first index with edge > th (NaN: beyond every edge)
            if (th != th) lo_i = p.H;
            while (lo_i < hi_i) {
                const int mid = (lo_i + hi_i) >> 1;
                if (p.row_edges[mid] <= th) lo_i = mid + 1; else hi_i = mid;
            }
            cnt_h = lo_i;
            const double tol = 8.9e-16 * fmax(fabs(th), 2.3e-308);
            near = (cnt_h > 0 && fabs(th - p.row_edges[cnt_h - 1]) <= tol) || (cnt_h < p.H && fabs(p.row_edges[cnt_h] - th) <= tol);
            r = p.row_edges_increasing ? cnt_h - 1 : p.H - 1 - cnt_h;
        } else {
            cnt_h = count_le(eh, p.theta[n], near);
            // the scan's own extreme points sit exactly ON the first/last edge by construction
            if (near && !p.use_range) near = !(p.theta[n] == lo || p.theta[n] == hi);
            r = p.H - 1 - cnt_h;                    // cnt_h in [0, H]: only -1 wraps
        }
        if (near) ++near_cnt;
        if (r < 0) r += p.H;
        const int px = r * p.W + p.col[n];
        p.pix[n] = px;
        atomicMin(&p.key[(long long)b * p.HW + px], p.rkey[n]);
    }
    near_cnt = __reduce_add_sync(0xffffffffu, near_cnt);
    if ((threadIdx.x & 31) == 0 && near_cnt) atomicAdd(&p.diag[2 * b + 1], near_cnt);
}


// ====================================================================================================
// Fast path of the batched entry point.  Two double atan2 per point made the exact kernels FP64-bound
// (46 us for 1.9 M points).  Here every point is first classified with fp32 angles; their error is
// bounded (fast_atan2 <= 5.9e-7 rad, plus 2.4e-7 for sqrtf and the pi/2 subtraction), so a
// point whose fp32 angle lies more than ANGLE_MARGIN from every bin edge is in the same bin as its fp64
// angle and never needs the fp64 evaluation; the others (~0.5 % of columns, ~0.2 % of rows) take the
// exact path unchanged.  The scan's theta min/max are found the same way: fp32 min/max first, then only
// the points within the margin of them are evaluated in fp64.  Results are bit-identical to the exact
// kernels (tests/test_gpu_project.py runs both against the reference's golden vectors).
// ====================================================================================================
constexpr double ANGLE_MARGIN = 8.0e-6;     // rad; >= 7x the fp32 angle error bound
constexpr int DEFER_CAP = 1024;             // per-block queue of near-edge points (overflow is handled inline)

// fp32 arctangent of the prefilter: atan(t) = t P(t^2) on [0,1] (degree-6 P, tools/fit_atan.py) plus octant folding,
// ~20 instructions against ~100 for atan2f with its special-case handling.  Certified maximum error 5.9e-7 rad over
// +-pi including a 2-ulp error of the approximate division (tools/fit_atan.py; tests/test_gpu_round2.py measures it on
// the device), i.e. 13x below ANGLE_MARGIN.  Zero, denormal, huge, infinite or NaN inputs return NaN, which no bin test
// accepts: such points always take the exact fp64 path.
__device__ __forceinline__ float fast_atan2(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    // fmaxf / fminf drop a NaN operand: both magnitudes are tested on their own (a comparison with NaN is false)
    if (!(ax < 1.0e30f) || !(ay < 1.0e30f) || !(mx > 1.0e-30f)) return __int_as_float(0x7fc00000);
    const float q = __fdividef(mn, mx);
    const float t = q * q;
    float a = fmaf(6.811646790e-03f, t, -3.360372957e-02f);
    a = fmaf(a, t, 7.962303474e-02f);
    a = fmaf(a, t, -1.323330193e-01f);
    a = fmaf(a, t, 1.980780307e-01f);
    a = fmaf(a, t, -3.331736634e-01f);
    a = fmaf(a, t, 9.999961108e-01f);
    float r = a * q;
    r = ay > ax ? 1.57079632679489662f - r : r;
    r = x < 0.f ? 3.14159265358979324f - r : r;
    return y < 0.f ? -r : r;
}

__device__ __noinline__ double exact_phi(const Pt q) { return atan2(q.y, q.x); }       // real calls: the fp64 arctangent would set the register budget of the point loops
__device__ __noinline__ double exact_theta(const Pt q) {
    const double rho = __dsqrt_rn(__dadd_rn(__dmul_rn(q.x, q.x), __dmul_rn(q.y, q.y)));
    return __dadd_rn(-atan2(rho, q.z), HALF_PI);
}

// certified bin from an fp32 angle, in fp32 (fp64 conversions and roundings run on a slow pipe): returns
// cnt = #{edges <= v}, or -1 if v is within the margin of an edge.  t = (a - start)/step carries a relative
// error of ~2e-7, i.e. < 5e-4 bins at 2048 bins; the margin in bin units is widened by 1e-3 to cover it.
struct FastEdges { float c0, inv_step, margin_bins, last; };

__device__ __forceinline__ FastEdges make_fast_edges(const Edges& e) {
    FastEdges f;
    const bool ok = e.step > 0.0 && e.num >= 2;
    const double inv = ok ? 1.0 / e.step : 0.0;
    f.inv_step = (float)inv;
    f.c0 = (float)(-e.start * inv);
    f.margin_bins = ok ? (float)(ANGLE_MARGIN * inv) + 1.0e-3f : 2.0f;     // 2 > 0.5: nothing passes when degenerate
    f.last = (float)(e.num - 2);
    return f;
}

__device__ __forceinline__ int fast_count_le(const FastEdges& f, float a32) {
    if (!(fabsf(a32) <= 4.0f)) return -1;
    const float t = fmaf(a32, f.inv_step, f.c0);
    const float fl = floorf(t);
    const float lo = t - fl;
    if (!(lo > f.margin_bins && lo < 1.0f - f.margin_bins) || fl < 0.0f || fl > f.last) return -1;
    return (int)fl + 1;
}

// ---- near-edge points decided by a float64 cross product instead of a float64 arctangent ------------------------------
// A point whose fp32 angle is within the margin of an edge e_k is on the upper side of that edge exactly when the cross
// product of its direction with the edge's direction is >= 0: y cos e_k - x sin e_k = rho sin(phi - e_k) for a column edge,
// z sin g_j - rho cos g_j = r sin(g_j - a) for a row edge (a = atan2(rho, z), theta = pi/2 - a, g_j = pi/2 - h_j).  The
// edge directions come from small shared-memory tables built once per block; the computed cross product is off by at
// most ~6e-16 (|x| + |y|) (two rounded products, one sum, table entries good to a few ulp, the table's angle within 1e-15
// of the float64 edge numpy builds), so outside a band of 1e-14 (|x| + |y|) its sign is the sign numpy's rounded arctangent
// (2 ulp) gives too -- checked on 8 M emulated points at offsets from 1e-17 to 1e-5 rad from an edge, zero disagreements
// outside the band.  Inside the band, on the outermost edges (the +-pi seam; the scan's own theta extremes, which sit ON
// their edges), or without a table (W > 8192, H > 256), the point takes the exact fp64 path through the queue as before,
// which is also where the 4-ulp `near` diagnostics come from.  With the test on, a handful of points per scan still queue
// up instead of ~0.5 % of them.
constexpr int COLT_MAX_A = 256;             // two-level column table: A[a] = dir(-pi + 32 a step), B[b] = dir(b step); W <= 8192
constexpr int ROWT_MAX = 256;               // row table: one entry per edge
constexpr double CROSS_BAND = 1.0e-14;

// Default: ON in the fixed-range fused kernel only.  Measured on B200: the tables' float64 sincos in every block's prologue
// cost what the shorter queue saves in the two-pass kernels (16 HDL-64 scans 0.0737 -> 0.0778 ms, OS1-128 0.1475 -> 0.1536 ms),
// while the fused kernel, whose queue carries BOTH angles, gains (0.1352 -> 0.1302 ms).  SLU_PROJECT_CROSS=1 switches it on
// everywhere, =0 off everywhere (bit-identical results in every setting: the projection tests pass with it on and off).
static int g_cross_mode = [] { const char* e = getenv("SLU_PROJECT_CROSS"); return (e && (e[0] == '0' || e[0] == '1')) ? e[0] - '0' : 2; }();   // 2 = fused kernel only

__device__ __forceinline__ void build_col_table(double2* A, double2* B, const Edges& ew, int W) {
    const int na = (W + 31) >> 5;
    for (int i = threadIdx.x; i < na + 32; i += blockDim.x) {
        const double ang = i < na ? __dadd_rn(__dmul_rn(32.0 * (double)i, ew.step), ew.start) : __dmul_rn((double)(i - na), ew.step);
        double sv, cv;
        sincos(ang, &sv, &cv);
        if (i < na) A[i] = make_double2(cv, sv); else B[i - na] = make_double2(cv, sv);
    }
}
// #{column edges <= phi} of a point near edge k = rint(t), or -1 when the cross product cannot tell
__device__ __forceinline__ int col_count_by_cross(const double2* A, const double2* B, const FastEdges& f, float phi32, double x, double y, int W) {
    if (!(fabsf(phi32) <= 4.0f)) return -1;
    const int k = __float2int_rn(fmaf(phi32, f.inv_step, f.c0));
    if (k < 1 || k > W - 2) return -1;
    const double2 a = A[k >> 5], bb = B[k & 31];
    const double C = a.x * bb.x - a.y * bb.y, S = a.y * bb.x + a.x * bb.y;
    const double cross = y * C - x * S;
    if (!(fabs(cross) > CROSS_BAND * (fabs(x) + fabs(y)))) return -1;
    return k + (cross >= 0.0 ? 1 : 0);
}
__device__ __forceinline__ void build_row_table(double2* R, const Edges& eh, int H) {
    for (int j = threadIdx.x; j < H; j += blockDim.x) {
        const double g = __dadd_rn(HALF_PI, -edge_at(eh, j));
        double sv, cv;
        sincos(g, &sv, &cv);
        R[j] = make_double2(cv, sv);
    }
}
__device__ __forceinline__ int row_count_by_cross(const double2* R, const FastEdges& f, float th32, const Pt q, int H) {
    if (!(fabsf(th32) <= 4.0f)) return -1;
    const int j = __float2int_rn(fmaf(th32, f.inv_step, f.c0));
    if (j < 1 || j > H - 2) return -1;
    const double rho = __dsqrt_rn(__dadd_rn(__dmul_rn(q.x, q.x), __dmul_rn(q.y, q.y)));
    const double2 g = R[j];
    const double cross = q.z * g.y - rho * g.x;
    if (!(fabs(cross) > CROSS_BAND * (rho + fabs(q.z)))) return -1;
    return j + (cross >= 0.0 ? 1 : 0);
}

__global__ void __launch_bounds__(PT_THREADS, 5) proj_fast_angles_kernel(const __grid_constant__ ProjParams p) {
    const int b = blockIdx.y;
    const long long n0 = p.offsets[b], n1 = p.offsets[b + 1];
    const bool check_ids = p.raw_label != nullptr && p.lut != nullptr;
    // the first point of every thread is requested before anything else, so its round trip overlaps the reset below
    const long long pt_stride = (long long)gridDim.x * blockDim.x;
    const long long n_first = n0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float4 v_nx = n_first < n1 ? __ldg(p.xyzi + n_first) : make_float4(1.f, 0.f, 0.f, 0.f);
    unsigned raw_nx = (check_ids && n_first < n1) ? __ldg(p.raw_label + n_first) : 0u;
    // this launch also resets the per-pixel depth-test state and the per-scan scalars
    if (p.use_cells) {
        ulonglong2* cell = reinterpret_cast<ulonglong2*>(p.cell + (long long)b * p.HW);
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.HW; i += (long long)gridDim.x * blockDim.x)
            cell[i] = make_ulonglong2(CELL_EMPTY, CELL_EMPTY);
    } else {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.HW; i += (long long)gridDim.x * blockDim.x) {
            p.key[(long long)b * p.HW + i] = ~0ull;
            p.winner[(long long)b * p.HW + i] = 0x7fffffff;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        p.tminmax[2 * b] = ~0ull; p.tminmax[2 * b + 1] = 0ull;
        p.diag[2 * b] = 0; p.diag[2 * b + 1] = 0;
    }
    const Edges ew = make_edges(-PI, PI, p.W);
    const FastEdges fw = make_fast_edges(ew);
    __shared__ int s_q[DEFER_CAP];
    __shared__ int s_qn;
    __shared__ double2 s_colA[COLT_MAX_A], s_colB[32];
    const bool col_cross = p.use_cross && p.W >= 4 && p.W <= 32 * COLT_MAX_A;
    if (col_cross) build_col_table(s_colA, s_colB, ew, p.W);
    if (threadIdx.x == 0) s_qn = 0;
    __syncthreads();
    float tmin = INFINITY, tmax = -INFINITY;
    int missing = 0, near_cnt = 0;
    int pend_lut = 0;
    // the point (and raw label) of the NEXT iteration is requested before the current one is worked on: with ~10 points per
    // thread the loop otherwise pays one exposed memory round trip per point
    for (long long n = n_first; n < n1; n += pt_stride) {
        const float4 v4 = v_nx;
        // label-id check (the reference raises KeyError on ids outside id_map, dataloader_semantic_KITTI.py:47): the raw
        // label is loaded together with the point, its LUT gather is issued here and only tested one iteration later,
        // so neither load stalls the angle arithmetic
        const unsigned raw = raw_nx & 0xffffu;
        if (n + pt_stride < n1) {
            v_nx = __ldg(p.xyzi + n + pt_stride);
            if (check_ids) raw_nx = __ldg(p.raw_label + n + pt_stride);
        }
        missing += pend_lut < 0 ? 1 : 0;
        const Pt q = load_pt(p, b, v4);
        const double x = q.x, y = q.y, z = q.z;
        const float4 v = make_float4((float)x, (float)y, (float)z, 0.f);     // fp32 view for the prefilter
        // depth key: the bits of r^2 = x^2 + y^2 + z^2 in numpy's operation order (utils.py:299 before its sqrt).  sqrt is
        // monotone, so ordering by r^2 is ordering by r except where two different r^2 round to the same r; the tie pass
        // treats those as the ties they are in the reference (proj_ties_kernel)
        const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
        const unsigned long long rb = (unsigned long long)__double_as_longlong(r2) & 0x7fffffffffffffffull;
        p.rkey[n] = p.farthest ? (0x7fffffffffffffffull - rb) : rb;
        const float phi32 = fast_atan2(v.y, v.x);
        const float th32 = 1.57079632679489662f - fast_atan2(sqrtf(fmaf(v.x, v.x, v.y * v.y)), v.z);
        p.theta32[n] = th32;
        pend_lut = check_ids ? __ldg(p.lut + raw) : 0;          // tested at the top of the next iteration
        if (th32 == th32) { tmin = fminf(tmin, th32); tmax = fmaxf(tmax, th32); }
        int cnt_w = fast_count_le(fw, phi32);
        if (cnt_w < 0 && col_cross) cnt_w = col_count_by_cross(s_colA, s_colB, fw, phi32, x, y, p.W);
        if (cnt_w < 0) {
            // near an edge and undecided: queue the point; the fp64 path runs once per block on the packed queue instead of
            // once per warp that happens to contain such a point
            // (a queue that overflows is dropped as a whole: the block then walks its points again, below)
            const int slot = atomicAdd(&s_qn, 1);
            if (slot < DEFER_CAP) s_q[slot] = (int)(n - n0);
            continue;
        }
        int c = p.W - 1 - cnt_w;                    // cnt_w in [0, W]: only -1 wraps (numpy's negative index)
        if (c < 0) c += p.W;
        p.col[n] = c;
    }
    missing += pend_lut < 0 ? 1 : 0;
    __syncthreads();
    if (s_qn <= DEFER_CAP) {
        for (int i = threadIdx.x; i < s_qn; i += blockDim.x) {
            const long long n = n0 + s_q[i];
            bool near;
            const int cnt_w = count_le(ew, exact_phi(load_pt(p, b, __ldg(p.xyzi + n))), near);
            if (near) ++near_cnt;
            int c = p.W - 1 - cnt_w;                    // cnt_w in [0, W]: only -1 wraps (numpy's negative index)
            if (c < 0) c += p.W;
            p.col[n] = c;
        }
    } else {
        // more near-edge points than the queue holds (adversarial inputs): every thread settles its own in place
        for (long long n = n0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; n < n1; n += (long long)gridDim.x * blockDim.x) {
            const Pt q = load_pt(p, b, __ldg(p.xyzi + n));
            const float a32 = fast_atan2((float)q.y, (float)q.x);
            if (fast_count_le(fw, a32) >= 0) continue;
            if (col_cross && col_count_by_cross(s_colA, s_colB, fw, a32, q.x, q.y, p.W) >= 0) continue;      // settled in the loop
            bool near;
            const int cnt_w = count_le(ew, exact_phi(q), near);
            if (near) ++near_cnt;
            int c = p.W - 1 - cnt_w;
            if (c < 0) c += p.W;
            p.col[n] = c;
        }
    }
    __shared__ float s_min[PT_THREADS / 32], s_max[PT_THREADS / 32];
    __shared__ int s_miss[PT_THREADS / 32], s_near[PT_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        tmin = fminf(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
        tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
        missing += __shfl_xor_sync(0xffffffffu, missing, o);
        near_cnt += __shfl_xor_sync(0xffffffffu, near_cnt, o);
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s_min[w] = tmin; s_max[w] = tmax; s_miss[w] = missing; s_near[w] = near_cnt; }
    __syncthreads();
    const long long slot = (long long)b * p.gx + blockIdx.x;           // every block writes its slot: no init needed
    float bmin = s_min[0], bmax = s_max[0];
    for (int i = 1; i < PT_THREADS / 32; ++i) { bmin = fminf(bmin, s_min[i]); bmax = fmaxf(bmax, s_max[i]); }
    if (threadIdx.x == 0) {
        for (int i = 1; i < PT_THREADS / 32; ++i) { missing += s_miss[i]; near_cnt += s_near[i]; }
        p.part_min[slot] = bmin; p.part_max[slot] = bmax;
        p.part_diag[2 * slot] = missing; p.part_diag[2 * slot + 1] = near_cnt;
    }
    if (p.use_range) return;
    // Exact fp64 theta min/max of the scan without a second pass over the points: the point P that attains the exact
    // minimum satisfies theta32(P) <= theta(P) + m/2 <= theta(Q) + m/2 <= theta32(Q) + m for every Q, in particular for
    // the Q of its own block, so it is among the block's points within the margin m of the BLOCK's fp32 minimum (same
    // for the maximum).  Each block evaluates those few candidates in fp64 and publishes its exact (min, max); the row
    // kernel reduces the gx partials.  (The separate extremes launch of the first version is gone.)
    const float m = (float)ANGLE_MARGIN;
    double emin = INFINITY, emax = -INFINITY;
    for (long long n = n0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; n < n1; n += (long long)gridDim.x * blockDim.x) {
        const float t = p.theta32[n];                      // written by this same thread above
        if (t <= bmin + m || t >= bmax - m || t != t) {
            const double th = exact_theta(load_pt(p, b, __ldg(p.xyzi + n)));
            if (th == th) { emin = fmin(emin, th); emax = fmax(emax, th); }
        }
    }
    __shared__ double s_emin[PT_THREADS / 32], s_emax[PT_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        emin = fmin(emin, __shfl_xor_sync(0xffffffffu, emin, o));
        emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
    }
    if ((threadIdx.x & 31) == 0) { s_emin[w] = emin; s_emax[w] = emax; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < PT_THREADS / 32; ++i) { emin = fmin(emin, s_emin[i]); emax = fmax(emax, s_emax[i]); }
        p.part_exact[2 * slot] = emin; p.part_exact[2 * slot + 1] = emax;
    }
}

// exact fp64 theta range of scan b from the angle kernel's per-block partials (block-wide call)
__device__ __forceinline__ void scan_theta_range_fast(const ProjParams& p, int b, double& lo, double& hi) {
    if (p.use_range) { lo = p.theta_lo; hi = p.theta_hi; return; }
    __shared__ double s_lo[PT_THREADS / 32], s_hi[PT_THREADS / 32];
    double a = INFINITY, c = -INFINITY;
    for (int i = threadIdx.x; i < p.gx; i += blockDim.x) {
        a = fmin(a, p.part_exact[2 * ((long long)b * p.gx + i)]);
        c = fmax(c, p.part_exact[2 * ((long long)b * p.gx + i) + 1]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a = fmin(a, __shfl_xor_sync(0xffffffffu, a, o));
        c = fmax(c, __shfl_xor_sync(0xffffffffu, c, o));
    }
    if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = a; s_hi[threadIdx.x >> 5] = c; }
    __syncthreads();
    a = s_lo[0]; c = s_hi[0];
    for (int i = 1; i < PT_THREADS / 32; ++i) { a = fmin(a, s_lo[i]); c = fmax(c, s_hi[i]); }
    if (!(a <= c)) { a = 0.0; c = 0.0; }                   // empty scan (or no finite angle)
    lo = a; hi = c;
}

template <bool CELLS>
__global__ void __launch_bounds__(PT_THREADS) proj_fast_rows_kernel(const __grid_constant__ ProjParams p) {
    const int b = blockIdx.y;
    const long long n0 = p.offsets[b], n1 = p.offsets[b + 1];
    double lo, hi;
    scan_theta_range_fast(p, b, lo, hi);
    __shared__ int s_q[DEFER_CAP];
    __shared__ int s_qn;
    __shared__ FastEdges s_fh;
    __shared__ double2 s_row[ROWT_MAX];
    const bool row_cross = p.use_cross && p.H >= 4 && p.H <= ROWT_MAX;
    if (row_cross && threadIdx.x < p.H) build_row_table(s_row, make_edges(lo, hi, p.H), p.H);
    if (threadIdx.x == 0) {                         // the float64 divisions of the edge set-up run once per block
        s_qn = 0;
        s_fh = make_fast_edges(make_edges(lo, hi, p.H));
    }
    __syncthreads();
    const FastEdges fh = s_fh;
    int near_cnt = 0;
    if (blockIdx.x == 0) {
        // fold the angle kernel's per-block diagnostics into the scan's counters (strided over the block's
        // threads: a single thread walking gx slots serially would be the longest chain in this kernel)
        if (threadIdx.x == 0 && p.theta_out) { p.theta_out[2 * b] = lo; p.theta_out[2 * b + 1] = hi; }
        int missing = 0;
        for (int i = threadIdx.x; i < p.gx; i += blockDim.x) {
            missing += p.part_diag[2 * ((long long)b * p.gx + i)];
            near_cnt += p.part_diag[2 * ((long long)b * p.gx + i) + 1];
        }
        missing = __reduce_add_sync(0xffffffffu, missing);
        if ((threadIdx.x & 31) == 0 && missing) atomicAdd(&p.diag[2 * b], missing);
    }
    // PT_BATCH points per thread and iteration: all their loads are issued before the first use, so a thread has
    // 3 * PT_BATCH memory requests in flight instead of paying one L2 round trip per point
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned long long* key = p.key + (long long)b * p.HW;
    Cell* cell = p.cell + (long long)b * p.HW;
    for (long long nb = n0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; nb < n1; nb += stride * PT_BATCH) {
        float t32[PT_BATCH];
        int col[PT_BATCH];
        unsigned long long rk[PT_BATCH];
#pragma unroll
        for (int k = 0; k < PT_BATCH; ++k) {
            const long long n = nb + k * stride;
            const bool in = n < n1;
            t32[k] = in ? p.theta32[n] : 0.f;
            col[k] = in ? p.col[n] : 0;
            rk[k] = in ? p.rkey[n] : 0ull;
        }
        int pxs[PT_BATCH];
        u128 old[PT_BATCH];
#pragma unroll
        for (int k = 0; k < PT_BATCH; ++k) {
            const long long n = nb + k * stride;
            pxs[k] = -1;
            if (n >= n1) continue;
            int cnt_h = fast_count_le(fh, t32[k]);
            if (cnt_h < 0 && row_cross) cnt_h = row_count_by_cross(s_row, fh, t32[k], load_pt(p, b, __ldg(p.xyzi + n)), p.H);
            if (cnt_h < 0) {
                const int slot = atomicAdd(&s_qn, 1);
                if (slot < DEFER_CAP) s_q[slot] = (int)(n - n0);
                continue;
            }
            int r = p.H - 1 - cnt_h;                    // cnt_h in [0, H]: only -1 wraps
            if (r < 0) r += p.H;
            const int px = r * p.W + col[k];
            pxs[k] = px;
            p.pix[n] = px;
            if (CELLS) old[k] = depth_test_issue(cell + px, rk[k], (unsigned long long)(n - n0));    // all first attempts of the batch in flight
            else atomicMin(&key[px], rk[k]);
        }
        if (CELLS) {
#pragma unroll
            for (int k = 0; k < PT_BATCH; ++k)
                if (pxs[k] >= 0) depth_test_settle(p, cell + pxs[k], rk[k], (unsigned long long)(nb + k * stride - n0), old[k]);
        }
    }
    __syncthreads();
    const Edges eh = make_edges(lo, hi, p.H);       // float64 edges: only the queued points need them
    const bool overflow = s_qn > DEFER_CAP;         // then every thread walks its points again and settles the near-edge ones
    const long long q_begin = overflow ? (long long)blockIdx.x * blockDim.x + threadIdx.x : threadIdx.x;
    const long long q_end = overflow ? n1 - n0 : s_qn, q_step = overflow ? stride : blockDim.x;
    for (long long i = q_begin; i < q_end; i += q_step) {
        const long long n = n0 + (overflow ? i : (long long)s_q[i]);
        if (overflow && fast_count_le(fh, p.theta32[n]) >= 0) continue;
        if (overflow && row_cross && row_count_by_cross(s_row, fh, p.theta32[n], load_pt(p, b, __ldg(p.xyzi + n)), p.H) >= 0) continue;   // settled in the loop
        bool near;
        const double th = exact_theta(load_pt(p, b, __ldg(p.xyzi + n)));
        const int cnt_h = count_le(eh, th, near);
        if (near && !p.use_range) near = !(th == lo || th == hi);
        if (near) ++near_cnt;
        int r = p.H - 1 - cnt_h;                    // cnt_h in [0, H]: only -1 wraps
        if (r < 0) r += p.H;
        const int px = r * p.W + p.col[n];
        p.pix[n] = px;
        if (CELLS) depth_test_settle(p, cell + px, p.rkey[n], (unsigned long long)(n - n0), depth_test_issue(cell + px, p.rkey[n], (unsigned long long)(n - n0)));
        else atomicMin(&p.key[(long long)b * p.HW + px], p.rkey[n]);
    }
    near_cnt = __reduce_add_sync(0xffffffffu, near_cnt);
    if ((threadIdx.x & 31) == 0 && near_cnt) atomicAdd(&p.diag[2 * b + 1], near_cnt);
}

// Fixed elevation range (theta_range given: SemanticCUDAL +-pi/8, SemanticWADS +-pi/2, src/dataset/dataloader_semantic_CUDAL.py:95):
// the row edges do not depend on the scan, so column, row and the depth test run in ONE pass over the points -- no fp32
// theta / column round trip through memory, and the 64-bit atomic stream overlaps the angle arithmetic.  The per-pixel
// state is reset by proj_init_kernel beforehand (the depth test cannot share a launch with its own initialisation).
template <bool CELLS>
__global__ void __launch_bounds__(PT_THREADS) proj_fast_fused_kernel(const __grid_constant__ ProjParams p) {
    const int b = blockIdx.y;
    const long long n0 = p.offsets[b], n1 = p.offsets[b + 1];
    const Edges ew = make_edges(-PI, PI, p.W);
    const Edges eh = make_edges(p.theta_lo, p.theta_hi, p.H);
    const FastEdges fw = make_fast_edges(ew), fh = make_fast_edges(eh);
    __shared__ int s_q[DEFER_CAP];
    __shared__ int s_qn;
    __shared__ double2 s_colA[COLT_MAX_A], s_colB[32], s_row[ROWT_MAX];
    const bool col_cross = p.use_cross && p.W >= 4 && p.W <= 32 * COLT_MAX_A;
    const bool row_cross = p.use_cross && p.H >= 4 && p.H <= ROWT_MAX && eh.step > 0.0;
    if (col_cross) build_col_table(s_colA, s_colB, ew, p.W);
    if (row_cross) build_row_table(s_row, eh, p.H);
    if (threadIdx.x == 0) s_qn = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0 && p.theta_out) { p.theta_out[2 * b] = p.theta_lo; p.theta_out[2 * b + 1] = p.theta_hi; }
    __syncthreads();
    int missing = 0, near_cnt = 0;
    const bool check_ids = p.raw_label != nullptr && p.lut != nullptr;
    int pend_lut = 0;
    unsigned long long* key = p.key + (long long)b * p.HW;
    Cell* cell = p.cell + (long long)b * p.HW;
    const long long pt_stride = (long long)gridDim.x * blockDim.x;       // next iteration's point requested ahead, as in the angle kernel
    const long long n_first = n0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float4 v_nx = n_first < n1 ? __ldg(p.xyzi + n_first) : make_float4(1.f, 0.f, 0.f, 0.f);
    unsigned raw_nx = (check_ids && n_first < n1) ? __ldg(p.raw_label + n_first) : 0u;
    for (long long n = n_first; n < n1; n += pt_stride) {
        const float4 v4 = v_nx;
        const unsigned raw = raw_nx & 0xffffu;
        if (n + pt_stride < n1) {
            v_nx = __ldg(p.xyzi + n + pt_stride);
            if (check_ids) raw_nx = __ldg(p.raw_label + n + pt_stride);
        }
        missing += pend_lut < 0 ? 1 : 0;
        const Pt q = load_pt(p, b, v4);
        const double x = q.x, y = q.y, z = q.z;
        const float4 v = make_float4((float)x, (float)y, (float)z, 0.f);
        const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));     // r^2 key, see proj_fast_angles_kernel
        const unsigned long long rb = (unsigned long long)__double_as_longlong(r2) & 0x7fffffffffffffffull;
        const unsigned long long rk = p.farthest ? (0x7fffffffffffffffull - rb) : rb;
        p.rkey[n] = rk;                                            // the tie pass compares against it
        const float phi32 = fast_atan2(v.y, v.x);
        const float th32 = 1.57079632679489662f - fast_atan2(sqrtf(fmaf(v.x, v.x, v.y * v.y)), v.z);
        pend_lut = check_ids ? __ldg(p.lut + raw) : 0;
        int cnt_w = fast_count_le(fw, phi32);
        int cnt_h = fast_count_le(fh, th32);
        if (cnt_w < 0 && col_cross) cnt_w = col_count_by_cross(s_colA, s_colB, fw, phi32, x, y, p.W);
        if (cnt_h < 0 && row_cross) cnt_h = row_count_by_cross(s_row, fh, th32, q, p.H);
        if (cnt_w < 0 || cnt_h < 0) {
            const int slot = atomicAdd(&s_qn, 1);
            if (slot < DEFER_CAP) s_q[slot] = (int)(n - n0);
            continue;
        }
        int c = p.W - 1 - cnt_w;
        if (c < 0) c += p.W;
        int rr = p.H - 1 - cnt_h;
        if (rr < 0) rr += p.H;
        const int px = rr * p.W + c;
        p.pix[n] = px;
        if (CELLS) depth_test_settle(p, cell + px, rk, (unsigned long long)(n - n0), depth_test_issue(cell + px, rk, (unsigned long long)(n - n0)));
        else atomicMin(&key[px], rk);
    }
    missing += pend_lut < 0 ? 1 : 0;
    __syncthreads();
    const bool overflow = s_qn > DEFER_CAP;         // then every thread walks its points again and settles the near-edge ones
    const long long q_begin = overflow ? (long long)blockIdx.x * blockDim.x + threadIdx.x : threadIdx.x;
    const long long q_end = overflow ? n1 - n0 : s_qn, q_step = overflow ? (long long)gridDim.x * blockDim.x : blockDim.x;
    for (long long i = q_begin; i < q_end; i += q_step) {
        // queued points: both bins from the exact fp64 angles (a bin the prefilter had certified cannot be near an edge)
        const long long n = n0 + (overflow ? i : (long long)s_q[i]);
        const Pt q = load_pt(p, b, __ldg(p.xyzi + n));
        if (overflow) {
            const float x32 = (float)q.x, y32 = (float)q.y, z32 = (float)q.z;
            const float a32 = fast_atan2(y32, x32), t32 = 1.57079632679489662f - fast_atan2(sqrtf(fmaf(x32, x32, y32 * y32)), z32);
            const bool w_ok = fast_count_le(fw, a32) >= 0 || (col_cross && col_count_by_cross(s_colA, s_colB, fw, a32, q.x, q.y, p.W) >= 0);
            const bool h_ok = fast_count_le(fh, t32) >= 0 || (row_cross && row_count_by_cross(s_row, fh, t32, q, p.H) >= 0);
            if (w_ok && h_ok) continue;                       // settled in the loop
        }
        bool near_w, near_h;
        const int cnt_w = count_le(ew, exact_phi(q), near_w);
        const int cnt_h = count_le(eh, exact_theta(q), near_h);
        near_cnt += (near_w ? 1 : 0) + (near_h ? 1 : 0);
        int c = p.W - 1 - cnt_w;
        if (c < 0) c += p.W;
        int rr = p.H - 1 - cnt_h;
        if (rr < 0) rr += p.H;
        const int px = rr * p.W + c;
        p.pix[n] = px;
        if (CELLS) depth_test_settle(p, cell + px, p.rkey[n], (unsigned long long)(n - n0), depth_test_issue(cell + px, p.rkey[n], (unsigned long long)(n - n0)));
        else atomicMin(&key[px], p.rkey[n]);
    }
    missing = __reduce_add_sync(0xffffffffu, missing);
    near_cnt = __reduce_add_sync(0xffffffffu, near_cnt);
    if ((threadIdx.x & 31) == 0) {
        if (missing) atomicAdd(&p.diag[2 * b], missing);
        if (near_cnt) atomicAdd(&p.diag[2 * b + 1], near_cnt);
    }
}

__global__ void __launch_bounds__(PT_THREADS) proj_ties_kernel(const __grid_constant__ ProjParams p) {
    const int b = blockIdx.y;
    const long long n0 = p.offsets[b], n1 = p.offsets[b + 1];
    const long long stride = (long long)gridDim.x * blockDim.x;
    const unsigned long long* key = p.key + (long long)b * p.HW;
    int* winner = p.winner + (long long)b * p.HW;
    for (long long nb = n0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; nb < n1; nb += stride * PT_BATCH) {
        int px[PT_BATCH];
        unsigned long long rk[PT_BATCH], best[PT_BATCH];
#pragma unroll
        for (int k = 0; k < PT_BATCH; ++k) {
            const long long n = nb + k * stride;
            const bool in = n < n1;
            px[k] = in ? p.pix[n] : 0;
            rk[k] = in ? p.rkey[n] : 0ull;
        }
#pragma unroll
        for (int k = 0; k < PT_BATCH; ++k) best[k] = (nb + k * stride < n1) ? key[px[k]] : ~0ull;
#pragma unroll
        for (int k = 0; k < PT_BATCH; ++k) {
            const long long n = nb + k * stride;
            if (n < n1 && same_range(p, rk[k], best[k])) atomicMin(&winner[px[k]], (int)(n - n0));
        }
    }
}


// ====================================================================================================
// Cell pipeline (round 2; an A/B alternative selected with slu_debug_project_exact(2), NOT the default): THREE launches, one
// pass of atomics.
//
//   P1 extremes  cells <- empty; per block the fp32 tangent of the elevation z / rho (monotone in theta: no arctangent)
//                gives the block's extreme candidates, which alone are evaluated in fp64 -> exact (min, max) partials
//                (with a fixed elevation range P1 only resets the cells)
//   P2 points    ONE fused pass per point: fp32-prefiltered column and row (fp64 near edges, as above), then the depth
//                test as a 128-bit compare-and-swap on the pixel's (range key, point index) cell.  The cell orders
//                points by (float64 range, index), so the winner INCLUDING the lowest-index rule for exact ties is
//                settled by this pass: no separate tie pass, no second gather of per-point state (theta32 / col / rkey
//                never exist in memory).
//   P3 resolve   per pixel: read the cell, gather the winner, write the planes.
//
// The first CAS of a point assumes an empty cell, so the common case (first point of its pixel) costs one L2 round
// trip; a failed CAS returns the resident (key, index), the point retires if the resident one is better and retries
// against the value it saw otherwise.  The cell only ever moves down the total order, whatever the interleaving: the
// result is deterministic and bit-identical to the exact kernels (tests/test_gpu_project.py).
//
// Measured on B200 (16 HDL-64 scans, graph replay): 0.076-0.078 ms against 0.076 ms for the four-launch path, and 0.158
// against 0.135 ms with a fixed range -- the pass it deletes (ties, 11 us) and the per-point state it never writes are
// paid back by the fused kernel: ATOMG.CAS.128 costs what RED.MIN.64 costs (the same 45.7 us kernel with either, 30.3 us
// with no atomic at all: ~15 us per 1.9 M L2 atomics whatever their width), the fused loop needs 48-64 registers (4-5 CTAs
// per SM) and its blocks finish 12 us apart (31.8 ... 44.5 us after a common start; tools/proj_timeline.py), so the
// slowest SMs set the time.  It stays in the library as the measured alternative (profiles/projection_r02.md).
// ====================================================================================================
// fp32 tangent of the elevation, z / sqrt(x^2 + y^2): relative error <= 5e-7 (two rounded products and a sum, rsqrt.approx
// <= 2 ulp, one product, plus the float32 view of yaw-rotated coordinates).  NaN when the operands leave the range in
// which that bound holds (rho^2 outside [1e-30, 1e30], |t| >= 1e9, non-finite input): such points are always candidates.
__device__ __forceinline__ float tan_elev32(float x, float y, float z) {
    const float s = fmaf(x, x, y * y);
    const float t = z * rsqrtf(s);
    const bool ok = s > 1.0e-30f && s < 1.0e30f && fabsf(t) < 1.0e9f;
    return ok ? t : __int_as_float(0x7fc00000);
}
// |t~ - t| <= TAN_BAND(t~): 4x the fp32 bound above, plus an absolute floor that covers the rounding of the float64 theta
// itself near the horizon (theta carries ~4e-16 absolute; d theta / d t = 1 / (1 + t^2))
__device__ __forceinline__ float tan_band(float t) { return fmaf(2.0e-6f, fabsf(t), 1.0e-9f); }

__device__ __forceinline__ void dbg_stamp(const ProjParams& p, int kernel, int slot) {
    if (p.dbg_times && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        unsigned long long* row = p.dbg_times + ((long long)kernel * 4096 + blockIdx.y * gridDim.x + blockIdx.x) * 8;
        row[slot] = t;
        if (slot == 0) {                             // slot 7: which SM ran the block
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            row[7] = smid + 1;
        }
    }
}
constexpr int P1_NB = 4;                    // points a thread of P1 has in flight per iteration

// float32 coordinates of a point as the prefilter sees them (float32(rotated float64 coordinate) under a yaw)
__device__ __forceinline__ void coords32(const ProjParams& p, int b, const float4 v, float& x, float& y, float& z) {
    x = v.x; y = v.y; z = v.z;
    if (p.yaw) {
        const Pt q = load_pt(p, b, v);
        x = (float)q.x; y = (float)q.y;
    }
}

__global__ void __launch_bounds__(PT_THREADS) proj3_extremes_kernel(const __grid_constant__ ProjParams p) {
    const int b = blockIdx.y;
    const long long n0 = p.offsets[b];
    const int npts = (int)(p.offsets[b + 1] - n0);
    const int tid0 = blockIdx.x * PT_THREADS + threadIdx.x, stride = gridDim.x * PT_THREADS;
    dbg_stamp(p, 0, 0);
    {
        ulonglong2* cell = reinterpret_cast<ulonglong2*>(p.cell + (long long)b * p.HW);
        const ulonglong2 empty = make_ulonglong2(CELL_EMPTY, CELL_EMPTY);
        const int hw = (int)p.HW;
        for (int i = tid0; i < hw; i += stride) cell[i] = empty;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { p.diag[2 * b] = 0; p.diag[2 * b + 1] = 0; }
    if (p.use_range) return;
    dbg_stamp(p, 0, 1);
    const float4* __restrict__ pts = p.xyzi + n0;
    // pass A: every thread keeps the smallest and second smallest (largest, second largest) fp32 tangent of its points
    // and the index of the extreme one; points whose tangent is not a number are always candidates and are evaluated in
    // fp64 on the spot
    float m1 = INFINITY, m2 = INFINITY, M1 = -INFINITY, M2 = -INFINITY;
    int i1 = -1, j1 = -1;
    double emin = INFINITY, emax = -INFINITY;
    for (int i = tid0; i < npts; i += stride * P1_NB) {
        float4 v[P1_NB];
#pragma unroll
        for (int k = 0; k < P1_NB; ++k) {
            const int idx = i + k * stride;
            v[k] = idx < npts ? __ldg(pts + idx) : make_float4(1.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < P1_NB; ++k) {
            const int idx = i + k * stride;
            if (idx >= npts) continue;
            float x, y, z;
            coords32(p, b, v[k], x, y, z);
            const float t = tan_elev32(x, y, z);
            if (t != t) {
                const double th = exact_theta(load_pt(p, b, v[k]));
                if (th == th) { emin = fmin(emin, th); emax = fmax(emax, th); }
                continue;
            }
            if (t < m1) { m2 = m1; m1 = t; i1 = idx; } else if (t < m2) m2 = t;
            if (t > M1) { M2 = M1; M1 = t; j1 = idx; } else if (t > M2) M2 = t;
        }
    }
    dbg_stamp(p, 0, 2);
    __shared__ float s_min[PT_THREADS / 32], s_max[PT_THREADS / 32];
    float tmin = m1, tmax = M1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        tmin = fminf(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
        tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s_min[w] = tmin; s_max[w] = tmax; }
    __syncthreads();
    float bmin = s_min[0], bmax = s_max[0];
    for (int i = 1; i < PT_THREADS / 32; ++i) { bmin = fminf(bmin, s_min[i]); bmax = fmaxf(bmax, s_max[i]); }
    // The point P attaining the scan's exact minimum has t(P) <= t(Q) for every Q (theta is monotone in t = tan theta), so
    // t~(P) - band <= t(P) <= t(Q) <= t~(Q) + band for the Q that set this block's fp32 minimum: P passes the test below in
    // its own block.  Usually only a thread's extreme point can pass; when a thread's SECOND point passes as well, the
    // block walks its points again and evaluates every one that passes.
    const float thr_lo = bmin + tan_band(bmin), thr_hi = bmax - tan_band(bmax);
    const bool c_lo = i1 >= 0 && !(m1 - tan_band(m1) > thr_lo), c_hi = j1 >= 0 && !(M1 + tan_band(M1) < thr_hi);
    const bool more = (m2 < INFINITY && !(m2 - tan_band(m2) > thr_lo)) || (M2 > -INFINITY && !(M2 + tan_band(M2) < thr_hi));
    if (!__syncthreads_or(more ? 1 : 0)) {
        if (c_lo) {
            const double th = exact_theta(load_pt(p, b, __ldg(pts + i1)));
            if (th == th) { emin = fmin(emin, th); emax = fmax(emax, th); }
        }
        if (c_hi && !(c_lo && j1 == i1)) {
            const double th = exact_theta(load_pt(p, b, __ldg(pts + j1)));
            if (th == th) { emin = fmin(emin, th); emax = fmax(emax, th); }
        }
    } else {
        for (int i = tid0; i < npts; i += stride) {
            const float4 v = __ldg(pts + i);
            float x, y, z;
            coords32(p, b, v, x, y, z);
            const float t = tan_elev32(x, y, z);
            if (t != t) continue;                               // already evaluated in pass A
            const float band = tan_band(t);
            if (!(t - band > thr_lo) || !(t + band < thr_hi)) {
                const double th = exact_theta(load_pt(p, b, v));
                if (th == th) { emin = fmin(emin, th); emax = fmax(emax, th); }
            }
        }
    }
    __shared__ double s_emin[PT_THREADS / 32], s_emax[PT_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        emin = fmin(emin, __shfl_xor_sync(0xffffffffu, emin, o));
        emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
    }
    if ((threadIdx.x & 31) == 0) { s_emin[w] = emin; s_emax[w] = emax; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < PT_THREADS / 32; ++i) { emin = fmin(emin, s_emin[i]); emax = fmax(emax, s_emax[i]); }
        const long long slot = (long long)b * p.gx + blockIdx.x;
        p.part_exact[2 * slot] = emin; p.part_exact[2 * slot + 1] = emax;
    }
    dbg_stamp(p, 0, 3);
}

// bins of one point from its exact fp64 angles (the queue / overflow path of P2)
// (a real call, not inlined: the fp64 arctangents would otherwise set the register budget of the whole point kernel)
__device__ __noinline__ int exact_pixel_call(double x, double y, double z, int cnt_w, int cnt_h, double lo, double hi, int W, int H,
                                             int use_range, int* near_cnt) {
    // the float64 edges are rebuilt here, on the rare path, rather than held in registers across the point loop
    Pt q; q.x = x; q.y = y; q.z = z;
    bool near;
    int nc = 0;
    if (cnt_w < 0) {
        cnt_w = count_le(make_edges(-PI, PI, W), exact_phi(q), near);
        if (near) ++nc;
    }
    if (cnt_h < 0) {
        const double th = exact_theta(q);
        cnt_h = count_le(make_edges(lo, hi, H), th, near);
        if (near && !use_range) near = !(th == lo || th == hi);         // the scan's own extremes sit ON the outer edges
        if (near) ++nc;
    }
    *near_cnt += nc;
    int c = W - 1 - cnt_w;                          // cnt in [0, n]: only -1 wraps (numpy's negative index)
    if (c < 0) c += W;
    int rr = H - 1 - cnt_h;
    if (rr < 0) rr += H;
    return rr * W + c;
}
// cnt_w / cnt_h >= 0: that bin is already certified by the prefilter and only the other angle is evaluated
__device__ __forceinline__ int exact_pixel(const ProjParams& p, const Pt q, int cnt_w, int cnt_h, double lo, double hi, int& near_cnt) {
    int nc = 0;
    const int px = exact_pixel_call(q.x, q.y, q.z, cnt_w, cnt_h, lo, hi, p.W, p.H, p.use_range, &nc);
    near_cnt += nc;
    return px;
}

// depth key of a point: the bits of r^2 = (x^2 + y^2) + z^2 in numpy's operation order (utils.py:299 before its sqrt; the
// tie rule of cell_before restores the order by r).  File coordinates are float32 values, whose squares are exact in
// float64, so fl(x^2 + y^2) and fl(that + z^2) are single-rounding fused operations; yaw-rotated coordinates are not
// float32 values and take the five separately rounded operations.  fmask = 0 (nearest wins) or 0x7fff... (farthest wins:
// MAX - bits == MAX ^ bits).
template <bool YAW>
__device__ __forceinline__ unsigned long long range_key(const ProjParams& p, int b, const float4 v, unsigned long long fmask) {
    double r2;
    if (YAW) {
        const Pt q = load_pt(p, b, v);
        r2 = __dadd_rn(__dadd_rn(__dmul_rn(q.x, q.x), __dmul_rn(q.y, q.y)), __dmul_rn(q.z, q.z));
    } else {
        const double x = (double)v.x, y = (double)v.y, z = (double)v.z;
        r2 = __fma_rn(z, z, __fma_rn(x, x, __dmul_rn(y, y)));
    }
    return ((unsigned long long)__double_as_longlong(r2) & 0x7fffffffffffffffull) ^ fmask;
}

// fast_atan2 without its operand checks: the caller guarantees finite operands with max(|x|, |y|) in [1e-15, 1e30)
__device__ __forceinline__ float fast_atan2_nc(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float q = __fdividef(mn, mx);
    const float t = q * q;
    float a = fmaf(6.811646790e-03f, t, -3.360372957e-02f);
    a = fmaf(a, t, 7.962303474e-02f);
    a = fmaf(a, t, -1.323330193e-01f);
    a = fmaf(a, t, 1.980780307e-01f);
    a = fmaf(a, t, -3.331736634e-01f);
    a = fmaf(a, t, 9.999961108e-01f);
    float r = a * q;
    r = ay > ax ? 1.57079632679489662f - r : r;
    r = x < 0.f ? 3.14159265358979324f - r : r;
    return y < 0.f ? -r : r;
}
__device__ __forceinline__ float sqrt_approx(float s) {          // MUFU.SQRT, relative error <= 2^-23
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s));
    return r;
}
// fast_count_le for an angle that is known to be finite
__device__ __forceinline__ int fast_count_le_nc(const FastEdges& f, float a32) {
    const float t = fmaf(a32, f.inv_step, f.c0);
    const float fl = floorf(t);
    const float lo = t - fl;
    const bool ok = lo > f.margin_bins && lo < 1.0f - f.margin_bins && fl >= 0.0f && fl <= f.last;
    return ok ? (int)fl + 1 : -1;
}

// MINB resident CTAs per SM (the register budget); YAW: the loaders' yaw augmentation is applied; IDS: raw label ids are
// checked against the look-up table.  The point loop is a three-stage software pipeline per thread: the point two
// iterations ahead is being LOADED, the next point's bins and key are COMPUTED, and the current point's compare-and-swap
// -- issued at the top of the iteration -- is only examined at the bottom, after that arithmetic, so neither the load nor
// the atomic round trip is waited for.  (Examining the result straight away, or carrying it into the next iteration,
// costs a full L2 round trip per point: ptxas copies a carried 128-bit result out of its register quad at once.)
template <int NB, int MINB, bool YAW, bool IDS>      // NB: unused (kept for the launch macro)
__global__ void __launch_bounds__(PT_THREADS, MINB) proj3_points_kernel(const __grid_constant__ ProjParams p) {
    const int b = blockIdx.y;
    const long long n0 = p.offsets[b];
    const int npts = (int)(p.offsets[b + 1] - n0);
    __shared__ int s_q[DEFER_CAP];
    __shared__ unsigned s_qc[DEFER_CAP];            // the bin counts the prefilter DID certify (0xffff = not certified), cnt_w | cnt_h << 16
    __shared__ int s_qn;
    __shared__ FastEdges s_fw, s_fh;
    __shared__ double s_lohi[2];
    dbg_stamp(p, 1, 0);
    {
        double lo, hi;
        scan_theta_range_fast(p, b, lo, hi);
        if (threadIdx.x == 0) {                     // the float64 divisions of the edge set-up run once per block
            s_qn = 0;
            s_lohi[0] = lo; s_lohi[1] = hi;
            s_fw = make_fast_edges(make_edges(-PI, PI, p.W));
            s_fh = make_fast_edges(make_edges(lo, hi, p.H));
            if (blockIdx.x == 0 && p.theta_out) { p.theta_out[2 * b] = lo; p.theta_out[2 * b + 1] = hi; }
        }
    }
    __syncthreads();
    dbg_stamp(p, 1, 1);
    const FastEdges fw = s_fw, fh = s_fh;
    int missing = 0, near_cnt = 0;
    const float4* __restrict__ pts = p.xyzi + n0;
    const unsigned* __restrict__ labs = p.raw_label + (IDS ? n0 : 0);
    int* __restrict__ pix = p.pix + n0;
    Cell* cell = p.cell + (long long)b * p.HW;
    const int W = p.W, H = p.H;
    const unsigned long long fmask = p.farthest ? 0x7fffffffffffffffull : 0ull;
    const int tid0 = blockIdx.x * PT_THREADS + threadIdx.x, stride = gridDim.x * PT_THREADS;
    const u128 EMPTY128 = ~(u128)0;

    // bins of one point from fp32 angles; -1 and a queue entry when it is near an edge
    auto classify = [&](int idx, const float4 v) -> int {
        float x = v.x, y = v.y;
        const float z = v.z;
        if (YAW) {
            const Pt q = load_pt(p, b, v);
            x = (float)q.x; y = (float)q.y;
        }
        // one operand check for both arctangents; whatever fails it takes the fp64 path through the queue
        const float s = fmaf(x, x, y * y);
        int cnt_w = -1, cnt_h = -1;
        if (s > 1.0e-30f && s < 1.0e30f && fabsf(z) < 1.0e30f) {
            cnt_w = fast_count_le_nc(fw, fast_atan2_nc(y, x));
            cnt_h = fast_count_le_nc(fh, 1.57079632679489662f - fast_atan2_nc(sqrt_approx(s), z));
        }
        if ((cnt_w | cnt_h) < 0) {
            // near an edge: queue the point, the fp64 angle runs once per block on the packed queue (a queue that
            // overflows is dropped as a whole: the block then re-scans its points for the near-edge ones, below)
            const int slot = atomicAdd(&s_qn, 1);
            if (slot < DEFER_CAP) { s_q[slot] = idx; s_qc[slot] = ((unsigned)cnt_w & 0xffffu) | ((unsigned)cnt_h << 16); }
            return -1;
        }
        int c = W - 1 - cnt_w;                                   // cnt in [0, n]: only -1 wraps (numpy's negative index)
        if (c < 0) c += W;
        int rr = H - 1 - cnt_h;
        if (rr < 0) rr += H;
        return rr * W + c;
    };

    int idx = tid0, idx_nx = tid0 + stride;
    bool have = idx < npts;
    float4 v = have ? __ldg(pts + idx) : make_float4(1.f, 0.f, 0.f, 0.f);
    unsigned raw = (IDS && have) ? (__ldg(labs + idx) & 0xffffu) : 0u;
    float4 v_nx = idx_nx < npts ? __ldg(pts + idx_nx) : make_float4(1.f, 0.f, 0.f, 0.f);
    unsigned raw_nx = (IDS && idx_nx < npts) ? (__ldg(labs + idx_nx) & 0xffffu) : 0u;
    int px = have ? classify(idx, v) : -1;
    unsigned long long rk = px >= 0 ? range_key<YAW>(p, b, v, fmask) : 0ull;
    while (have) {
        // stage 3 of the current point: issue its depth test and its label look-up
        u128 old = EMPTY128;
        if (px >= 0) {
            if (p.want_pix) pix[idx] = px;
            if (p.dbg_atom == 0) old = cas128q(cell + px, EMPTY128, ((u128)rk << 64) | (unsigned long long)(unsigned)idx);
            else if (p.dbg_atom == 1) atomicMin(&cell[px].key, rk);
        }
        int lutv = 0;
        if (IDS) lutv = __ldg(p.lut + raw);
        // stage 1 of the point two ahead, stage 2 of the next point
        const bool have_nx = idx_nx < npts;
        const float4 v2 = v_nx;
        const unsigned raw2 = raw_nx;
        const int idx_nn = idx_nx + stride;
        v_nx = idx_nn < npts ? __ldg(pts + idx_nn) : make_float4(1.f, 0.f, 0.f, 0.f);
        if (IDS) raw_nx = idx_nn < npts ? (__ldg(labs + idx_nn) & 0xffffu) : 0u;
        const int px2 = have_nx ? classify(idx_nx, v2) : -1;
        const unsigned long long rk2 = px2 >= 0 ? range_key<YAW>(p, b, v2, fmask) : 0ull;
        // the current point's results are back by now
        if (IDS) missing += lutv < 0 ? 1 : 0;                    // KeyError of the reference's remap loop
        if (old != EMPTY128)
            depth_test_retry(cell + px, rk, (unsigned long long)(unsigned)idx, (unsigned long long)old, (unsigned long long)(old >> 64), p.key_sq, p.farthest);
        idx = idx_nx; idx_nx = idx_nn; px = px2; rk = rk2; raw = raw2; have = have_nx;
    }
    dbg_stamp(p, 1, 2);
    __syncthreads();
    dbg_stamp(p, 1, 3);
    const int qn = s_qn;
    const double lo = s_lohi[0], hi = s_lohi[1];
    if (qn <= DEFER_CAP) {
        for (int i = threadIdx.x; i < qn; i += PT_THREADS) {
            const int idx = s_q[i];
            const unsigned qc = s_qc[i];
            const float4 v = __ldg(pts + idx);
            // only the angle the prefilter could not certify is evaluated in fp64 (a certified bin is not near an edge)
            const int cw = (qc & 0xffffu) == 0xffffu || W > 0xfffe ? -1 : (int)(qc & 0xffffu);
            const int ch = (qc >> 16) == 0xffffu || H > 0xfffe ? -1 : (int)(qc >> 16);
            const int px = exact_pixel(p, load_pt(p, b, v), cw, ch, lo, hi, near_cnt);
            if (p.want_pix) pix[idx] = px;
            unsigned long long ol, oh;
            const unsigned long long rk = range_key<YAW>(p, b, v, fmask);
            cas128(cell + px, CELL_EMPTY, CELL_EMPTY, (unsigned long long)idx, rk, ol, oh);
            depth_test_settle(p, cell + px, rk, (unsigned long long)idx, ol, oh);
        }
    } else {
        // more near-edge points than the queue holds (adversarial inputs): every thread walks its points again and
        // settles the near-edge ones in place
        for (int idx = tid0; idx < npts; idx += stride) {
            const float4 v = __ldg(pts + idx);
            float x = v.x, y = v.y;
            if (YAW) {
                const Pt q = load_pt(p, b, v);
                x = (float)q.x; y = (float)q.y;
            }
            const float s = fmaf(x, x, y * y);
            if (s > 1.0e-30f && s < 1.0e30f && fabsf(v.z) < 1.0e30f &&
                fast_count_le_nc(fw, fast_atan2_nc(y, x)) >= 0 &&
                fast_count_le_nc(fh, 1.57079632679489662f - fast_atan2_nc(sqrt_approx(s), v.z)) >= 0) continue;
            const int px = exact_pixel(p, load_pt(p, b, v), -1, -1, lo, hi, near_cnt);
            if (p.want_pix) pix[idx] = px;
            unsigned long long ol, oh;
            const unsigned long long rk = range_key<YAW>(p, b, v, fmask);
            cas128(cell + px, CELL_EMPTY, CELL_EMPTY, (unsigned long long)idx, rk, ol, oh);
            depth_test_settle(p, cell + px, rk, (unsigned long long)idx, ol, oh);
        }
    }
    missing = __reduce_add_sync(0xffffffffu, missing);
    near_cnt = __reduce_add_sync(0xffffffffu, near_cnt);
    if ((threadIdx.x & 31) == 0) {
        if (missing) atomicAdd(&p.diag[2 * b], missing);
        if (near_cnt) atomicAdd(&p.diag[2 * b + 1], near_cnt);
    }
    __syncthreads();
    dbg_stamp(p, 1, 4);
}

// planes: 0 x, 1 y, 2 z, 3 range (fp32 norm of the fp32 xyz, dataloader_semantic_KITTI.py:83),
//         4 intensity, 5 label (train id as float, as the reference's float32 image carries it)
template <bool CELLS>
__global__ void __launch_bounds__(PT_THREADS) proj_resolve_planes_kernel(const __grid_constant__ ProjParams p) {
    const int b = blockIdx.y;
    const long long n0 = p.offsets[b];
    const long long stride = (long long)gridDim.x * blockDim.x;
    int* winner = p.winner + (long long)b * p.HW;
    const ulonglong2* cell = reinterpret_cast<const ulonglong2*>(p.cell + (long long)b * p.HW);
    if (CELLS) dbg_stamp(p, 2, 0);
    // the gather is a chain of dependent loads (winner -> point -> raw label -> LUT): PT_BATCH pixels per thread go
    // through it together, one L2 round trip per link for the whole batch
    for (long long pb = (long long)blockIdx.x * blockDim.x + threadIdx.x; pb < p.HW; pb += stride * PT_BATCH) {
        int w[PT_BATCH];
        float4 v[PT_BATCH];
        unsigned raw[PT_BATCH];
        float lab[PT_BATCH];
#pragma unroll
        for (int k = 0; k < PT_BATCH; ++k) {
            if (CELLS) {
                ulonglong2 c = make_ulonglong2(CELL_EMPTY, CELL_EMPTY);
                if (pb + k * stride < p.HW) c = cell[pb + k * stride];
                w[k] = c.y == CELL_EMPTY ? 0x7fffffff : (int)c.x;
            } else {
                w[k] = (pb + k * stride < p.HW) ? winner[pb + k * stride] : 0x7fffffff;
            }
        }
#pragma unroll
        for (int k = 0; k < PT_BATCH; ++k) {
            const bool has = w[k] != 0x7fffffff;
            v[k] = has ? __ldg(p.xyzi + n0 + w[k]) : make_float4(0.f, 0.f, 0.f, 0.f);
            raw[k] = (has && p.raw_label) ? (__ldg(p.raw_label + n0 + w[k]) & 0xffffu) : 0u;
        }
#pragma unroll
        for (int k = 0; k < PT_BATCH; ++k) {
            const bool has = w[k] != 0x7fffffff && p.raw_label;
            lab[k] = has ? (float)(p.lut ? __ldg(p.lut + raw[k]) : (int)raw[k]) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < PT_BATCH; ++k) {
            const long long px = pb + k * stride;
            if (px >= p.HW) continue;
            float rng = 0.f;
            float4 q = v[k];
            if (w[k] != 0x7fffffff) {
                if (p.yaw) {                               // the image carries float32(rotated float64 coordinate)
                    const Pt t = load_pt(p, b, q);
                    q.x = (float)t.x; q.y = (float)t.y;
                }
                rng = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(q.x, q.x), __fmul_rn(q.y, q.y)), __fmul_rn(q.z, q.z)));
            }
            const long long cell = (long long)b * p.HW + px;
            if (!CELLS || p.want_winner) winner[px] = w[k] == 0x7fffffff ? -1 : w[k];
            if (p.label_img) p.label_img[cell] = (long long)lab[k];
            if (p.img) {
                float* o = p.img + (long long)b * 6 * p.HW + px;
                o[0] = q.x; o[p.HW] = q.y; o[2 * p.HW] = q.z; o[3 * p.HW] = rng; o[4 * p.HW] = q.w; o[5 * p.HW] = lab[k];
            }
        }
    }
    if (CELLS) dbg_stamp(p, 2, 1);
}

// generic form: [H,W,Cin] float32 = float(pc[winner]) per channel (utils.py:341-344)
__global__ void __launch_bounds__(PT_THREADS) proj_resolve_hwc_kernel(const __grid_constant__ ProjParams p) {
    const long long total = p.HW * p.cin;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long px = i / p.cin;
        const int c = (int)(i - px * p.cin);
        const int w = p.winner[px];
        if (p.img) p.img[i] = (w == 0x7fffffff || w < 0) ? 0.f : (float)p.pc[(long long)w * p.cin + c];
    }
}
__global__ void __launch_bounds__(PT_THREADS) proj_fix_winner_kernel(const __grid_constant__ ProjParams p) {
    for (long long px = (long long)blockIdx.x * blockDim.x + threadIdx.x; px < p.HW; px += (long long)gridDim.x * blockDim.x)
        if (p.winner[px] == 0x7fffffff) p.winner[px] = -1;
}

__global__ void __launch_bounds__(PT_THREADS) backproject_kernel(const long long* __restrict__ label_img,
                                                                 const int* __restrict__ pix,
                                                                 const __grid_constant__ ProjParams p,
                                                                 long long* __restrict__ out) {
    const int b = blockIdx.y;
    const long long n0 = p.offsets[b], n1 = p.offsets[b + 1];
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long* img = label_img + (long long)b * p.HW;
    for (long long nb = n0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; nb < n1; nb += stride * PT_BATCH) {
        int px[PT_BATCH];
        long long v[PT_BATCH];
#pragma unroll
        for (int k = 0; k < PT_BATCH; ++k) px[k] = (nb + k * stride < n1) ? __ldg(pix + nb + k * stride) : 0;
#pragma unroll
        for (int k = 0; k < PT_BATCH; ++k) v[k] = (nb + k * stride < n1) ? __ldg(img + px[k]) : 0ll;
#pragma unroll
        for (int k = 0; k < PT_BATCH; ++k)
            if (nb + k * stride < n1) out[nb + k * stride] = v[k];
    }
}

// ---- host ------------------------------------------------------------------------------------------
static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

constexpr int MAX_GX = 2048;         // upper bound of blocks per scan (partials are sized for it)
struct Workspace {
    int64_t theta, rkey, col, key, tminmax, pix, winner, diag, theta32, part_min, part_max, part_diag, part_exact, cell, total;
};
static Workspace carve(int64_t n_total, int B, int64_t HW) {
    Workspace w;
    int64_t o = 0;
    w.theta = o;   o = align_up(o + n_total * 8, 256);
    w.rkey = o;    o = align_up(o + n_total * 8, 256);
    w.key = o;     o = align_up(o + (int64_t)B * HW * 8, 256);
    w.tminmax = o; o = align_up(o + (int64_t)B * 16, 256);
    w.col = o;     o = align_up(o + n_total * 4, 256);
    w.pix = o;     o = align_up(o + n_total * 4, 256);
    w.winner = o;  o = align_up(o + (int64_t)B * HW * 4, 256);
    w.diag = o;    o = align_up(o + (int64_t)B * 8, 256);
    w.theta32 = o; o = align_up(o + n_total * 4, 256);
    // blocks per scan never exceed ceil(max(points of a scan, pixels) / PT_THREADS) (the extremes kernel also resets the cells)
    const int64_t items_max = n_total > HW ? n_total : HW;
    const int64_t gx_max = items_max / PT_THREADS + 1 < MAX_GX ? items_max / PT_THREADS + 1 : MAX_GX;
    w.part_min = o;  o = align_up(o + (int64_t)B * gx_max * 4, 256);
    w.part_max = o;  o = align_up(o + (int64_t)B * gx_max * 4, 256);
    w.part_diag = o; o = align_up(o + (int64_t)B * gx_max * 8, 256);
    w.part_exact = o; o = align_up(o + (int64_t)B * gx_max * 16, 256);
    w.cell = o;    o = align_up(o + (int64_t)B * HW * 16, 256);
    w.total = o;
    return w;
}

// CTAs per SM the point kernels are sized for (8 = one resident wave of 256-thread CTAs; more = several waves of shorter
// CTAs).  SLU_PT_CTAS_PER_SM overrides it for experiments.
static int g_pt_ctas_per_sm = [] { const char* e = getenv("SLU_PT_CTAS_PER_SM"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 8; }();

// The point / pixel kernels are latency-bound chains of dependent loads, atomics and barriers; they run best as EXACTLY
// one resident wave (measured: 4 or 8 CTAs/SM = 1 or 2 full waves of the angle kernel 0.080 / 0.084 ms, 3, 5 or 6 = a
// partial last wave 0.090-0.094 ms per 16 HDL-64 scans).  Each kernel's grid is therefore sized from ITS occupancy.
template <typename K>
static int resident_ctas_per_sm(K kernel) {
    // cached per kernel ADDRESS (several kernels share one function type, so a per-type static would hand every one of
    // them the occupancy of whichever was asked first)
    static const void* seen[32];
    static int value[32];
    static int n_seen = 0;
    const char* e = getenv("SLU_PT_CTAS_PER_SM");
    const int forced = e ? atoi(e) : 0;
    if (forced > 0) return forced;
    const void* key = reinterpret_cast<const void*>(kernel);
    for (int i = 0; i < n_seen; ++i)
        if (seen[i] == key) return value[i];
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, PT_THREADS, 0) != cudaSuccess || n < 1) { cudaGetLastError(); n = 4; }
    if (n_seen < 32) { seen[n_seen] = key; value[n_seen] = n; ++n_seen; }
    return n;
}

// waves of the kernels WITHOUT per-CTA set-up (ties, resolve, back-projection): short CTAs that the hardware scheduler deals
// out as SMs free up, which evens out the 30-40 % spread in per-SM speed that a single static wave leaves as its tail
static int g_light_waves = [] { const char* e = getenv("SLU_PT_LIGHT_WAVES"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 1; }();

template <typename K>
static int wave_grid_x(K kernel, long long items_max, int B, int sms, int waves = 1) {
    long long gx = (items_max + PT_THREADS - 1) / PT_THREADS;
    long long cap = ((long long)resident_ctas_per_sm(kernel) * sms * waves) / B;      // floor: never more than `waves` waves
    if (cap < 1) cap = 1;
    if (gx > cap) gx = cap;
    if (gx > MAX_GX) gx = MAX_GX;
    return (int)(gx < 1 ? 1 : gx);
}

static long long max_points(const long long* offsets, int B) {
    long long m = 1;
    for (int b = 0; b < B; ++b) m = offsets[b + 1] - offsets[b] > m ? offsets[b + 1] - offsets[b] : m;
    return m;
}

static int point_grid_x(const long long* offsets, int B, int sms) {
    long long max_n = 1;
    for (int b = 0; b < B; ++b) max_n = offsets[b + 1] - offsets[b] > max_n ? offsets[b + 1] - offsets[b] : max_n;
    long long gx = (max_n + PT_THREADS - 1) / PT_THREADS;
    const long long cap = ((long long)g_pt_ctas_per_sm * sms + B - 1) / B;   // CTAs per SM over the batch
    if (gx > cap) gx = cap;
    if (gx > MAX_GX) gx = MAX_GX;
    return (int)(gx < 1 ? 1 : gx);
}

// A/B switch for tests and profiles/: run the exact fp64 kernels in the batched entry point too.
// Initial value from SLU_PROJECT_EXACT=1, changed at run time by slu_debug_project_exact().
static int g_exact_only = [] { const char* e = getenv("SLU_PROJECT_EXACT"); return (e && e[0] >= '1' && e[0] <= '3') ? e[0] - '0' : 0; }();
static bool exact_only() { return g_exact_only == 1; }
static bool cell_path() { return g_exact_only == 2; }             // A/B: the three-launch cell pipeline (128-bit compare-and-swap depth test)
static bool cells_in_rows_path() { return g_exact_only == 3; }    // A/B: the default kernels with the depth test as a 128-bit compare-and-swap (no tie pass)
static int g_no_fused = [] { const char* e = getenv("SLU_PROJECT_NO_FUSED"); return (e && e[0] == '1') ? 1 : 0; }();   // A/B: two-pass kernels with a fixed range too

static int project_common(ProjParams& p, int64_t n_total, void* d_work, int32_t* d_pix, int32_t* d_winner,
                          int32_t* d_diag, bool generic, cudaStream_t st) {
    const int sms = sm_count_current_device();
    if (sms <= 0) return fail(SLU_E_DEVICE, "no CUDA device");
    if ((reinterpret_cast<uintptr_t>(d_work) & 15) != 0) return fail(SLU_E_ALIGN, "d_work not 16-byte aligned");
    const Workspace w = carve(n_total, p.B, p.HW);
    char* base = static_cast<char*>(d_work);
    p.theta = reinterpret_cast<double*>(base + w.theta);
    p.rkey = reinterpret_cast<unsigned long long*>(base + w.rkey);
    p.col = reinterpret_cast<int*>(base + w.col);
    p.key = reinterpret_cast<unsigned long long*>(base + w.key);
    p.tminmax = reinterpret_cast<unsigned long long*>(base + w.tminmax);
    p.pix = d_pix ? d_pix : reinterpret_cast<int*>(base + w.pix);
    p.winner = d_winner ? d_winner : reinterpret_cast<int*>(base + w.winner);
    p.diag = d_diag ? d_diag : reinterpret_cast<int*>(base + w.diag);
    p.theta32 = reinterpret_cast<float*>(base + w.theta32);
    p.part_min = reinterpret_cast<float*>(base + w.part_min);
    p.part_max = reinterpret_cast<float*>(base + w.part_max);
    p.part_diag = reinterpret_cast<int*>(base + w.part_diag);
    p.part_exact = reinterpret_cast<double*>(base + w.part_exact);
    p.cell = reinterpret_cast<Cell*>(base + w.cell);
    p.want_pix = d_pix != nullptr; p.want_winner = d_winner != nullptr;
    const dim3 gp(point_grid_x(p.offsets, p.B, sms), p.B);
    p.gx = (int)gp.x;
    if (!generic && cell_path()) {
        // cell pipeline: extremes (+ cell reset) -> fused point pass with the 128-bit depth test; the caller resolves
        p.key_sq = 1;
        p.use_cells = 1;
        const long long nmax = max_points(p.offsets, p.B);
        long long items = p.use_range ? p.HW : (nmax > p.HW ? nmax : p.HW);
        static const int dbg_times = [] { const char* e = getenv("SLU_P3_TIMES"); return e ? atoi(e) : 0; }();
        p.dbg_times = dbg_times ? reinterpret_cast<unsigned long long*>(base + w.theta) : nullptr;   // the theta region is unused on this path
        p.gx = wave_grid_x(proj3_extremes_kernel, items, p.B, sms);          // P2 reads gx per-block partials
        proj3_extremes_kernel<<<dim3(p.gx, p.B), PT_THREADS, 0, st>>>(p);
        SLU_LAUNCH_CHECK("proj3_extremes_kernel");
        static const int variant = [] { const char* e = getenv("SLU_P3_VARIANT"); return e ? atoi(e) : 0; }();
        static const int dbg_atom = [] { const char* e = getenv("SLU_P3_ATOM"); return e ? atoi(e) : 0; }();
        p.dbg_atom = dbg_atom;
#define SLU_P3_LAUNCH2(NB, MINB, YAW, IDS)                                                                                               \
    proj3_points_kernel<NB, MINB, YAW, IDS><<<dim3(wave_grid_x(proj3_points_kernel<NB, MINB, YAW, IDS>, (nmax + NB - 1) / NB, p.B, sms), p.B), PT_THREADS, 0, st>>>(p)
#define SLU_P3_LAUNCH(NB, MINB)                                                        \
    do {                                                                               \
        if (p.yaw) { if (ids) SLU_P3_LAUNCH2(NB, MINB, true, true); else SLU_P3_LAUNCH2(NB, MINB, true, false); }      \
        else { if (ids) SLU_P3_LAUNCH2(NB, MINB, false, true); else SLU_P3_LAUNCH2(NB, MINB, false, false); }          \
    } while (0)
        const bool ids = p.raw_label != nullptr && p.lut != nullptr;
        switch (variant) {
            case 1: SLU_P3_LAUNCH(1, 4); break;
            case 2: SLU_P3_LAUNCH(1, 6); break;
            default: SLU_P3_LAUNCH(1, 5); break;
        }
#undef SLU_P3_LAUNCH2
#undef SLU_P3_LAUNCH
        SLU_LAUNCH_CHECK("proj3_points_kernel");
        return 0;
    }
    if (!generic && !exact_only()) {
        p.key_sq = 1;
        // default: fire-and-forget RED.MIN.64 on the range key, then the tie pass.  Measured alternative (mode 3): a 128-bit
        // compare-and-swap on a (range, index) cell settles the winner in the row kernel and deletes the tie pass (11 us per 16
        // HDL-64 scans) -- but the row kernel is an atomic stream, and atomics that RETURN a value cost 36.0 us there against
        // 15.2 us for the reductions: 0.090 ms in total against 0.072.
        p.use_cells = cells_in_rows_path() ? 1 : 0;
        p.use_cross = g_cross_mode == 1 || (g_cross_mode == 2 && p.use_range && !g_no_fused) ? 1 : 0;
        if (p.use_range && !g_no_fused) {
            // fixed elevation range: init -> one fused pass (angles + rows + depth test) -> ties
            const long long cells = (long long)p.B * p.HW;
            const long long gi = (cells + PT_THREADS - 1) / PT_THREADS;
            proj_init_kernel<<<(unsigned)(gi < 8LL * sms ? (gi < 1 ? 1 : gi) : 8LL * sms), PT_THREADS, 0, st>>>(p);
            SLU_LAUNCH_CHECK("proj_init_kernel");
            if (p.use_cells) proj_fast_fused_kernel<true><<<dim3(wave_grid_x(proj_fast_fused_kernel<true>, max_points(p.offsets, p.B), p.B, sms), p.B), PT_THREADS, 0, st>>>(p);
            else proj_fast_fused_kernel<false><<<dim3(wave_grid_x(proj_fast_fused_kernel<false>, max_points(p.offsets, p.B), p.B, sms), p.B), PT_THREADS, 0, st>>>(p);
            SLU_LAUNCH_CHECK("proj_fast_fused_kernel");
        } else {
            // fp32-prefiltered path: angles (+ init, + exact theta extremes per block) -> rows -> ties
            const long long nmax = max_points(p.offsets, p.B);
            p.gx = wave_grid_x(proj_fast_angles_kernel, nmax, p.B, sms);      // the row kernel reads gx per-block partials
            proj_fast_angles_kernel<<<dim3(p.gx, p.B), PT_THREADS, 0, st>>>(p);
            SLU_LAUNCH_CHECK("proj_fast_angles_kernel");
            if (p.use_cells) proj_fast_rows_kernel<true><<<dim3(wave_grid_x(proj_fast_rows_kernel<true>, nmax, p.B, sms), p.B), PT_THREADS, 0, st>>>(p);
            else proj_fast_rows_kernel<false><<<dim3(wave_grid_x(proj_fast_rows_kernel<false>, nmax, p.B, sms), p.B), PT_THREADS, 0, st>>>(p);
            SLU_LAUNCH_CHECK("proj_fast_rows_kernel");
        }
        if (n_total > 0 && !p.use_cells) {
            proj_ties_kernel<<<dim3(wave_grid_x(proj_ties_kernel, (max_points(p.offsets, p.B) + PT_BATCH - 1) / PT_BATCH, p.B, sms, g_light_waves), p.B), PT_THREADS, 0, st>>>(p);
            SLU_LAUNCH_CHECK("proj_ties_kernel");
        }
        return 0;
    }

    const long long cells = (long long)p.B * p.HW;
    const int gi = (int)((cells + PT_THREADS - 1) / PT_THREADS < 8LL * sms ? (cells + PT_THREADS - 1) / PT_THREADS : 8LL * sms);
    proj_init_kernel<<<gi < 1 ? 1 : gi, PT_THREADS, 0, st>>>(p);
    SLU_LAUNCH_CHECK("proj_init_kernel");
    if (n_total > 0) {
        if (generic) proj_angles_kernel<true><<<gp, PT_THREADS, 0, st>>>(p);
        else proj_angles_kernel<false><<<gp, PT_THREADS, 0, st>>>(p);
        SLU_LAUNCH_CHECK("proj_angles_kernel");
    }
    proj_rows_kernel<<<gp, PT_THREADS, 0, st>>>(p);           // also writes theta_out for empty scans
    SLU_LAUNCH_CHECK("proj_rows_kernel");
    if (n_total > 0) {
        proj_ties_kernel<<<gp, PT_THREADS, 0, st>>>(p);
        SLU_LAUNCH_CHECK("proj_ties_kernel");
    }
    return 0;
}

}  // namespace slu

extern "C" int slu_debug_project_exact(int on) {
    const int prev = slu::g_exact_only;
    if (on >= 0) slu::g_exact_only = on > 3 ? 1 : on;
    return prev;
}

extern "C" int64_t slu_project_workspace_bytes(int64_t n_total, int B, int64_t HW) {
    if (n_total < 0 || B < 1 || HW < 1) return slu::fail(SLU_E_ARG, "bad workspace query");
    return slu::carve(n_total, B, HW).total;
}

extern "C" int slu_project_batch(const float* d_xyzi, const uint32_t* d_raw_label, const int32_t* d_lut,
                                 const int64_t* h_offsets, int64_t n_total, int B, int H, int W,
                                 int use_theta_range, double theta_lo, double theta_hi, int farthest_wins,
                                 const double* d_yaw_cs, void* d_work,
                                 float* d_img, int64_t* d_label, int32_t* d_pix, int32_t* d_winner, double* d_theta,
                                 int32_t* d_diag, slu_stream_t stream) {
    using namespace slu;
    if (B < 1 || B > MAX_SCANS) return fail(SLU_E_RANGE, "B=%d outside [1,%d]", B, MAX_SCANS);
    if (H < 1 || W < 1 || (long long)H * W > 0x7fffffffLL) return fail(SLU_E_RANGE, "image %dx%d unsupported", H, W);
    if (!h_offsets || !d_work) return fail(SLU_E_ARG, "h_offsets / d_work is NULL");
    if (n_total < 0 || h_offsets[0] != 0 || h_offsets[B] != n_total) return fail(SLU_E_ARG, "offsets do not span [0,n_total]");
    for (int b = 0; b < B; ++b) {
        if (h_offsets[b + 1] < h_offsets[b]) return fail(SLU_E_ARG, "offsets must be non-decreasing");
        if (h_offsets[b + 1] - h_offsets[b] > 0x7ffffff0LL) return fail(SLU_E_RANGE, "scan %d has too many points", b);
    }
    if (n_total > 0 && !d_xyzi) return fail(SLU_E_ARG, "d_xyzi is NULL");
    if ((reinterpret_cast<uintptr_t>(d_xyzi) & 15) != 0) return fail(SLU_E_ALIGN, "d_xyzi not 16-byte aligned");
    if (use_theta_range && !(theta_lo == theta_lo && theta_hi == theta_hi)) return fail(SLU_E_ARG, "theta range is NaN");
    ProjParams p{};
    p.xyzi = reinterpret_cast<const float4*>(d_xyzi);
    p.raw_label = d_raw_label;
    p.lut = d_lut;
    p.yaw = d_yaw_cs;
    for (int b = 0; b <= B; ++b) p.offsets[b] = h_offsets[b];
    p.B = B; p.H = H; p.W = W; p.HW = (long long)H * W;
    p.use_range = use_theta_range; p.theta_lo = theta_lo; p.theta_hi = theta_hi;
    p.farthest = farthest_wins;
    p.img = d_img; p.theta_out = d_theta;
    p.label_img = reinterpret_cast<long long*>(d_label);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc = project_common(p, n_total, d_work, d_pix, d_winner, d_diag, false, st);
    if (rc) return rc;
    const int sms = sm_count_current_device();
    if (p.use_cells) {
        const int gx = wave_grid_x(proj_resolve_planes_kernel<true>, (p.HW + PT_BATCH - 1) / PT_BATCH, B, sms, g_light_waves);
        proj_resolve_planes_kernel<true><<<dim3((unsigned)gx, B), PT_THREADS, 0, st>>>(p);
    } else {
        const int gx = wave_grid_x(proj_resolve_planes_kernel<false>, (p.HW + PT_BATCH - 1) / PT_BATCH, B, sms, g_light_waves);
        proj_resolve_planes_kernel<false><<<dim3((unsigned)gx, B), PT_THREADS, 0, st>>>(p);
    }
    SLU_LAUNCH_CHECK("proj_resolve_planes_kernel");
    return 0;
}

extern "C" int slu_project_points_bins(const double* d_pc, int64_t N, int Cin, int H, int W,
                                       int use_theta_range, double theta_lo, double theta_hi, int farthest_wins,
                                       const double* d_row_edges_ascending, int edges_were_increasing, void* d_work,
                                       float* d_img_hwc, int32_t* d_pix, int32_t* d_winner, double* d_theta, int32_t* d_diag,
                                       slu_stream_t stream);

extern "C" int slu_project_points(const double* d_pc, int64_t N, int Cin, int H, int W,
                                  int use_theta_range, double theta_lo, double theta_hi, int farthest_wins,
                                  void* d_work,
                                  float* d_img_hwc, int32_t* d_pix, int32_t* d_winner, double* d_theta, int32_t* d_diag,
                                  slu_stream_t stream) {
    return slu_project_points_bins(d_pc, N, Cin, H, W, use_theta_range, theta_lo, theta_hi, farthest_wins, nullptr, 0, d_work,
                                   d_img_hwc, d_pix, d_winner, d_theta, d_diag, stream);
}

extern "C" int slu_project_points_bins(const double* d_pc, int64_t N, int Cin, int H, int W,
                                       int use_theta_range, double theta_lo, double theta_hi, int farthest_wins,
                                       const double* d_row_edges_ascending, int edges_were_increasing, void* d_work,
                                       float* d_img_hwc, int32_t* d_pix, int32_t* d_winner, double* d_theta, int32_t* d_diag,
                                       slu_stream_t stream) {
    using namespace slu;
    if (N < 0 || N > 0x7ffffff0LL) return fail(SLU_E_RANGE, "N=%lld unsupported", (long long)N);
    if (Cin < 3) return fail(SLU_E_ARG, "Cin=%d < 3", Cin);
    if (H < 1 || W < 1 || (long long)H * W > 0x7fffffffLL) return fail(SLU_E_RANGE, "image %dx%d unsupported", H, W);
    if (!d_work) return fail(SLU_E_ARG, "d_work is NULL");
    if (N > 0 && !d_pc) return fail(SLU_E_ARG, "d_pc is NULL");
    if (use_theta_range && !(theta_lo == theta_lo && theta_hi == theta_hi)) return fail(SLU_E_ARG, "theta range is NaN");
    ProjParams p{};
    p.pc = d_pc; p.cin = Cin;
    p.offsets[0] = 0; p.offsets[1] = N;
    p.B = 1; p.H = H; p.W = W; p.HW = (long long)H * W;
    p.use_range = use_theta_range; p.theta_lo = theta_lo; p.theta_hi = theta_hi;
    p.farthest = farthest_wins;
    p.row_edges = d_row_edges_ascending; p.row_edges_increasing = edges_were_increasing;
    p.img = d_img_hwc; p.theta_out = d_theta;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc = project_common(p, N, d_work, d_pix, d_winner, d_diag, true, st);
    if (rc) return rc;
    const int sms = sm_count_current_device();
    if (d_img_hwc) {
        long long gx = (p.HW * Cin + PT_THREADS - 1) / PT_THREADS;
        if (gx > 8LL * sms) gx = 8LL * sms;
        proj_resolve_hwc_kernel<<<(unsigned)gx, PT_THREADS, 0, st>>>(p);
        SLU_LAUNCH_CHECK("proj_resolve_hwc_kernel");
    }
    long long gx = (p.HW + PT_THREADS - 1) / PT_THREADS;
    if (gx > 8LL * sms) gx = 8LL * sms;
    proj_fix_winner_kernel<<<(unsigned)gx, PT_THREADS, 0, st>>>(p);
    SLU_LAUNCH_CHECK("proj_fix_winner_kernel");
    return 0;
}

extern "C" int slu_backproject(const int64_t* d_label_img, const int32_t* d_pix, const int64_t* h_offsets,
                               int64_t n_total, int B, int64_t HW, int64_t* d_out, slu_stream_t stream) {
    using namespace slu;
    if (B < 1 || B > MAX_SCANS) return fail(SLU_E_RANGE, "B=%d outside [1,%d]", B, MAX_SCANS);
    if (!h_offsets || h_offsets[0] != 0 || h_offsets[B] != n_total || n_total < 0) return fail(SLU_E_ARG, "offsets do not span [0,n_total]");
    if (n_total == 0) return 0;
    if (!d_label_img || !d_pix || !d_out) return fail(SLU_E_ARG, "NULL pointer");
    if (HW < 1) return fail(SLU_E_ARG, "HW < 1");
    ProjParams p{};
    for (int b = 0; b <= B; ++b) p.offsets[b] = h_offsets[b];
    p.B = B; p.HW = HW;
    const int sms = sm_count_current_device();
    if (sms <= 0) return fail(SLU_E_DEVICE, "no CUDA device");
    const dim3 g(wave_grid_x(backproject_kernel, (max_points(p.offsets, B) + PT_BATCH - 1) / PT_BATCH, B, sms, g_light_waves), B);
    backproject_kernel<<<g, PT_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const long long*>(d_label_img), d_pix, p, reinterpret_cast<long long*>(d_out));
    SLU_LAUNCH_CHECK("backproject_kernel");
    return 0;
}

/* diagnostic: the prefilter's fp32 arctangent on arrays (tests measure its error against float64 atan2) */
namespace slu {
__global__ void fast_atan2_kernel(const float* __restrict__ y, const float* __restrict__ x, long long n, float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = fast_atan2(y[i], x[i]);
}
}  // namespace slu
extern "C" int slu_diag_fast_atan2(const float* d_y, const float* d_x, int64_t n, float* d_out, slu_stream_t stream) {
    using namespace slu;
    if (!d_y || !d_x || !d_out || n < 1) return fail(SLU_E_ARG, "slu_diag_fast_atan2: bad arguments");
    fast_atan2_kernel<<<1024, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_y, d_x, n, d_out);
    SLU_LAUNCH_CHECK("fast_atan2_kernel");
    return 0;
}
