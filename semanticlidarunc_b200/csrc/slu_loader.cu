// slu_loader.cu -- loader glue behind the projection, on the device.
//
// Replaces (reference file:line): src/dataset/dataloader_semantic_KITTI.py:60-97 --
// cv2.resize(..., INTER_NEAREST) :61-62, horizontal flip + y negation :71-73, channel split :75-78,
// range = ||xyz|| :83, build_normal_xyz :85 (src/dataset/utils.py:30-59: six 3x3 Scharr filters, cross
// product, normalisation) and the tensor packing :91-97.
//
//   frame_resample_kernel  one thread per OUTPUT pixel: nearest-neighbour source index exactly as
//                          OpenCV computes it (sx = min(floor(x * (1 / (dst/src))), src-1)), optional
//                          column flip with y -> -y; writes xyz, range, reflectivity, semantics.
//   frame_normals_kernel   one thread per output pixel: Scharr d/dx and d/dy of x, y, z with
//                          BORDER_REFLECT_101, in OpenCV's separable order (row pass, then column
//                          pass; derivative taps -1,0,1; smoothing taps 3,10,3 scaled by 1/norm_factor),
//                          n = -(dy x dz style cross product), n /= (||n|| + 1e-10).
// Pure copies are bit-exact; the normals agree with cv2 to float32 rounding (cv2's SIMD paths may fuse
// multiply-adds differently), tested at 1e-5.
#include <stdlib.h>
#include "slu_common.cuh"

namespace slu {

constexpr int FR_THREADS = 256;

struct FrameParams {
    const float* img;        // [B,6,Hs*Ws] x,y,z,range,intensity,label
    int B, Hs, Ws, Hd, Wd;
    double ify, ifx;         // OpenCV's inverse scale factors
    unsigned long long flip_bits[4];   // bit b = flip scan b (B <= 256)
    float tap_c, tap_s;      // Scharr smoothing taps (10, 3) * scale
    const int* rowmap;       // [B, Hs+1]: count, then the source rows kept (drop_empty_rows) or NULL
    int keep_native_h;       // no resize requested: output row r <- kept row r, rows beyond the count are zero
    float* range;            // [B,1,Hd*Wd]
    float* refl;             // [B,1,Hd*Wd]
    float* xyz;              // [B,3,Hd*Wd]
    long long xyz_bstride;   // elements between the xyz planes of consecutive scans (3*Hd*Wd; 6*HW when reading the image)
    float* normals;          // [B,3,Hd*Wd]
    long long* sem;          // [B,1,Hd*Wd]
};

__global__ void __launch_bounds__(FR_THREADS) frame_resample_kernel(const __grid_constant__ FrameParams p) {
    const int b = blockIdx.y;
    const long long HWs = (long long)p.Hs * p.Ws, HWd = (long long)p.Hd * p.Wd;
    const bool flip = (p.flip_bits[b >> 6] >> (b & 63)) & 1ull;
    const float* src = p.img + (long long)b * 6 * HWs;
    // WADS drops image rows without any return before resizing (dataloader_semantic_WADS.py:125): the
    // source height is then the per-scan count of kept rows and source rows go through the row map
    const int* rmap = p.rowmap ? p.rowmap + (long long)b * (p.Hs + 1) : nullptr;
    const int hs_eff = rmap ? rmap[0] : p.Hs;
    const int hd_eff = p.keep_native_h ? hs_eff : p.Hd;
    const double ify = rmap ? (hs_eff > 0 ? 1.0 / ((double)hd_eff / (double)hs_eff) : 0.0) : p.ify;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HWd; i += (long long)gridDim.x * blockDim.x) {
        const int yd = (int)(i / p.Wd);
        int xd = (int)(i - (long long)yd * p.Wd);
        if (flip) xd = p.Wd - 1 - xd;                                   // resized[:, ::-1]
        int ys = (int)floor((double)yd * ify), xs = (int)floor((double)xd * p.ifx);
        ys = ys < hs_eff - 1 ? ys : hs_eff - 1;
        xs = xs < p.Ws - 1 ? xs : p.Ws - 1;
        const bool empty = hs_eff <= 0 || yd >= hd_eff;
        if (rmap && !empty) ys = rmap[1 + ys];
        const long long s = empty ? 0 : (long long)ys * p.Ws + xs;
        const float x = empty ? 0.f : src[s], z = empty ? 0.f : src[2 * HWs + s];
        float y = empty ? 0.f : src[HWs + s];
        if (flip) y = -y;                                               // :73
        const long long o = (long long)b * HWd + i;
        if (p.xyz) {
            float* q = p.xyz + (long long)b * 3 * HWd + i;
            q[0] = x; q[HWd] = y; q[2 * HWd] = z;
        }
        if (p.range) p.range[o] = empty ? 0.f : src[3 * HWs + s];
        if (p.refl) p.refl[o] = empty ? 0.f : src[4 * HWs + s];
        if (p.sem) p.sem[o] = empty ? 0ll : (long long)src[5 * HWs + s];
    }
}

// rows with at least one non-zero x, y, z, intensity or label value (the reference tests the 5-channel norm)
__global__ void __launch_bounds__(FR_THREADS) frame_rowmap_kernel(const float* __restrict__ img, int Hs, int Ws, int* __restrict__ rowmap) {
    extern __shared__ int s_flag[];
    const int b = blockIdx.x;
    const long long HWs = (long long)Hs * Ws;
    const float* src = img + (long long)b * 6 * HWs;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    for (int r = warp; r < Hs; r += nwarp) {
        bool any = false;
        for (int x = lane; x < Ws && !any; x += 32) {
            const long long s = (long long)r * Ws + x;
            any = src[s] != 0.f || src[HWs + s] != 0.f || src[2 * HWs + s] != 0.f || src[4 * HWs + s] != 0.f || src[5 * HWs + s] != 0.f;
        }
        any = __any_sync(0xffffffffu, any);
        if (lane == 0) s_flag[r] = any ? 1 : 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int* m = rowmap + (long long)b * (Hs + 1);
        int n = 0;
        for (int r = 0; r < Hs; ++r)
            if (s_flag[r]) m[1 + n++] = r;
        m[0] = n;
    }
}

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i < 0 ? 0 : (i >= n ? n - 1 : i);
}

__global__ void __launch_bounds__(FR_THREADS) frame_normals_kernel(const __grid_constant__ FrameParams p) {
    const int b = blockIdx.y;
    const long long HW = (long long)p.Hd * p.Wd;
    const float* base = p.xyz + (long long)b * p.xyz_bstride;
    float* out = p.normals + (long long)b * 3 * HW;
    // with dropped rows and no resize the image really has rowmap[0] rows: reflect at that border
    const int h_eff = (p.rowmap && p.keep_native_h) ? p.rowmap[(long long)b * (p.Hs + 1)] : p.Hd;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / p.Wd), x = (int)(i - (long long)y * p.Wd);
        if (y >= h_eff) { out[i] = 0.f; out[HW + i] = 0.f; out[2 * HW + i] = 0.f; continue; }
        const int ym = reflect101(y - 1, h_eff), yp = reflect101(y + 1, h_eff);
        const int xm = reflect101(x - 1, p.Wd), xp = reflect101(x + 1, p.Wd);
        float gx[3], gy[3];                      // d/dx and d/dy of the planes x, y, z
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* P = base + (long long)c * HW;
            const float a00 = P[(long long)ym * p.Wd + xm], a01 = P[(long long)ym * p.Wd + x], a02 = P[(long long)ym * p.Wd + xp];
            const float a10 = P[(long long)y * p.Wd + xm], a11 = P[(long long)y * p.Wd + x], a12 = P[(long long)y * p.Wd + xp];
            const float a20 = P[(long long)yp * p.Wd + xm], a21 = P[(long long)yp * p.Wd + x], a22 = P[(long long)yp * p.Wd + xp];
            // Scharr(dx=1): rows (a02-a00, a12-a10, a22-a20), then column taps (s, c, s)
            const float d0 = __fsub_rn(a02, a00), d1 = __fsub_rn(a12, a10), d2 = __fsub_rn(a22, a20);
            gx[c] = __fadd_rn(__fmul_rn(p.tap_c, d1), __fmul_rn(p.tap_s, __fadd_rn(d0, d2)));
            // Scharr(dy=1): rows smoothed with (s, c, s), then column taps (-1, 0, 1)
            const float r0 = __fadd_rn(__fmul_rn(p.tap_c, a01), __fmul_rn(p.tap_s, __fadd_rn(a00, a02)));
            const float r2 = __fadd_rn(__fmul_rn(p.tap_c, a21), __fmul_rn(p.tap_s, __fadd_rn(a20, a22)));
            gy[c] = __fsub_rn(r2, r0);
            (void)a11;
        }
        // utils.py:49-51: normal = -(Syx*Szy - Szx*Syy, Szx*Sxy - Szy*Sxx, Sxx*Syy - Syx*Sxy)
        const float Sxx = gx[0], Sxy = gy[0], Syx = gx[1], Syy = gy[1], Szx = gx[2], Szy = gy[2];
        float nx = -__fsub_rn(__fmul_rn(Syx, Szy), __fmul_rn(Szx, Syy));
        float ny = -__fsub_rn(__fmul_rn(Szx, Sxy), __fmul_rn(Szy, Sxx));
        float nz = -__fsub_rn(__fmul_rn(Sxx, Syy), __fmul_rn(Syx, Sxy));
        const float n = __fadd_rn(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz))), 1e-10f);
        out[i] = __fdiv_rn(nx, n);
        out[HW + i] = __fdiv_rn(ny, n);
        out[2 * HW + i] = __fdiv_rn(nz, n);
    }
}

// Four consecutive pixels of a row per thread: per plane three 16-byte row loads plus the two neighbours left and right
// of the strip (9 load instructions instead of 36 for four pixels), float4 stores.  Same arithmetic, operation for
// operation, as frame_normals_kernel (the results are bit-identical); used when W % 4 == 0, the planes are 16-byte
// aligned and no rows were dropped.
__global__ void __launch_bounds__(FR_THREADS) frame_normals4_kernel(const __grid_constant__ FrameParams p) {
    const int b = blockIdx.y;
    const long long HW = (long long)p.Hd * p.Wd;
    const float* base = p.xyz + (long long)b * p.xyz_bstride;
    float* out = p.normals + (long long)b * 3 * HW;
    const int W4 = p.Wd >> 2;
    const long long n4 = (long long)p.Hd * W4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / W4), x0 = (int)(i - (long long)y * W4) << 2;
        const int ym = reflect101(y - 1, p.Hd), yp = reflect101(y + 1, p.Hd);
        const int xm = reflect101(x0 - 1, p.Wd), xp = reflect101(x0 + 4, p.Wd);
        float gx[3][4], gy[3][4];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* P = base + (long long)c * HW;
            float r[3][6];                                   // rows ym, y, yp; columns x0-1 .. x0+4
            const int ys[3] = {ym, y, yp};
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const float* row = P + (long long)ys[j] * p.Wd;
                const float4 m = *reinterpret_cast<const float4*>(row + x0);
                r[j][0] = row[xm]; r[j][1] = m.x; r[j][2] = m.y; r[j][3] = m.z; r[j][4] = m.w; r[j][5] = row[xp];
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float a00 = r[0][k], a01 = r[0][k + 1], a02 = r[0][k + 2];
                const float a10 = r[1][k], a12 = r[1][k + 2];
                const float a20 = r[2][k], a21 = r[2][k + 1], a22 = r[2][k + 2];
                const float d0 = __fsub_rn(a02, a00), d1 = __fsub_rn(a12, a10), d2 = __fsub_rn(a22, a20);
                gx[c][k] = __fadd_rn(__fmul_rn(p.tap_c, d1), __fmul_rn(p.tap_s, __fadd_rn(d0, d2)));
                const float r0 = __fadd_rn(__fmul_rn(p.tap_c, a01), __fmul_rn(p.tap_s, __fadd_rn(a00, a02)));
                const float r2 = __fadd_rn(__fmul_rn(p.tap_c, a21), __fmul_rn(p.tap_s, __fadd_rn(a20, a22)));
                gy[c][k] = __fsub_rn(r2, r0);
            }
        }
        float o[3][4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float Sxx = gx[0][k], Sxy = gy[0][k], Syx = gx[1][k], Syy = gy[1][k], Szx = gx[2][k], Szy = gy[2][k];
            const float nx = -__fsub_rn(__fmul_rn(Syx, Szy), __fmul_rn(Szx, Syy));
            const float ny = -__fsub_rn(__fmul_rn(Szx, Sxy), __fmul_rn(Szy, Sxx));
            const float nz = -__fsub_rn(__fmul_rn(Sxx, Syy), __fmul_rn(Syx, Sxy));
            const float n = __fadd_rn(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz))), 1e-10f);
            o[0][k] = __fdiv_rn(nx, n); o[1][k] = __fdiv_rn(ny, n); o[2][k] = __fdiv_rn(nz, n);
        }
        const long long off = (long long)y * p.Wd + x0;
#pragma unroll
        for (int c = 0; c < 3; ++c)
            *reinterpret_cast<float4*>(out + (long long)c * HW + off) = make_float4(o[c][0], o[c][1], o[c][2], o[c][3]);
    }
}

static void launch_normals(const FrameParams& p, int B, int sms, cudaStream_t st) {
    const long long HWd = (long long)p.Hd * p.Wd;
    const bool vec = (p.Wd & 3) == 0 && p.Wd >= 8 && !(p.rowmap && p.keep_native_h) && (p.xyz_bstride & 3) == 0 &&
                     ((reinterpret_cast<uintptr_t>(p.xyz) | reinterpret_cast<uintptr_t>(p.normals)) & 15) == 0;
    const long long work = vec ? HWd / 4 : HWd;
    long long gx = (work + FR_THREADS - 1) / FR_THREADS;
    static const int ctas_per_sm = [] { const char* e = getenv("SLU_FR_CTAS_PER_SM"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 8; }();
    const long long cap = ((long long)ctas_per_sm * sms + B - 1) / B;
    if (gx > cap) gx = cap;
    if (vec) frame_normals4_kernel<<<dim3((unsigned)gx, B), FR_THREADS, 0, st>>>(p);
    else frame_normals_kernel<<<dim3((unsigned)gx, B), FR_THREADS, 0, st>>>(p);
}

// ---- organised clouds (Ouster / SemanticTHAB): the sensor already delivers H x W points, pixel = point -----------
// Replaces src/dataset/dataloader_semantic_THAB.py:35-66: label remap, reshape to (H,W,.), optional flip
// (columns reversed, y negated), optional yaw as an image roll (rotate_equirectangular_image,
// src/dataset/utils.py:21-28) plus rotate_z on the coordinates, range = ||xyz|| in float64 (the THAB image is
// float64, unlike the projected loaders), cast to float32.
struct OrgParams {
    const float4* xyzi;        // [B*HW]
    const unsigned* raw;       // [B*HW] or NULL
    const int* lut;            // [65536] or NULL
    int B, H, W;
    unsigned long long flip_bits[4];
    const int* shift;          // [B] column roll per scan (device) or NULL
    const double* yaw;         // [B,2] cos, sin (device) or NULL
    float* img;                // [B,6,HW]
    int* missing;              // [B] or NULL
};

__global__ void __launch_bounds__(FR_THREADS) organized_planes_kernel(const __grid_constant__ OrgParams p) {
    const int b = blockIdx.y;
    const long long HW = (long long)p.H * p.W;
    const bool flip = (p.flip_bits[b >> 6] >> (b & 63)) & 1ull;
    const int shift = p.shift ? p.shift[b] : 0;
    int miss = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / p.W), x = (int)(i - (long long)y * p.W);
        int xs = ((x - shift) % p.W + p.W) % p.W;           // np.roll(img, shift, axis=1): out[x] = in[x - shift]
        if (flip) xs = p.W - 1 - xs;                          // the flip is applied before the roll
        const long long src = (long long)b * HW + (long long)y * p.W + xs;
        const float4 v = __ldg(p.xyzi + src);
        double X = (double)v.x, Y = flip ? -(double)v.y : (double)v.y, Z = (double)v.z;
        if (p.yaw) {
            const double c = p.yaw[2 * b], s = p.yaw[2 * b + 1];
            const double xr = __dadd_rn(__dmul_rn(X, c), __dmul_rn(Y, s));
            const double yr = __dadd_rn(__dmul_rn(X, -s), __dmul_rn(Y, c));
            X = xr; Y = yr;
        }
        float lab = 0.f;
        if (p.raw) {
            const unsigned id = __ldg(p.raw + src) & 0xffffu;
            const int t = p.lut ? __ldg(p.lut + id) : (int)id;
            if (t < 0) ++miss;
            lab = (float)t;
        }
        const double r = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(X, X), __dmul_rn(Y, Y)), __dmul_rn(Z, Z)));
        float* o = p.img + (long long)b * 6 * HW + i;
        o[0] = (float)X; o[HW] = (float)Y; o[2 * HW] = (float)Z; o[3 * HW] = (float)r; o[4 * HW] = v.w; o[5 * HW] = lab;
    }
    miss = __reduce_add_sync(0xffffffffu, miss);
    if ((threadIdx.x & 31) == 0 && miss && p.missing) atomicAdd(&p.missing[b], miss);
}

}  // namespace slu

extern "C" int slu_organized_planes(const float* d_xyzi, const uint32_t* d_raw_label, const int32_t* d_lut,
                                    int B, int H, int W, const uint8_t* h_flip, const int32_t* d_col_shift,
                                    const double* d_yaw_cs, float* d_img, int32_t* d_missing, slu_stream_t stream) {
    using namespace slu;
    if (!d_xyzi || !d_img) return fail(SLU_E_ARG, "d_xyzi / d_img is NULL");
    if (B < 1 || B > 256 || H < 1 || W < 1) return fail(SLU_E_RANGE, "B=%d H=%d W=%d unsupported", B, H, W);
    if ((reinterpret_cast<uintptr_t>(d_xyzi) & 15) != 0) return fail(SLU_E_ALIGN, "d_xyzi not 16-byte aligned");
    OrgParams p{};
    p.xyzi = reinterpret_cast<const float4*>(d_xyzi); p.raw = d_raw_label; p.lut = d_lut;
    p.B = B; p.H = H; p.W = W;
    for (int b = 0; b < B && h_flip; ++b)
        if (h_flip[b]) p.flip_bits[b >> 6] |= 1ull << (b & 63);
    p.shift = d_col_shift; p.yaw = d_yaw_cs; p.img = d_img; p.missing = d_missing;
    const int sms = sm_count_current_device();
    if (sms <= 0) return fail(SLU_E_DEVICE, "no CUDA device");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (d_missing) SLU_CUDA(cudaMemsetAsync(d_missing, 0, sizeof(int32_t) * B, st));
    long long gx = ((long long)H * W + FR_THREADS - 1) / FR_THREADS;
    const long long cap = (8LL * sms + B - 1) / B;
    if (gx > cap) gx = cap;
    organized_planes_kernel<<<dim3((unsigned)gx, B), FR_THREADS, 0, st>>>(p);
    SLU_LAUNCH_CHECK("organized_planes_kernel");
    return 0;
}

extern "C" int slu_frame_tensors(const float* d_img, int B, int Hs, int Ws, int Hd, int Wd,
                                 const uint8_t* h_flip, float norm_factor,
                                 float* d_range, float* d_refl, float* d_xyz, float* d_normals, int64_t* d_sem,
                                 int32_t* d_rowmap, int resize_rows, slu_stream_t stream) {
    using namespace slu;
    if (!d_img) return fail(SLU_E_ARG, "d_img is NULL");
    if (B < 1 || B > 256) return fail(SLU_E_RANGE, "B=%d outside [1,256]", B);
    if (Hs < 1 || Ws < 1 || Hd < 1 || Wd < 1) return fail(SLU_E_ARG, "image sizes must be >= 1");
    if (d_normals && !d_xyz) return fail(SLU_E_ARG, "normals need the xyz output");
    if (!(norm_factor > 0.f)) return fail(SLU_E_ARG, "norm_factor must be > 0");
    FrameParams p{};
    p.img = d_img; p.B = B; p.Hs = Hs; p.Ws = Ws; p.Hd = Hd; p.Wd = Wd;
    p.ify = 1.0 / ((double)Hd / (double)Hs);          // cv::resize: inv_scale = dsize/ssize, ify = 1/inv_scale
    p.ifx = 1.0 / ((double)Wd / (double)Ws);
    for (int b = 0; b < B && h_flip; ++b)
        if (h_flip[b]) p.flip_bits[b >> 6] |= 1ull << (b & 63);
    const float scale = 1.0f / norm_factor;
    p.tap_c = 10.0f * scale; p.tap_s = 3.0f * scale;
    p.range = d_range; p.refl = d_refl; p.xyz = d_xyz; p.normals = d_normals;
    p.xyz_bstride = 3LL * Hd * Wd;
    p.sem = reinterpret_cast<long long*>(d_sem);
    p.rowmap = d_rowmap;
    p.keep_native_h = (d_rowmap && !resize_rows) ? 1 : 0;
    if (p.keep_native_h && Hd != Hs) return fail(SLU_E_ARG, "without a row resize the output must have Hs rows");
    if (d_rowmap && Hs > 8192) return fail(SLU_E_RANGE, "Hs=%d too large for row compaction", Hs);
    const int sms = sm_count_current_device();
    if (sms <= 0) return fail(SLU_E_DEVICE, "no CUDA device");
    if (d_rowmap) {
        frame_rowmap_kernel<<<B, FR_THREADS, sizeof(int) * Hs, reinterpret_cast<cudaStream_t>(stream)>>>(d_img, Hs, Ws, d_rowmap);
        SLU_LAUNCH_CHECK("frame_rowmap_kernel");
    }
    const long long HWd = (long long)Hd * Wd;
    long long gx = (HWd + FR_THREADS - 1) / FR_THREADS;
    const long long cap = (8LL * sms + B - 1) / B;
    if (gx > cap) gx = cap;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    frame_resample_kernel<<<dim3((unsigned)gx, B), FR_THREADS, 0, st>>>(p);
    SLU_LAUNCH_CHECK("frame_resample_kernel");
    if (d_normals) {
        launch_normals(p, B, sms, st);
        SLU_LAUNCH_CHECK("frame_normals_kernel");
    }
    return 0;
}

/* Normals straight from planar xyz (no resize, no flip): the x, y, z planes of scan b start at d_xyz + b * batch_stride
 * (3*H*W for a packed [B,3,H,W] tensor, 6*H*W for the first three planes of the projection image [B,6,H,W]). */
extern "C" int slu_frame_normals(const float* d_xyz, int B, int H, int W, int64_t batch_stride, float norm_factor,
                                 float* d_normals, slu_stream_t stream) {
    using namespace slu;
    if (!d_xyz || !d_normals) return fail(SLU_E_ARG, "d_xyz / d_normals is NULL");
    if (B < 1 || B > 65535 || H < 1 || W < 1) return fail(SLU_E_RANGE, "B=%d H=%d W=%d unsupported", B, H, W);
    if (batch_stride < 3LL * H * W) return fail(SLU_E_ARG, "batch_stride smaller than three planes");
    if (!(norm_factor > 0.f)) return fail(SLU_E_ARG, "norm_factor must be > 0");
    FrameParams p{};
    p.B = B; p.Hs = H; p.Ws = W; p.Hd = H; p.Wd = W;
    const float scale = 1.0f / norm_factor;
    p.tap_c = 10.0f * scale; p.tap_s = 3.0f * scale;
    p.xyz = const_cast<float*>(d_xyz);
    p.xyz_bstride = batch_stride;
    p.normals = d_normals;
    const int sms = sm_count_current_device();
    if (sms <= 0) return fail(SLU_E_DEVICE, "no CUDA device");
    launch_normals(p, B, sms, reinterpret_cast<cudaStream_t>(stream));
    SLU_LAUNCH_CHECK("frame_normals_kernel");
    return 0;
}
