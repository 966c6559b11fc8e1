// slu_api.cu -- version, error reporting and device queries of libslu.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <atomic>
#include "slu_common.cuh"

namespace slu {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return (int)e;
}

int sm_count_current_device() {
    static thread_local int cached_dev = -1, cached_sms = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (dev != cached_dev) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
        cached_dev = dev;
        cached_sms = sms;
    }
    return cached_sms;
}

bool one_step_bin_search_ok(const float* e, int n_bins) {
    if (!(e[0] >= 0.f) || !(e[n_bins] <= 1.f)) return false;
    for (int i = 0; i < n_bins; ++i) {
        const float lo = e[i], hi = nextafterf(e[i + 1], -1.0f);
        int g0 = (int)(lo * (float)n_bins), g1 = (int)(hi * (float)n_bins);
        g0 = g0 > n_bins - 1 ? n_bins - 1 : g0;
        g1 = g1 > n_bins - 1 ? n_bins - 1 : g1;
        if (g0 < i - 1 || g1 > i + 1) return false;
    }
    return true;
}

}  // namespace slu

extern "C" int slu_version(void) { return SLU_VERSION; }

extern "C" const char* slu_last_error(void) { return slu::g_err; }

extern "C" int64_t slu_launch_count(void) { return (int64_t)slu::g_launches.load(std::memory_order_relaxed); }

extern "C" int slu_device_info(int device, int* sm_count, int* cc_major, int* cc_minor) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return slu::fail(SLU_E_DEVICE, "no CUDA device visible (%s)", e == cudaSuccess ? "count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= n) return slu::fail(SLU_E_ARG, "device %d outside [0,%d)", device, n);
    int sms = 0, maj = 0, min = 0;
    SLU_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    SLU_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, device));
    SLU_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, device));
    if (sm_count) *sm_count = sms;
    if (cc_major) *cc_major = maj;
    if (cc_minor) *cc_minor = min;
    if (maj != 10) return slu::fail(SLU_E_DEVICE, "libslu is built for sm_100a only; device %d is cc %d.%d", device, maj, min);
    return 0;
}

// ---- diagnostic read-only stream (HBM yardstick) ---------------------------------------------------
namespace slu {
__global__ void __launch_bounds__(256) read_stream_kernel(const float4* __restrict__ in, long long n4, float* out) {
    float acc = 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 7 * stride < n4; i += 8 * stride) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
            asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(in + i + u * stride));
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += (v[u].x + v[u].y) + (v[u].z + v[u].w);
    }
    for (; i < n4; i += stride) { const float4 v = __ldg(in + i); acc += (v.x + v.y) + (v.z + v.w); }
    if (acc == 1.2345e-30f) *out = acc;       // keeps the loads alive without a store on the hot path
}
}  // namespace slu

extern "C" int slu_diag_read_stream(const float* d_in, int64_t n, float* d_out, slu_stream_t stream) {
    using namespace slu;
    if (!d_in || !d_out || n < 4 || (n & 3)) return fail(SLU_E_ARG, "slu_diag_read_stream: bad arguments");
    if (reinterpret_cast<uintptr_t>(d_in) & 15) return fail(SLU_E_ALIGN, "d_in not 16-byte aligned");
    const int sms = sm_count_current_device();
    if (sms <= 0) return fail(SLU_E_DEVICE, "no CUDA device");
    read_stream_kernel<<<8 * sms, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(d_in), n / 4, d_out);
    SLU_LAUNCH_CHECK("read_stream_kernel");
    return 0;
}
