// slu_api.cu -- version, error reporting and device queries of libslu.
#include <stdarg.h>
#include <stdio.h>
#include "slu_common.cuh"

namespace slu {

static thread_local char g_err[512] = "";

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

}  // namespace slu

extern "C" int slu_version(void) { return SLU_VERSION; }

extern "C" const char* slu_last_error(void) { return slu::g_err; }

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
