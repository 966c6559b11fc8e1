// Dependent-issue latency of FFMA / FFMA2 / MUFU on sm_100a: ONE chain per thread, one warp per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 8192
template <int MODE>
__global__ void k(float* out, float seed, long long* cycles) {
    float2 a = make_float2(seed, seed * 0.5f);
    const float2 b = make_float2(0.999f, 1.001f), c = make_float2(0.001f, -0.001f);
    const long long t0 = clock64();
#pragma unroll 16
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) a.x = fmaf(a.x, b.x, c.x);
        if (MODE == 1) a = __ffma2_rn(a, b, c);
        if (MODE == 2) a = __ffma2_rn(a, b, make_float2(0.001f, 0.001f));
        if (MODE == 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a.x));
        if (MODE == 4) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a.x));
        if (MODE == 5) a.x = a.x > 0.5f ? a.x * 0.999f : a.x + 0.25f;      // FSETP + FSEL style dependent pair
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a.x + a.y;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int MODE> void run(const char* name) {
    float* out; long long* cyc; long long h = 0;
    cudaMalloc(&out, 148 * 128 * sizeof(float)); cudaMalloc(&cyc, 8);
    k<MODE><<<148, 32>>>(out, 1.0001f, cyc); k<MODE><<<148, 32>>>(out, 1.0001f, cyc);
    cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s %.2f cycles per dependent step\n", name, (double)h / ITERS);
}
int main() {
    run<0>("FFMA"); run<1>("FFMA2 reg"); run<2>("FFMA2 imm"); run<3>("MUFU.EX2"); run<4>("MUFU.RCP"); run<5>("compare+select+mul/add");
    return 0;
}
