// Micro-benchmark (sm_100a): issue cost of scalar FFMA (register / immediate form) against packed FFMA2 (register /
// immediate form), per SM sub-partition.  8 independent chains per thread, 8 warps per SM sub-partition... prints
// warp-instructions per cycle per SM and fp32 FMAs per cycle per SM.   nvcc -arch=sm_100a -O3 -o fma_rates fma_rates.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CHAINS 8
#define ITERS 4096

template <int MODE>
__global__ void k(float* out, float seed, long long* cycles) {
    float2 a[CHAINS];
    for (int i = 0; i < CHAINS; ++i) a[i] = make_float2(seed + i, seed - i);
    const float2 b = make_float2(seed * 0.5f, seed * 0.25f);
    const float2 c = make_float2(seed * 0.125f, seed * 0.0625f);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }                  // 2 scalar FFMA, registers
            if (MODE == 1) { a[i].x = fmaf(a[i].x, b.x, 1.25f); a[i].y = fmaf(a[i].y, b.y, -0.75f); }             // 2 scalar FFMA, immediate addend
            if (MODE == 2) a[i] = __ffma2_rn(a[i], b, c);                                                          // 1 FFMA2, registers
            if (MODE == 3) a[i] = __ffma2_rn(a[i], b, make_float2(1.25f, 1.25f));                                  // 1 FFMA2, immediate addend
            if (MODE == 4) { a[i].x = a[i].x * b.x; a[i].y = a[i].y + c.y; }                                       // FMUL + FADD scalar
            if (MODE == 5) a[i] = __fmul2_rn(a[i], b);                                                             // FMUL2
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < CHAINS; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE>
void run(const char* name, int fma_per_instr, int instr_per_step) {
    float* out; long long* cyc; long long h = 0;
    cudaMalloc(&out, 148 * 1024 * sizeof(float)); cudaMalloc(&cyc, 8);
    for (int warps = 4; warps <= 32; warps *= 2) {
        k<MODE><<<148, warps * 32>>>(out, 1.0001f, cyc);
        k<MODE><<<148, warps * 32>>>(out, 1.0001f, cyc);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        const double winstr = (double)ITERS * CHAINS * instr_per_step * warps;      // warp instructions per SM
        printf("%-34s warps/SM %2d: %.2f warp-instr/clk/SM, %.1f fp32 FMA-lanes/clk/SM\n", name, warps, winstr / h,
               winstr * 32 * fma_per_instr / h);
    }
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("FFMA reg (2 per step)", 1, 2);
    run<1>("FFMA imm addend (2 per step)", 1, 2);
    run<2>("FFMA2 reg", 2, 1);
    run<3>("FFMA2 imm addend", 2, 1);
    run<4>("FMUL + FADD scalar", 1, 2);
    run<5>("FMUL2", 2, 1);
    return 0;
}
