// slu_special.cuh -- fp32 digamma / trigamma for positive arguments (evidential kernels).
//
// The reference calls torch.special.digamma / torch.digamma (src/models/probability_helper.py:128,
// src/losses/regularizers.py:335-337); ATen's float kernel is the cephes scheme: recurrence
// psi(x) = psi(x+1) - 1/x up to x >= 10, then the asymptotic series.  Same scheme here with the
// recurrence stopped at x >= 6 (the series' first dropped term is then 691/(32760 x^12) < 2e-11),
// and the whole evaluation kept in fp32.  Parity is by tolerance (1e-5 relative on the reduced maps).
#pragma once
#include <cuda_runtime.h>

namespace slu {

// psi(x), x > 0
__device__ __forceinline__ float digamma_pos(float x) {
    float r = 0.f;
#pragma unroll 1
    while (x < 2.f) {                 // rare in this code base: arguments are alpha + 1 >= 2
        r -= __frcp_rn(x);
        x += 1.f;
    }
    if (x < 6.f) {
        // four recurrence steps at once: 1/x + 1/(x+1) + 1/(x+2) + 1/(x+3) = (2x+3)(2u+2) / (u(u+2)), u = x(x+3)
        const float u = x * (x + 3.f);
        r -= __fdividef(fmaf(2.f, x, 3.f) * fmaf(2.f, u, 2.f), u * (u + 2.f));
        x += 4.f;
    }
    const float inv = __frcp_rn(x);
    const float z = inv * inv;
    // 1/12 z - 1/120 z^2 + 1/252 z^3 - 1/240 z^4 + 1/132 z^5
    float y = fmaf(z, 7.57575757575757576e-3f, -4.16666666666666667e-3f);
    y = fmaf(z, y, 3.96825396825396825e-3f);
    y = fmaf(z, y, -8.33333333333333333e-3f);
    y = fmaf(z, y, 8.33333333333333333e-2f);
    y *= z;
    return r + (logf(x) - 0.5f * inv - y);
}

// psi'(x), x > 0
__device__ __forceinline__ float trigamma_pos(float x) {
    float r = 0.f;
#pragma unroll 1
    while (x < 6.f) {
        const float i = __frcp_rn(x);
        r = fmaf(i, i, r);
        x += 1.f;
    }
    const float inv = __frcp_rn(x);
    const float z = inv * inv;
    // 1/x + 1/(2x^2) + 1/(6x^3) - 1/(30x^5) + 1/(42x^7) - 1/(30x^9) + 5/(66 x^11)
    float y = fmaf(z, 7.57575757575757576e-2f, -3.33333333333333333e-2f);
    y = fmaf(z, y, 2.38095238095238095e-2f);
    y = fmaf(z, y, -3.33333333333333333e-2f);
    y = fmaf(z, y, 1.66666666666666667e-1f);
    y = fmaf(y, z * inv, fmaf(0.5f, z, inv));
    return r + y;
}

}  // namespace slu
