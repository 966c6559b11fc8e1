// slu_loss_terms.cu -- the remaining evidential loss terms and regularisers, forward + analytic backward in
// one pass each.  Same contract as slu_loss.cu: a thread owns a pixel, reads its C
// concentrations once, adds the per-pixel value to a float64 block sum and writes d(value)/d(alpha); the
// caller divides by the term's denominator (sums[1]) and applies the upstream gradient, so every term stays
// an independently differentiable scalar (GradNorm: src/utils/grad_norm.py:52).
//
// Terms (reference file:line) and their per-pixel algebra, a0 = sum(alpha), y = target:
//   SLU_TERM_COMP_KL    ComplementKLUniform      src/losses/dirichlet_losses.py:228-314
//        A = a0 + e, p = alpha/A, py = max(p_y, e), d = max(1 - py, e), t_c = p_c/d (c != y)
//        kl = sum_{c!=y} t_c log max(t_c, e) + log(C-1)   [/ log(C-1) if normalize]
//        w  = (1-py)^gamma sigmoid((tau - py)/sigma) [* s/(A + s) if s_target]; value = w kl, sums[1] += 1
//        backward, with L_c = log max(t_c,e) + [t_c > e], S = sum L_c t_c, k = w / log(C-1):
//          G_c = k L_c / d (c != y),  G_y = k S [1-py > e][p_y >= e] / d + E,  E = kl dw/dpy if the gate is not detached
//          dvalue/dalpha_j = (G_j - sum_c G_c p_c) / A;  with no clamp active this is (k/d (L_j - S) - E p_y)/A for
//          j != y and E (1 - p_y)/A for j = y
//   SLU_TERM_WRONG_LOW  WrongLowEvidence         src/losses/regularizers.py:218-289
//        a0' = max(a0, e), p = alpha/a0', wrong = argmax p != y, m = max(pmax,e) - max(p_y,e)
//        gate = wrong * {1 | [m > margin] | sigmoid((m - margin)/k)}   (no gradient)
//        value = gate relu(log a0' - log(C + s_low + e))^2,  sums[1] += gate
//        dvalue/dalpha_j = 2 gate relu(.) / a0'   (every class; 0 if the a0 clamp is active)
//   SLU_TERM_EVID_BAND  EvidenceRegBand          :116-147   a0' = a0 + 1e-8
//        value = relu(log(a0'/s_hi))^2 + relu(log(s_lo/a0'))^2,  d/dalpha_j = 2 (over - under)/a0'
//   SLU_TERM_EVID_REG   EvidenceReg              :149-212   a0' = a0 + 1e-8
//        log_squared: lr = log(a0'/s); value = lr^2 [* a0'/s];  d = 2 lr/a0'  [(lr^2 + 2 lr)/s]
//        one_sided:   r = relu(a0' - s(1+margin)); value = r^2; d = 2 r        l2: value = (a0'-s)^2; d = 2(a0'-s)
//   SLU_TERM_KL_CONF    KL_offClasses_to_uniform(with_conf_weighting=True)  :291-389
//        w = clamp(1 - alpha_y/(a0 + e), 0, 1)^gamma (no gradient); value = w KL, sums[1] += w, grad = w dKL/dalpha
//   slu_logit_regularizer  LogitRegularizer      :75-110    value = z^2 or relu(z - thr)^2 per element
// Bound: all of these read 80-88 B/px and write 80 B/px; COMP_KL (C logs) and KL_CONF (lgamma/digamma/trigamma per
// class) are instruction-bound like the terms of slu_loss.cu, the others stream.
#include <math.h>
#include "slu_common.cuh"
#include "slu_special.cuh"

namespace slu {

constexpr int TERM_THREADS = 256;
constexpr int TERM_MAX_IGNORE = 8;
constexpr int TERM_MAX_PRM = 8;

struct EvTermParams {
    const float* alpha;
    const long long* target;       // may be NULL for the a0-only regularisers
    const unsigned char* keep;
    int B, C;
    long long HW, n_px;
    long long ignore[TERM_MAX_IGNORE];
    int n_ignore;
    int term;
    float prm[TERM_MAX_PRM];
    double* sums;                  // [2] sum of values | denominator
    float* grad;                   // [B,C,HW] or NULL
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int CP>
__global__ void __launch_bounds__(TERM_THREADS) evidence_term_kernel(const __grid_constant__ EvTermParams p) {
    const int tid = threadIdx.x;
    double acc = 0.0, den = 0.0;
    const long long chunks = (p.n_px + TERM_THREADS - 1) / TERM_THREADS;
    for (long long ch = blockIdx.x; ch < chunks; ch += gridDim.x) {
        const long long g = ch * TERM_THREADS + tid;
        if (g >= p.n_px) continue;
        const int b = (int)(g / p.HW);
        const long long px = g - (long long)b * p.HW;
        const long long tgt = p.target ? p.target[g] : 0;
        bool valid = true;
        if (p.keep) {
            valid = p.keep[g] != 0;
        } else if (p.target) {
#pragma unroll
            for (int i = 0; i < TERM_MAX_IGNORE; ++i)
                if (i < p.n_ignore && tgt == p.ignore[i]) valid = false;
        }
        const float* base = p.alpha + ((long long)b * p.C) * p.HW + px;
        float* go = p.grad ? p.grad + ((long long)b * p.C) * p.HW + px : nullptr;
        if (!valid) {
            if (go) for (int c = 0; c < p.C; ++c) go[(long long)c * p.HW] = 0.f;
            continue;
        }
        const int y = (int)tgt;
        float a[CP];
        float a0 = 0.f, ay = 0.f;
#pragma unroll
        for (int c = 0; c < CP; ++c) {
            a[c] = (c < p.C) ? ldg_stream(base + (long long)c * p.HW) : 0.f;
            a0 += a[c];
            if (c == y) ay = a[c];
        }

        if (p.term == SLU_TERM_COMP_KL) {
            const float gamma = p.prm[0], tau = p.prm[1], sigma = p.prm[2], s_t = p.prm[3], eps = p.prm[5];
            const bool normalize = p.prm[4] != 0.f, detach = p.prm[6] != 0.f;
            const float A = a0 + eps, invA = 1.0f / A;
            const float py_raw = ay * invA;
            const float py = fmaxf(py_raw, eps);
            // 1 - p_y without the cancellation the literal fp32 form has at confident pixels: (sum_{c!=y} alpha_c + e)/A
            float off = 0.f;
#pragma unroll
            for (int c = 0; c < CP; ++c)
                if (c != y) off += a[c];
            const float om = (py_raw >= eps) ? (off + eps) * invA : 1.0f - eps;
            const float d = fmaxf(om, eps), invd = 1.0f / d;
            const float logCm1 = logf((float)(p.C - 1));
            // L_c is kept relative to the first off-class' value: the gradient only needs differences L_j - sum t_c L_c,
            // and centring removes their common part before it can cancel (uniform alpha gives an exact zero)
            float S = 0.f, kl = 0.f, Lref = 0.f, T = 0.f;
            bool have = false;
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                if (c < p.C && c != y) {
                    const float t = a[c] * invA * invd;
                    const float lt = logf(fmaxf(t, eps));
                    kl = fmaf(t, lt, kl);
                    const float L = lt + (t > eps ? 1.0f : 0.0f);
                    if (!have) { Lref = L; have = true; }
                    S = fmaf(L - Lref, t, S);
                    T += t;
                    a[c] = L - Lref;                // a[] now holds the centred L_c for c != y
                }
            }
            kl += logCm1;
            const float nrm = normalize ? 1.0f / logCm1 : 1.0f;
            kl *= nrm;
            const float sg = sigmoidf_((tau - py) / sigma);
            const float pw = powf(om, gamma);
            float w = pw * sg;
            float wev = 1.0f;
            if (s_t >= 0.f) wev = s_t / (A + s_t);
            acc += (double)(w * wev * kl);
            den += 1.0;
            if (go) {
                const float k = w * wev * nrm;
                const bool f_d = om > eps, f_py = py_raw >= eps;
                float E = 0.f;                      // d(value)/d(p_y) through the gate, when it is not detached
                if (!detach && f_py) {
                    const float dpw = (om > 0.f) ? -gamma * powf(om, gamma - 1.0f) : 0.0f;
                    E = kl * wev * (dpw * sg - pw * sg * (1.0f - sg) / sigma);
                }
                const float kd = k * invd;
                if (f_d && f_py) {
                    // no clamp active: sum_c t_c = 1 and d + p_y = 1, so dvalue/dalpha_j = (k/d (L_j - S) - E p_y)/A, dvalue/dalpha_y = E (1-p_y)/A
                    const float Epy = E * py_raw;
#pragma unroll
                    for (int c = 0; c < CP; ++c)
                        if (c < p.C) go[(long long)c * p.HW] = (c == y ? E * om : fmaf(kd, a[c] - S, -Epy)) * invA;
                } else {
                    // a clamp is active (p_y < e or 1 - p_y < e): general chain rule with the uncentred sums
                    const float Sfull = fmaf(Lref, T, S);
                    const float Gy = E;             // the clamp cuts the path through d
                    const float Gp = fmaf(Gy, py_raw, k * Sfull);
#pragma unroll
                    for (int c = 0; c < CP; ++c)
                        if (c < p.C) go[(long long)c * p.HW] = ((c == y ? Gy : kd * (a[c] + Lref)) - Gp) * invA;
                }
            }
        } else if (p.term == SLU_TERM_WRONG_LOW) {
            const float s_low = p.prm[0], margin = p.prm[1], kk = p.prm[2], eps = p.prm[3];
            const float a0c = fmaxf(a0, eps);
            float pmax = -1.f, pyv = 0.f;
            int pred = 0;
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                if (c < p.C) {
                    const float pc = __fdiv_rn(a[c], a0c);
                    if (pc > pmax) { pmax = pc; pred = c; }      // first maximum, as torch.argmax
                    if (c == y) pyv = pc;
                }
            }
            float gate = 0.f;
            if (pred != y) {
                const float m = fmaxf(pmax, eps) - fmaxf(pyv, eps);
                if (margin > 0.f) gate = kk > 0.f ? sigmoidf_((m - margin) / kk) : (m > margin ? 1.f : 0.f);
                else gate = 1.f;
            }
            const float r = fmaxf(logf(a0c) - logf((float)p.C + s_low + eps), 0.f);
            acc += (double)(r * r * gate);
            den += (double)gate;
            if (go) {
                const float gj = (a0 >= eps) ? 2.0f * r * gate / a0c : 0.f;
#pragma unroll
                for (int c = 0; c < CP; ++c)
                    if (c < p.C) go[(long long)c * p.HW] = gj;
            }
        } else if (p.term == SLU_TERM_EVID_BAND || p.term == SLU_TERM_EVID_REG) {
            const float A = a0 + 1e-8f;
            float val, gj;
            if (p.term == SLU_TERM_EVID_BAND) {
                const float s = p.prm[0], band = p.prm[1];
                const float over = fmaxf(logf(A / (s * (1.0f + band))), 0.f);
                const float under = fmaxf(logf((s * (1.0f - band)) / A), 0.f);
                val = over * over + under * under;
                gj = 2.0f * (over - under) / A;
            } else {
                const float s = p.prm[0], margin = p.prm[2];
                const int mode = (int)p.prm[1];
                if (mode == 0) {
                    const float lr = logf(A / s);
                    if (p.prm[3] != 0.f) { val = (A / s) * lr * lr; gj = (lr * lr + 2.0f * lr) / s; }
                    else { val = lr * lr; gj = 2.0f * lr / A; }
                } else if (mode == 1) {
                    const float r = fmaxf(A - s * (1.0f + margin), 0.f);
                    val = r * r; gj = 2.0f * r;
                } else {
                    const float r = A - s;
                    val = r * r; gj = 2.0f * r;
                }
            }
            acc += (double)val;
            den += 1.0;
            if (go) {
#pragma unroll
                for (int c = 0; c < CP; ++c)
                    if (c < p.C) go[(long long)c * p.HW] = gj;
            }
        } else {   // SLU_TERM_KL_CONF
            const float gamma = p.prm[0], eps = p.prm[1];
            const float om = fminf(fmaxf(1.0f - ay / (a0 + eps), 0.f), 1.f);
            const float w = powf(om, gamma);
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < CP; ++c)
                if (c < p.C) s += fmaxf(c == y ? 1.0f : a[c], eps);
            const LDT fs = ldt_pos(s);
            float kl = fs.lg;
            const float tail = (s - (float)p.C) * fs.tri;
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                if (c < p.C) {
                    float gj = 0.f;
                    if (c != y) {
                        const float ac = fmaxf(a[c], eps);
                        const LDT f = ldt_pos(ac);
                        kl -= f.lg;
                        kl = fmaf(ac - 1.0f, f.psi - fs.psi, kl);
                        if (a[c] > eps) gj = w * fmaf(ac - 1.0f, f.tri, -tail);
                    }
                    if (go) go[(long long)c * p.HW] = gj;
                }
            }
            acc += (double)(w * kl);
            den += (double)w;
        }
    }
    __shared__ double s_v[TERM_THREADS / 32], s_d[TERM_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        den += __shfl_xor_sync(0xffffffffu, den, o);
    }
    if ((tid & 31) == 0) { s_v[tid >> 5] = acc; s_d[tid >> 5] = den; }
    __syncthreads();
    if (tid == 0) {
        double v = 0.0, n = 0.0;
        for (int i = 0; i < TERM_THREADS / 32; ++i) { v += s_v[i]; n += s_d[i]; }
        atomicAdd(&p.sums[0], v);
        atomicAdd(&p.sums[1], n);
    }
}

template <int CP>
static int launch_ev_term(const EvTermParams& p, cudaStream_t st) {
    const int sms = sm_count_current_device();
    if (sms <= 0) return fail(SLU_E_DEVICE, "no CUDA device");
    const long long chunks = (p.n_px + TERM_THREADS - 1) / TERM_THREADS;
    const long long cap = 6LL * sms;
    evidence_term_kernel<CP><<<(unsigned)(chunks < cap ? chunks : cap), TERM_THREADS, 0, st>>>(p);
    SLU_LAUNCH_CHECK("evidence_term_kernel");
    return 0;
}

// ---- LogitRegularizer: elementwise over [B,Cz,HW] with a per-pixel mask ---------------------------------
struct LogitRegParams {
    const float* z;
    const long long* target;
    const unsigned char* keep;
    int B, Cz;
    long long HW;
    long long ignore[TERM_MAX_IGNORE];
    int n_ignore;
    int use_thr;
    float thr;
    double* sums;                  // [2] sum of per-element values | number of valid PIXELS
    float* grad;
};

__global__ void __launch_bounds__(TERM_THREADS) logit_reg_kernel(const __grid_constant__ LogitRegParams p) {
    const int tid = threadIdx.x;
    double acc = 0.0;
    unsigned n_valid = 0;
    const long long n_px = (long long)p.B * p.HW;
    for (long long g = (long long)blockIdx.x * TERM_THREADS + tid; g < n_px; g += (long long)gridDim.x * TERM_THREADS) {
        const int b = (int)(g / p.HW);
        const long long px = g - (long long)b * p.HW;
        bool valid = true;
        if (p.keep) {
            valid = p.keep[g] != 0;
        } else if (p.target) {
            const long long tgt = p.target[g];
#pragma unroll
            for (int i = 0; i < TERM_MAX_IGNORE; ++i)
                if (i < p.n_ignore && tgt == p.ignore[i]) valid = false;
        }
        const float* base = p.z + ((long long)b * p.Cz) * p.HW + px;
        float* go = p.grad ? p.grad + ((long long)b * p.Cz) * p.HW + px : nullptr;
        if (valid) ++n_valid;
        float v = 0.f;
        for (int c = 0; c < p.Cz; ++c) {
            float gz = 0.f;
            if (valid) {
                const float z = ldg_stream(base + (long long)c * p.HW);
                const float r = p.use_thr ? fmaxf(z - p.thr, 0.f) : z;
                v = fmaf(r, r, v);
                gz = 2.0f * r;
            }
            if (go) go[(long long)c * p.HW] = gz;
        }
        acc += (double)v;
    }
    __shared__ double s_v[TERM_THREADS / 32];
    __shared__ unsigned s_n[TERM_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        n_valid += __shfl_xor_sync(0xffffffffu, n_valid, o);
    }
    if ((tid & 31) == 0) { s_v[tid >> 5] = acc; s_n[tid >> 5] = n_valid; }
    __syncthreads();
    if (tid == 0) {
        double v = 0.0, n = 0.0;
        for (int i = 0; i < TERM_THREADS / 32; ++i) { v += s_v[i]; n += (double)s_n[i]; }
        atomicAdd(&p.sums[0], v);
        atomicAdd(&p.sums[1], n);
    }
}

}  // namespace slu

extern "C" int slu_evidence_term(const float* d_alpha, const int64_t* d_target, const uint8_t* d_keep_mask,
                                 int B, int C, int64_t HW, const int64_t* h_ignore, int n_ignore,
                                 int term, const float* h_params, int n_params,
                                 double* d_sums, float* d_grad, slu_stream_t stream) {
    using namespace slu;
    if (!d_alpha || !d_sums) return fail(SLU_E_ARG, "d_alpha / d_sums is NULL");
    if (B < 1 || HW < 1) return fail(SLU_E_ARG, "B=%d HW=%lld must be >= 1", B, (long long)HW);
    if (C < 2 || C > SLU_MAX_CLASSES) return fail(SLU_E_RANGE, "C=%d outside [2,%d]", C, SLU_MAX_CLASSES);
    if (n_ignore < 0 || n_ignore > TERM_MAX_IGNORE || (n_ignore > 0 && !h_ignore)) return fail(SLU_E_RANGE, "n_ignore=%d outside [0,%d]", n_ignore, TERM_MAX_IGNORE);
    int need;
    switch (term) {
        case SLU_TERM_COMP_KL: need = 7; break;
        case SLU_TERM_WRONG_LOW: need = 4; break;
        case SLU_TERM_EVID_BAND: need = 2; break;
        case SLU_TERM_EVID_REG: need = 4; break;
        case SLU_TERM_KL_CONF: need = 2; break;
        default: return fail(SLU_E_ARG, "unknown term %d", term);
    }
    if (!h_params || n_params != need) return fail(SLU_E_ARG, "term %d takes %d parameters, got %d", term, need, n_params);
    const bool needs_target = term == SLU_TERM_COMP_KL || term == SLU_TERM_WRONG_LOW || term == SLU_TERM_KL_CONF;
    if (needs_target && !d_target) return fail(SLU_E_ARG, "term %d needs d_target", term);
    if (term == SLU_TERM_COMP_KL && C < 3) return fail(SLU_E_RANGE, "SLU_TERM_COMP_KL needs C >= 3 (the reference returns 0 for C <= 2)");
    if (term == SLU_TERM_COMP_KL && !(h_params[2] > 0.f)) return fail(SLU_E_ARG, "sigma must be > 0");
    EvTermParams p{};
    p.alpha = d_alpha; p.target = reinterpret_cast<const long long*>(d_target); p.keep = d_keep_mask;
    p.B = B; p.C = C; p.HW = HW; p.n_px = (long long)B * HW;
    for (int i = 0; i < n_ignore; ++i) p.ignore[i] = h_ignore[i];
    p.n_ignore = n_ignore; p.term = term;
    for (int i = 0; i < need; ++i) p.prm[i] = h_params[i];
    p.sums = d_sums; p.grad = d_grad;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch ((C + 3) / 4 * 4) {
        case 4: return launch_ev_term<4>(p, st);
        case 8: return launch_ev_term<8>(p, st);
        case 12: return launch_ev_term<12>(p, st);
        case 16: return launch_ev_term<16>(p, st);
        case 20: return launch_ev_term<20>(p, st);
        case 24: return launch_ev_term<24>(p, st);
        case 28: return launch_ev_term<28>(p, st);
        default: return launch_ev_term<32>(p, st);
    }
}

extern "C" int slu_logit_regularizer(const float* d_logits, const int64_t* d_target, const uint8_t* d_keep_mask,
                                     int B, int Cz, int64_t HW, const int64_t* h_ignore, int n_ignore,
                                     int use_threshold, float threshold, double* d_sums, float* d_grad,
                                     slu_stream_t stream) {
    using namespace slu;
    if (!d_logits || !d_sums) return fail(SLU_E_ARG, "d_logits / d_sums is NULL");
    if (B < 1 || HW < 1 || Cz < 1) return fail(SLU_E_ARG, "B=%d Cz=%d HW=%lld must be >= 1", B, Cz, (long long)HW);
    if (n_ignore < 0 || n_ignore > TERM_MAX_IGNORE || (n_ignore > 0 && !h_ignore)) return fail(SLU_E_RANGE, "n_ignore=%d outside [0,%d]", n_ignore, TERM_MAX_IGNORE);
    LogitRegParams p{};
    p.z = d_logits; p.target = reinterpret_cast<const long long*>(d_target); p.keep = d_keep_mask;
    p.B = B; p.Cz = Cz; p.HW = HW;
    for (int i = 0; i < n_ignore; ++i) p.ignore[i] = h_ignore[i];
    p.n_ignore = n_ignore; p.use_thr = use_threshold ? 1 : 0; p.thr = threshold;
    p.sums = d_sums; p.grad = d_grad;
    const int sms = sm_count_current_device();
    if (sms <= 0) return fail(SLU_E_DEVICE, "no CUDA device");
    const long long chunks = ((long long)B * HW + TERM_THREADS - 1) / TERM_THREADS;
    const long long cap = 8LL * sms;
    logit_reg_kernel<<<(unsigned)(chunks < cap ? chunks : cap), TERM_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    SLU_LAUNCH_CHECK("logit_reg_kernel");
    return 0;
}
