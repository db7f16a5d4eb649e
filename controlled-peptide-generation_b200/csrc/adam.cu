// Global-norm gradient clipping + Adam on the flat parameter buffer, reproducing the
// reference optimizer semantics of train_vae.py:15,41-42 -- including the fact that
// `RNN_VAE.vae_params()` (models/model.py:88-94) yields the shared embedding matrix TWICE:
//   (i)  its squared gradient norm enters the total norm twice,
//   (ii) `clip_grad_norm_` scales its gradient once per list entry (coef^2 overall),
//   (iii) Adam applies two sequential updates per iteration to it (its step counter += 2).
#include "kernels.h"
#include "adam.h"

namespace cpg {

constexpr int NORM_THREADS = 256;

__global__ void __launch_bounds__(NORM_THREADS)
k_sumsq_partial(const float* __restrict__ g, int64_t n, int64_t dup_off, int64_t dup_n, float* __restrict__ part) {
    __shared__ float red[NORM_THREADS / 32];
    float s = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float v = g[i];
        float w = (i >= dup_off && i < dup_off + dup_n) ? 2.f : 1.f;
        s = fmaf(w * v, v, s);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < NORM_THREADS / 32; ++w) t += red[w];
        part[blockIdx.x] = t;
    }
}

__device__ __forceinline__ void adam_update(float& p, float& m, float& v, float g, const AdamStep& st, float b1,
                                            float b2, float eps) {
    m = m + (g - m) * (1.0f - b1);                       // exp_avg.lerp_(grad, 1 - beta1)
    v = v * b2 + (1.0f - b2) * g * g;                    // mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    float denom = sqrtf(v) / st.bc2_sqrt + eps;
    p = p - st.step_size * (m / denom);                  // addcdiv_(exp_avg, denom, value=-step_size)
}

// The total norm is finished here, by every block on its own (the same ordered sum of the same partials: identical
// in every block and run to run), instead of in a one-thread kernel between the two passes.
__global__ void __launch_bounds__(NORM_THREADS)
k_clip_adam(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            int64_t n, int64_t dup_off, int64_t dup_n, const float* __restrict__ part, int nparts, float max_norm,
            float* __restrict__ norm_out, float* __restrict__ coef_out, AdamHyper h, const StepDyn* __restrict__ dyn) {
    __shared__ float coef_s;
    if (dyn != nullptr) {
        h.single = AdamStep{dyn->step_size[0], dyn->bc2_sqrt[0]};
        h.dup_first = AdamStep{dyn->step_size[1], dyn->bc2_sqrt[1]};
        h.dup_second = AdamStep{dyn->step_size[2], dyn->bc2_sqrt[2]};
    }
    if (threadIdx.x < 32) {
        double s = 0.0;
        for (int i = threadIdx.x; i < nparts; i += 32) s += (double)part[i];
        s = warp_sum(s);
        if (threadIdx.x == 0) {
            const float total = (float)sqrt(s);
            const float c = max_norm / (total + 1e-6f);           // torch: clamp(max_norm / (total + 1e-6), max=1)
            coef_s = c < 1.0f ? c : 1.0f;
            if (blockIdx.x == 0) {
                if (norm_out != nullptr) *norm_out = total;
                if (coef_out != nullptr) *coef_out = coef_s;
            }
        }
    }
    __syncthreads();
    const float coef = coef_s;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const bool dup = (i >= dup_off && i < dup_off + dup_n);
        float gi = g[i] * coef;
        if (dup) gi *= coef;
        g[i] = gi;
        float pi = p[i], mi = m[i], vi = v[i];
        if (dup) {
            adam_update(pi, mi, vi, gi, h.dup_first, h.beta1, h.beta2, h.eps);
            adam_update(pi, mi, vi, gi, h.dup_second, h.beta1, h.beta2, h.eps);
        } else {
            adam_update(pi, mi, vi, gi, h.single, h.beta1, h.beta2, h.eps);
        }
        p[i] = pi; m[i] = mi; v[i] = vi;
    }
}

int adam_norm_parts(int64_t n, int sm_count) {
    int64_t want = (n + NORM_THREADS * 4 - 1) / (NORM_THREADS * 4);
    int cap = 2 * (sm_count > 0 ? sm_count : 1);
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

// Both passes in ONE launch: every block publishes the sum of squares of its slice, passes a grid-wide barrier (all
// blocks are co-resident: at most 2 per SM; sense-reversing counter that needs no reset between launches), finishes the
// total norm by itself (the same ordered sum of the same partials in every block) and updates its slice.
__global__ void __launch_bounds__(NORM_THREADS)
k_sumsq_clip_adam(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                  int64_t n, int64_t dup_off, int64_t dup_n, float* __restrict__ part, unsigned* __restrict__ bar, float max_norm,
                  float* __restrict__ norm_out, float* __restrict__ coef_out, AdamHyper h, const StepDyn* __restrict__ dyn) {
    __shared__ float red[NORM_THREADS / 32];
    __shared__ float coef_s;
    if (dyn != nullptr) {
        h.single = AdamStep{dyn->step_size[0], dyn->bc2_sqrt[0]};
        h.dup_first = AdamStep{dyn->step_size[1], dyn->bc2_sqrt[1]};
        h.dup_second = AdamStep{dyn->step_size[2], dyn->bc2_sqrt[2]};
    }
    float s = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float gv = g[i];
        const float w = (i >= dup_off && i < dup_off + dup_n) ? 2.f : 1.f;
        s = fmaf(w * gv, gv, s);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < NORM_THREADS / 32; ++w) t += red[w];
        part[blockIdx.x] = t;
        // grid barrier: bar[0] = arrivals of this round, bar[1] = generation
        volatile unsigned* vb = bar;
        const unsigned gen = vb[1];
        __threadfence();
        if (atomicAdd(bar, 1u) == gridDim.x - 1) {
            vb[0] = 0;
            __threadfence();
            atomicAdd(bar + 1, 1u);
        } else {
            while (vb[1] == gen) {}
        }
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const volatile float* vp = part;
        double t = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) t += (double)vp[i];
        t = warp_sum(t);
        if (threadIdx.x == 0) {
            const float total = (float)sqrt(t);
            const float c = max_norm / (total + 1e-6f);           // torch: clamp(max_norm / (total + 1e-6), max=1)
            coef_s = c < 1.0f ? c : 1.0f;
            if (blockIdx.x == 0) {
                if (norm_out != nullptr) *norm_out = total;
                if (coef_out != nullptr) *coef_out = coef_s;
            }
        }
    }
    __syncthreads();
    const float coef = coef_s;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const bool dup = (i >= dup_off && i < dup_off + dup_n);
        float gi = g[i] * coef;
        if (dup) gi *= coef;
        g[i] = gi;
        float pi = p[i], mi = m[i], vi = v[i];
        if (dup) {
            adam_update(pi, mi, vi, gi, h.dup_first, h.beta1, h.beta2, h.eps);
            adam_update(pi, mi, vi, gi, h.dup_second, h.beta1, h.beta2, h.eps);
        } else {
            adam_update(pi, mi, vi, gi, h.single, h.beta1, h.beta2, h.eps);
        }
        p[i] = pi; m[i] = mi; v[i] = vi;
    }
}

// `bar`: two zero-initialised unsigned words owned by the context (grid barrier); null = two launches
void launch_clip_adam_fused(cudaStream_t s, float* p, float* g, float* m, float* v, int64_t n, int64_t dup_off,
                            int64_t dup_n, float max_norm, float* part, float* norm_out, float* coef_out,
                            const AdamHyper& h, int sm_count, unsigned* bar) {
    int parts = adam_norm_parts(n, sm_count);
#ifndef CPG_EMU
    if (bar != nullptr) {
        CPG_LAUNCH(k_sumsq_clip_adam, parts, NORM_THREADS, 0, s, p, g, m, v, n, dup_off, dup_n, part, bar, max_norm, norm_out,
                   coef_out, h, g_dyn);
        return;
    }
#endif
    CPG_LAUNCH(k_sumsq_partial, parts, NORM_THREADS, 0, s, g, n, dup_off, dup_n, part);
    CPG_LAUNCH(k_clip_adam, parts, NORM_THREADS, 0, s, p, g, m, v, n, dup_off, dup_n, part, parts, max_norm, norm_out,
               coef_out, h, g_dyn);
}

}  // namespace cpg
