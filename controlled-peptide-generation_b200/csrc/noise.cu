// Counter-based (Philox4x32-10) generation of the per-iteration noise of the WAE step, so that
// "perf mode" draws nothing on the host.  The reference interleaves three generators
// (torch-CPU randn for eps, numpy for c and word dropout, the device
// generator for nn.Dropout / randn_like); here every tensor is a pure function of
// (seed, step, element index), which also makes data-parallel ranks reproducible.
// Parity runs bypass this file: they pass the reference's own noise tensors in.
#include "ctx.h"

namespace cpg {
int check_launch(const char* where);

__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
    float u1 = u32_to_unit_open(a), u2 = u32_to_unit_open(b);
    float r = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincosf(6.283185307179586f * u2, &s, &c);
    n0 = r * c;
    n1 = r * s;
}


// stream ids: 16*step + {0: normals, 1: c, 2: word dropout, 3: out dropout}
// part: NOISE_WORD = word dropout (needed first), NOISE_LATENT = eps, c, NOISE_LATE = what is only needed later
// (z_prior x2, out-dropout mask -- the bulk of the work)
__global__ void k_step_noise(StepNoiseArgs a, int part, const StepDyn* __restrict__ dyn) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n_lat = (int64_t)a.B * ZD;
    const int64_t n_tok = (int64_t)a.B * a.L;
    const int64_t n_out4 = ((int64_t)a.B * a.L * DEC_H + 3) / 4;
    const uint32_t base = (dyn != nullptr ? dyn->noise_step : a.step) * 16u;
    uint32_t r[4];
    if (i < n_lat && (part & (NOISE_LATENT | NOISE_LATE))) {
        Philox::gen(a.seed, (uint64_t)i, base + 0, r);
        float n0, n1, n2, n3;
        box_muller(r[0], r[1], n0, n1);
        box_muller(r[2], r[3], n2, n3);
        if (a.eps && (part & NOISE_LATENT)) a.eps[i] = n0;
        if (a.zp_full && (part & NOISE_LATE)) a.zp_full[i] = n1;
        if (a.zp_rf && (part & NOISE_LATE)) a.zp_rf[i] = n2;
    }
    if (i < a.B && a.c && (part & NOISE_LATENT)) {
        Philox::gen(a.seed, (uint64_t)i, base + 1, r);
        int k = r[0] >> 31;                                 // Cat([.5, .5]) one-hot (model.py:121-126)
        a.c[i * CD + 0] = k == 0 ? 1.f : 0.f;
        a.c[i * CD + 1] = k == 1 ? 1.f : 0.f;
    }
    if (i < n_tok && a.word_drop && (part & NOISE_WORD)) {
        Philox::gen(a.seed, (uint64_t)i, base + 2, r);
        a.word_drop[i] = u32_to_unit_open(r[0]) < a.p_word ? 1 : 0;      // decoder.py:124-127
    }
    if (i < n_out4 && a.out_keep && (part & NOISE_LATE)) {
        Philox::gen(a.seed, (uint64_t)i, base + 3, r);
        const int64_t n_out = (int64_t)a.B * a.L * DEC_H;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int64_t j = i * 4 + q;
            if (j < n_out) a.out_keep[j] = u32_to_unit_open(r[q]) >= a.p_out ? 1 : 0;   // keep w.p. 1-p
        }
    }
}

// out[i] = normal (kind 0) or uniform[0,1) * scale (kind 1)
__global__ void k_fill_random(uint64_t seed, uint32_t stream, int kind, float scale, int64_t n, float* __restrict__ out) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // handles 4 outputs
    if (q * 4 >= n) return;
    uint32_t r[4];
    Philox::gen(seed, (uint64_t)q, stream, r);
    float v[4];
    if (kind == 0) {
        box_muller(r[0], r[1], v[0], v[1]);
        box_muller(r[2], r[3], v[2], v[3]);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (float)(r[k] >> 8) * (1.0f / 16777216.0f) * scale;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (q * 4 + k < n) out[q * 4 + k] = v[k];
}

void launch_step_noise(cudaStream_t s, const StepNoiseArgs& a, int part) {
    int64_t n = 0;
    if (part & (NOISE_LATENT | NOISE_LATE)) n = (int64_t)a.B * ZD;         // normals
    if (part & NOISE_LATENT) n = std::max<int64_t>(n, a.B);
    if (part & NOISE_WORD) n = std::max<int64_t>(n, (int64_t)a.B * a.L);
    if (part & NOISE_LATE) n = std::max<int64_t>(n, ((int64_t)a.B * a.L * DEC_H + 3) / 4);
    CPG_LAUNCH(k_step_noise, (unsigned)((n + 255) / 256), 256, 0, s, a, part, g_dyn);
}

}  // namespace cpg

using namespace cpg;

extern "C" {

int cpg_fill_step_noise(cpg_ctx* ctx, cpg_stream stream, uint64_t seed, uint32_t step, int B, int L, float p_word,
                        float p_out, float* eps, float* c, uint8_t* word_drop, uint8_t* out_keep,
                        float* z_prior_full, float* z_prior_rf) {
    if (!ctx || B < 1 || L < 1) { set_error("cpg_fill_step_noise: bad argument"); return CPG_EINVAL; }
    StepNoiseArgs a;
    a.seed = seed; a.step = step; a.B = B; a.L = L; a.p_word = p_word; a.p_out = p_out;
    a.eps = eps; a.c = c; a.word_drop = word_drop; a.out_keep = out_keep; a.zp_full = z_prior_full; a.zp_rf = z_prior_rf;
    if (ctx->gen_deferred) {                       // an unconsumed request of cpg_fill_step_noise_overlapped comes first
        ctx->gen_deferred = false;
        launch_step_noise((cudaStream_t)stream, ctx->gen_args, NOISE_ALL);
    }
    launch_step_noise((cudaStream_t)stream, a, NOISE_ALL);
    return check_launch("cpg_fill_step_noise");
}

int cpg_fill_normal(cpg_ctx* ctx, cpg_stream stream, uint64_t seed, uint32_t stream_id, int64_t n, float* out) {
    if (!ctx || !out || n < 1) { set_error("cpg_fill_normal: bad argument"); return CPG_EINVAL; }
    CPG_LAUNCH(k_fill_random, (unsigned)(((n + 3) / 4 + 255) / 256), 256, 0, (cudaStream_t)stream, seed, stream_id, 0,
               1.0f, n, out);
    return check_launch("cpg_fill_normal");
}

int cpg_fill_uniform(cpg_ctx* ctx, cpg_stream stream, uint64_t seed, uint32_t stream_id, float scale, int64_t n,
                     float* out) {
    if (!ctx || !out || n < 1) { set_error("cpg_fill_uniform: bad argument"); return CPG_EINVAL; }
    CPG_LAUNCH(k_fill_random, (unsigned)(((n + 3) / 4 + 255) / 256), 256, 0, (cudaStream_t)stream, seed, stream_id, 1,
               scale, n, out);
    return check_launch("cpg_fill_uniform");
}

}  // extern "C"
