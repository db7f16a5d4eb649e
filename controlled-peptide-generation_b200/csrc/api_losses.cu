// C ABI: the individual loss functions of losses.py (module-level API), built from the same
// kernels as the fused step.
#include <string.h>
#include "ctx.h"

namespace cpg {
int check_launch(const char* where);

// recon_dec on caller-provided logits (losses.py:18-31): one warp per (b,t) row, lane per class.
__global__ void k_xent_count(const int64_t* __restrict__ tokens, int B, int L, int* __restrict__ ntok) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int cnt = 0;
    if (i < B * L) {
        int t = i % L;
        int64_t nx = (t + 1 < L) ? tokens[i + 1] : (int64_t)PAD;
        cnt = nx != PAD;
    }
    unsigned m = __ballot_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(ntok, __popc(m));
}

constexpr int XE_WARPS = 8;
__global__ void __launch_bounds__(XE_WARPS * 32)
k_xent_rows(const float* __restrict__ logits, const int64_t* __restrict__ tokens, int B, int L, int V,
            const int* __restrict__ ntok, float* __restrict__ dlogits, float* __restrict__ part) {
    __shared__ float red[XE_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = *ntok;
    const float inv = n > 0 ? 1.0f / (float)n : 0.f;
    float acc = 0.f;
    for (int row = blockIdx.x * XE_WARPS + warp; row < B * L; row += gridDim.x * XE_WARPS) {
        int t = row % L;
        int tg = (t + 1 < L) ? (int)tokens[row + 1] : PAD;
        float lg = lane < V ? logits[(size_t)row * V + lane] : -INFINITY;
        float mx = warp_max(lg);
        float e = lane < V ? expf(lg - mx) : 0.f;
        float se = warp_sum(e);
        float lt = __shfl_sync(0xffffffffu, lg, tg & 31);
        float dl = 0.f;
        if (tg != PAD) {
            acc += (mx + logf(se)) - lt;
            dl = (e / se - (lane == tg ? 1.f : 0.f)) * inv;
        }
        if (dlogits != nullptr && lane < V) dlogits[(size_t)row * V + lane] = dl;
    }
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < XE_WARPS; ++w) s += red[w];
        part[blockIdx.x] = s;
    }
}
__global__ void k_xent_final(const float* __restrict__ part, int nparts, const int* __restrict__ ntok,
                             float* __restrict__ out) {
    if (threadIdx.x != 0) return;
    double s = 0.0;
    for (int i = 0; i < nparts; ++i) s += (double)part[i];
    int n = *ntok;
    out[0] = n > 0 ? (float)(s / (double)n) : 0.f;
    out[1] = (float)n;
}

// means as the reference reports them: kl, kl_sharedmu, logvar L1 (mean over batch), mean|mu|, mean logvar
__global__ void k_latent_means(const float* __restrict__ sums, int B, float* __restrict__ out) {
    if (threadIdx.x == 0) {
        float invB = 1.0f / (float)B;
        out[0] = sums[0] * invB; out[1] = sums[1] * invB; out[2] = sums[2] * invB;
        out[3] = sums[3] * invB / (float)ZD; out[4] = sums[4] * invB / (float)ZD;
    }
}

}  // namespace cpg

using namespace cpg;

extern "C" {

int cpg_softmax_xent(cpg_ctx* ctx, cpg_stream stream, const float* logits, const int64_t* tokens, int B, int L, int V,
                     float* loss_out, float* d_logits) {
    if (!ctx || !logits || !tokens || !loss_out) { set_error("cpg_softmax_xent: null argument"); return CPG_EINVAL; }
    if (V < 2 || V > VMAX || B < 1 || L < 2) { set_error("cpg_softmax_xent: bad shape"); return CPG_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    if (ctx->base == nullptr && (rc = ensure_workspace(ctx, B, L > LMAX ? LMAX : L, V < 4 ? 4 : V, 500, s))) return rc;
    int nparts = std::max(1, std::min(ceil_div(B * L, XE_WARPS), 2 * ctx->sm_count));
    float* part = ctx->ws.norm_part;          // 2*sm+8 floats
    cudaMemsetAsync(ctx->ints + 2, 0, sizeof(int), s);
    CPG_LAUNCH(k_xent_count, ceil_div(B * L, 256), 256, 0, s, tokens, B, L, ctx->ints + 2);
    CPG_LAUNCH(k_xent_rows, nparts, XE_WARPS * 32, 0, s, logits, tokens, B, L, V, ctx->ints + 2, d_logits, part);
    CPG_LAUNCH(k_xent_final, 1, 32, 0, s, part, nparts, ctx->ints + 2, loss_out);
    return check_launch("cpg_softmax_xent");
}

int cpg_latent_stats(cpg_ctx* ctx, cpg_stream stream, const float* mu, const float* logvar, int B, float* out5) {
    if (!ctx || !mu || !logvar || !out5 || B < 1) { set_error("cpg_latent_stats: bad argument"); return CPG_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    if (ctx->base == nullptr && (rc = ensure_workspace(ctx, B, 2, 4, 500, s))) return rc;
    // partials live in norm_part (2*sm+8 floats): use at most (2*sm)/5 CTAs
    int nparts = std::max(1, std::min(ceil_div(B, 8), (2 * ctx->sm_count) / 5));
    launch_latent_stats(s, mu, logvar, B, ctx->ws.norm_part, nparts, ctx->ws.lat_sums);
    CPG_LAUNCH(k_latent_means, 1, 32, 0, s, ctx->ws.lat_sums, B, out5);
    return check_launch("cpg_latent_stats");
}

int cpg_mmd_full(cpg_ctx* ctx, cpg_stream stream, const float* z, const float* zp, int B, float sigma, float* out) {
    if (!ctx || !z || !zp || !out || B < 2) { set_error("cpg_mmd_full: bad argument"); return CPG_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    if ((rc = ensure_aux(ctx, mmd_ws_floats(B) * sizeof(float), s))) return rc;
    if ((rc = launch_mmd_full(s, z, zp, B, sigma, (float*)ctx->aux, out))) return rc;
    return check_launch("cpg_mmd_full");
}

int cpg_mmd_full_grad(cpg_ctx* ctx, cpg_stream stream, const float* z, const float* zp, int B, float sigma, float* dz) {
    if (!ctx || !z || !zp || !dz || B < 2) { set_error("cpg_mmd_full_grad: bad argument"); return CPG_EINVAL; }
    launch_mmd_full_grad((cudaStream_t)stream, z, zp, B, sigma, 1.0f, dz);
    return check_launch("cpg_mmd_full_grad");
}

int cpg_mmd_rf(cpg_ctx* ctx, cpg_stream stream, const float* z, const float* zp, const float* rf_w, const float* rf_b,
               int B, int R, float sigma, float* loss_out, float* dz) {
    if (!ctx || !z || !zp || !rf_w || !rf_b || !loss_out || B < 1 || R < 1) { set_error("cpg_mmd_rf: bad argument"); return CPG_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    // own scratch (never the step workspace: a loss evaluated between forward and backward must not touch the stash)
    const int nchunk = std::max(1, std::min(B, 2 * ctx->sm_count));
    const size_t nB = align_up((size_t)B * R, 64), nC = align_up((size_t)nchunk * R, 64), nR = align_up((size_t)R, 64);
    if ((rc = ensure_aux(ctx, (2 * nB + nC + 3 * nR) * sizeof(float), s))) return rc;
    float* pre1 = (float*)ctx->aux;
    float* pre2 = pre1 + nB;
    float* part = pre2 + nB;
    float* sum1 = part + nC;
    float* sum2 = sum1 + nR;
    float* coef = sum2 + nR;
    launch_sgemm(s, B, R, ZD, 1.f, z, ZD, 1, rf_w, R, 1, 0.f, pre1, R, nullptr, 1, nullptr);
    launch_sgemm(s, B, R, ZD, 1.f, zp, ZD, 1, rf_w, R, 1, 0.f, pre2, R, nullptr, 1, nullptr);
    launch_rf_colsum(s, pre1, rf_b, B, R, sigma, part, nchunk, sum1);
    launch_rf_colsum(s, pre2, rf_b, B, R, sigma, part, nchunk, sum2);
    launch_rf_loss(s, sum1, sum2, R, B, sigma, 1.0f, coef, loss_out);
    if (dz != nullptr) {
        launch_rf_grad_prep(s, pre1, rf_b, coef, B, R, sigma);
        launch_sgemm(s, B, ZD, R, 1.f, pre1, R, 1, rf_w, 1, R, 0.f, dz, ZD, nullptr, 1, nullptr);
    }
    return check_launch("cpg_mmd_rf");
}

int cpg_set_option(const char* name, int value) {
    if (name == nullptr) { set_error("cpg_set_option: null name"); return CPG_EINVAL; }
    if (strcmp(name, "mmd_tensor_core") == 0) { g_opt_mmd_tc = value; return CPG_OK; }
    if (strcmp(name, "wgrad_tensor_core") == 0) { g_opt_wgrad_tc = value; return CPG_OK; }
    if (strcmp(name, "gru_tensor_core") == 0) { g_opt_gru_tc = value; return CPG_OK; }
    if (strcmp(name, "dec_out_tensor_core") == 0) { g_opt_dec_out_tc = value; return CPG_OK; }
    if (strcmp(name, "side_stream") == 0) { g_opt_side_stream = value; return CPG_OK; }
    if (strcmp(name, "bptt_fused") == 0) { g_opt_bptt_fused = value; return CPG_OK; }
    if (strcmp(name, "cuda_graph") == 0) { g_opt_graph = value; return CPG_OK; }
    if (strcmp(name, "latent_tensor_core") == 0) { g_opt_latent_tc = value; return CPG_OK; }
    if (strcmp(name, "rf_tensor_core") == 0) { g_opt_rf_tc = value; return CPG_OK; }
    if (strcmp(name, "wgrad_dense_tensor_core") == 0) { g_opt_wgrad_dense_tc = value; return CPG_OK; }
    if (strcmp(name, "matmul_terms") == 0) {
        if (value != 1 && value != 3) { set_error("matmul_terms must be 3 (split-bf16, fp32-grade) or 1 (single bf16 product)"); return CPG_EINVAL; }
        g_opt_matmul_terms = value;
        return CPG_OK;
    }
    if (strcmp(name, "rf_grid") == 0) { g_opt_rf_grid = value; return CPG_OK; }
    if (strcmp(name, "wgrad_dense_grid") == 0) { g_opt_wd_grid = value < 64 ? value : 64; return CPG_OK; }
    if (strcmp(name, "mmd_grid") == 0) { g_opt_mmd_grid = value; return CPG_OK; }
    if (strcmp(name, "adam_fused") == 0) { g_opt_adam_fused = value; return CPG_OK; }
    if (strcmp(name, "chain_priority") == 0) { g_opt_chain_priority = value; return CPG_OK; }
    if (strcmp(name, "latent_tile_rows") == 0) {
        if (value != 64 && value != 128) { set_error("latent_tile_rows must be 64 or 128"); return CPG_EINVAL; }
        g_opt_latent_rows = value;
        return CPG_OK;
    }
    set_error(std::string("cpg_set_option: unknown option ") + name);
    return CPG_EINVAL;
}

}  // extern "C"
