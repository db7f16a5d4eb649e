// Normalising-flow transforms of the latent code (models/flow.py:30-160 of the reference: PlanarFlow, RadialFlow,
// AlternatingFlow), forward pass with the log-determinant "loss" the reference returns in train mode.
// One warp per latent row, lanes over the 100 dimensions; all layers of the flow in one launch.
//
//   planar:  a = z.w + b;  z' = z + s tanh(a);  det = 1 + (1 - tanh(a)^2) (w.s)
//   radial:  r = z - z0;  act = 1 / (alpha + |r|);  z' = z + beta act r;
//            det = (1 + beta act)^(D-1) (1 + beta act - beta act^2 |R|_F)   with |R|_F the norm of the WHOLE [B, D]
//            matrix of radii -- `radius.norm(2)` in flow.py:87 is not per row; reproduced as executed.
//   loss = mean_b sum_layers log(|det| + 1e-7)
// The radial determinant needs a batch-wide reduction per layer, so the row kernel stores (act, |r|) per radial layer
// and ordered partial sums of |r|^2; a second kernel finishes the loss.  Parameter maintenance ("maintain
// invertibility", flow.py:44-48,77-79) is scalar host logic and stays in models/flow.py of this package.
#include "ctx.h"

namespace cpg {
int check_launch(const char* where);

constexpr int FLOW_MAX = 16;
constexpr int FL_WARPS = 8;
struct FlowSpec {
    int n_layers, train;
    int kind[FLOW_MAX];              // 0 planar, 1 radial
    const float* va[FLOW_MAX];       // planar weight [D] | radial initial point z0 [D]
    const float* vb[FLOW_MAX];       // planar scale [D]  | unused
    float sa[FLOW_MAX];              // planar bias       | radial alpha
    float sb[FLOW_MAX];              // planar w.s        | radial beta
};

__global__ void __launch_bounds__(FL_WARPS * 32)
k_flow_rows(FlowSpec f, const float* __restrict__ z_in, int B, float* __restrict__ z_out, float* __restrict__ row_loss,
            float* __restrict__ rad_aux, float* __restrict__ rad_part) {
    __shared__ float sq_s[FL_WARPS][FLOW_MAX];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = blockIdx.x * FL_WARPS + warp;
    float z[4];
    float loss = 0.f;
    const bool live = row < B;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int d = lane + 32 * q;
        z[q] = (live && d < ZD) ? z_in[(size_t)row * ZD + d] : 0.f;
    }
    for (int l = 0; l < f.n_layers; ++l) {
        float sq = 0.f;
        if (f.kind[l] == 0) {
            float dot = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) { const int d = lane + 32 * q; if (d < ZD) dot = fmaf(z[q], f.va[l][d], dot); }
            const float t = tanhf(warp_sum(dot) + f.sa[l]);
#pragma unroll
            for (int q = 0; q < 4; ++q) { const int d = lane + 32 * q; if (d < ZD) z[q] = fmaf(f.vb[l][d], t, z[q]); }
            if (f.train) loss += logf(fabsf(1.0f + (1.0f - t * t) * f.sb[l]) + 1e-7f);
        } else {
            float r[4], s2 = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int d = lane + 32 * q;
                r[q] = d < ZD ? z[q] - f.va[l][d] : 0.f;
                s2 = fmaf(r[q], r[q], s2);
            }
            s2 = warp_sum(s2);
            const float rn = sqrtf(s2);
            const float act = 1.0f / (f.sa[l] + rn);
            const float g = f.sb[l] * act;
#pragma unroll
            for (int q = 0; q < 4; ++q) z[q] = fmaf(g, r[q], z[q]);
            if (f.train && live && lane == 0) {
                rad_aux[((size_t)row * f.n_layers + l) * 2 + 0] = act;
                rad_aux[((size_t)row * f.n_layers + l) * 2 + 1] = rn;
            }
            sq = live ? s2 : 0.f;
        }
        if (lane == 0) sq_s[warp][l] = sq;
    }
    if (live) {
#pragma unroll
        for (int q = 0; q < 4; ++q) { const int d = lane + 32 * q; if (d < ZD) z_out[(size_t)row * ZD + d] = z[q]; }
        if (f.train && lane == 0) row_loss[row] = loss;
    }
    if (f.train) {
        __syncthreads();
        if (threadIdx.x < f.n_layers) {                       // ordered partial of sum_rows |r|^2 per layer
            float s = 0.f;
            for (int w = 0; w < FL_WARPS; ++w) s += sq_s[w][threadIdx.x];
            rad_part[(size_t)blockIdx.x * FLOW_MAX + threadIdx.x] = s;
        }
    }
}

// one CTA: Frobenius norms per layer (ordered), then mean over rows of (planar part + radial part)
__global__ void __launch_bounds__(256)
k_flow_loss(FlowSpec f, int B, int nblocks, const float* __restrict__ row_loss, const float* __restrict__ rad_aux,
            const float* __restrict__ rad_part, float* __restrict__ loss_out) {
    __shared__ float fro[FLOW_MAX];
    __shared__ double red[256];
    if (threadIdx.x < FLOW_MAX) {
        double s = 0.0;
        if (threadIdx.x < f.n_layers)
            for (int b = 0; b < nblocks; ++b) s += (double)rad_part[(size_t)b * FLOW_MAX + threadIdx.x];
        fro[threadIdx.x] = (float)sqrt(s);
    }
    __syncthreads();
    double acc = 0.0;
    for (int row = threadIdx.x; row < B; row += 256) {
        float loss = row_loss[row];
        for (int l = 0; l < f.n_layers; ++l) {
            if (f.kind[l] != 1) continue;
            const float act = rad_aux[((size_t)row * f.n_layers + l) * 2];
            const float g = f.sb[l] * act;
            const float diag = powf(1.0f + g, (float)(ZD - 1));
            const float det = diag * (1.0f + g + f.sb[l] * (-act * act) * fro[l]);
            loss += logf(fabsf(det) + 1e-7f);
        }
        acc += (double)loss;
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < 256; ++i) s += red[i];
        *loss_out = (float)(s / (double)B);
    }
}

}  // namespace cpg

using namespace cpg;

extern "C" int cpg_flow_forward(cpg_ctx* ctx, cpg_stream stream, const float* z_in, int B, int n_layers, const int* kind,
                                const float* const* vec_a, const float* const* vec_b, const float* scalar_a,
                                const float* scalar_b, int train, float* z_out, float* loss_out) {
    if (!ctx || !z_in || !z_out || !kind || !vec_a || !vec_b || !scalar_a || !scalar_b || B < 1) { set_error("cpg_flow_forward: bad argument"); return CPG_EINVAL; }
    if (n_layers < 1 || n_layers > FLOW_MAX) { set_error("cpg_flow_forward: 1 <= flow layers <= 16"); return CPG_EINVAL; }
    if (train && !loss_out) { set_error("cpg_flow_forward: train mode needs loss_out"); return CPG_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    FlowSpec f;
    memset(&f, 0, sizeof(f));
    f.n_layers = n_layers; f.train = train ? 1 : 0;
    for (int l = 0; l < n_layers; ++l) {
        if (kind[l] != 0 && kind[l] != 1) { set_error("cpg_flow_forward: layer kind must be 0 (planar) or 1 (radial)"); return CPG_EINVAL; }
        if (!vec_a[l] || (kind[l] == 0 && !vec_b[l])) { set_error("cpg_flow_forward: null layer parameter"); return CPG_EINVAL; }
        f.kind[l] = kind[l]; f.va[l] = vec_a[l]; f.vb[l] = vec_b[l]; f.sa[l] = scalar_a[l]; f.sb[l] = scalar_b[l];
    }
    const int nblocks = ceil_div(B, FL_WARPS);
    float *row_loss = nullptr, *rad_aux = nullptr, *rad_part = nullptr;
    if (train) {
        const size_t n1 = align_up((size_t)B, 64), n2 = align_up((size_t)B * n_layers * 2, 64);
        int rc = ensure_aux(ctx, (n1 + n2 + (size_t)nblocks * FLOW_MAX) * sizeof(float), s);
        if (rc) return rc;
        row_loss = (float*)ctx->aux; rad_aux = row_loss + n1; rad_part = rad_aux + n2;
    }
    CPG_LAUNCH(k_flow_rows, nblocks, FL_WARPS * 32, 0, s, f, z_in, B, z_out, row_loss, rad_aux, rad_part);
    if (train) CPG_LAUNCH(k_flow_loss, 1, 256, 0, s, f, B, nblocks, row_loss, rad_aux, rad_part, loss_out);
    return check_launch("cpg_flow_forward");
}
