// Decoder output layer: out-dropout -> Linear(102 -> V) -> softmax cross-entropy,
// forward and backward fused, one warp per (sample, step) row and one lane per
// vocabulary class (V <= 32).
//
// Replaces nn.Dropout + nn.Linear at models/decoder.py:43-45,83 and
// F.cross_entropy(..., reduction='mean', ignore_index=PAD) at losses.py:27-30:
//   nll_row = logsumexp(logits) - logits[tgt]   (0 where tgt == <pad>)
//   dlogits = (softmax - onehot(tgt)) / N_tok    (0 where tgt == <pad>)
// with N_tok the number of non-<pad> targets of the WHOLE (global) batch.
#include "kernels.h"
#include "dec_out.h"

namespace cpg {

constexpr int DO_WARPS = 4;
constexpr int WSTR = DEC_HP + 1;      // odd stride: lane v reading W[v][j] is conflict free

__global__ void __launch_bounds__(DO_WARPS * 32)
k_dec_out(DecOutArgs a) {
    __shared__ float Ws[VMAX * WSTR];
    __shared__ float hd_s[DO_WARPS][DEC_HP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int V = a.V;
    for (int i = threadIdx.x; i < VMAX * DEC_HP; i += blockDim.x) {
        int v = i / DEC_HP, j = i % DEC_HP;
        Ws[v * WSTR + j] = a.fc_w[i];
    }
    __syncthreads();
    const float bias = (lane < V) ? a.fc_b[lane] : 0.f;
    const float inv_ntok = (a.ntok != nullptr && *a.ntok > 0.f) ? 1.0f / *a.ntok : 0.f;
    const bool want_grad = a.dh_out != nullptr;

    float accW[VMAX][4];
    float accb = 0.f, accn = 0.f;
#pragma unroll
    for (int v = 0; v < VMAX; ++v)
#pragma unroll
        for (int m = 0; m < 4; ++m) accW[v][m] = 0.f;

    const int nrows = a.B * a.L;
    const int gw = blockIdx.x * DO_WARPS + warp, nw = gridDim.x * DO_WARPS;
    for (int row = gw; row < nrows; row += nw) {
        float hd[4], ks[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            int j = lane + 32 * m;
            float h = 0.f, k = 1.f;
            if (j < DEC_HP) h = a.hs[(size_t)row * DEC_HP + j];
            if (a.out_keep != nullptr) k = (j < DEC_H) ? (a.out_keep[(size_t)row * DEC_H + j] ? a.keep_scale : 0.f) : 0.f;
            hd[m] = h * k;
            ks[m] = k;
            if (j < DEC_HP) hd_s[warp][j] = hd[m];
        }
        __syncwarp();
        float logit = -INFINITY;
        if (lane < V) {
            float s = 0.f;
            const float* w = Ws + lane * WSTR;
#pragma unroll 8
            for (int j = 0; j < DEC_H; ++j) s = fmaf(hd_s[warp][j], w[j], s);
            logit = s + bias;
            if (a.logits_out != nullptr) a.logits_out[(size_t)row * V + lane] = logit;
        }
        float dl = 0.f;
        if (a.fused_ce) {
            const float mx = warp_max(logit);
            const float e = (lane < V) ? expf(logit - mx) : 0.f;
            const float se = warp_sum(e);
            const int tg = a.tgt[row];
            const float lt = __shfl_sync(0xffffffffu, logit, tg);
            if (tg != PAD) {
                accn += (mx + logf(se)) - lt;                    // every lane holds the same value
                dl = (e / se - (lane == tg ? 1.f : 0.f)) * inv_ntok;
            }
        } else if (a.dlogits_in != nullptr && lane < V) {
            dl = a.dlogits_in[(size_t)row * V + lane];
        }
        if (want_grad) {
            float dhd[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int v = 0; v < VMAX; ++v) {
                if (v < V) {
                    const float dv = __shfl_sync(0xffffffffu, dl, v);
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        int j = lane + 32 * m;
                        float w = (j < DEC_HP) ? Ws[v * WSTR + j] : 0.f;
                        dhd[m] = fmaf(dv, w, dhd[m]);
                        accW[v][m] = fmaf(dv, hd[m], accW[v][m]);
                    }
                }
            }
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                int j = lane + 32 * m;
                if (j < DEC_HP) a.dh_out[(size_t)row * DEC_HP + j] = dhd[m] * ks[m];
            }
            accb += dl;
        }
        __syncwarp();
    }
    if (want_grad) {
        float* pw = a.part_w + (size_t)gw * VMAX * DEC_HP;
#pragma unroll
        for (int v = 0; v < VMAX; ++v)
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                int j = lane + 32 * m;
                if (j < DEC_HP) pw[v * DEC_HP + j] = accW[v][m];
            }
        a.part_b[(size_t)gw * VMAX + lane] = accb;
    }
    if (a.part_nll != nullptr && lane == 0) a.part_nll[gw] = accn;
}

// ordered reduction of the per-warp partials into dW_fc [V][102], db_fc [V], sum nll
__global__ void k_dec_out_reduce(const float* __restrict__ part_w, const float* __restrict__ part_b,
                                 const float* __restrict__ part_nll, int nparts, int V,
                                 float* __restrict__ dW, float* __restrict__ db, float* __restrict__ nll_sum) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < V * DEC_H) {
        int v = i / DEC_H, j = i % DEC_H;
        float s = 0.f;
        for (int p = 0; p < nparts; ++p) s += part_w[((size_t)p * VMAX + v) * DEC_HP + j];
        dW[i] = s;
    } else if (i < V * DEC_H + V) {
        int v = i - V * DEC_H;
        float s = 0.f;
        for (int p = 0; p < nparts; ++p) s += part_b[(size_t)p * VMAX + v];
        db[v] = s;
    } else if (i == V * DEC_H + V && nll_sum != nullptr) {
        double s = 0.0;
        for (int p = 0; p < nparts; ++p) s += (double)part_nll[p];
        *nll_sum = (float)s;
    }
}

int dec_out_parts(int B, int L, int sm_count) {
    int rows = B * L;
    int ctas = min(ceil_div(rows, DO_WARPS), max(1, sm_count));
    return ctas * DO_WARPS;
}

void launch_dec_out(cudaStream_t s, const DecOutArgs& a, int sm_count) {
    int parts = dec_out_parts(a.B, a.L, sm_count);
    CPG_LAUNCH(k_dec_out, parts / DO_WARPS, DO_WARPS * 32, 0, s, a);
}

void launch_dec_out_reduce(cudaStream_t s, const DecOutArgs& a, int sm_count, float* dW, float* db, float* nll_sum) {
    int parts = dec_out_parts(a.B, a.L, sm_count);
    int n = a.V * DEC_H + a.V + 1;
    CPG_LAUNCH(k_dec_out_reduce, ceil_div(n, 128), 128, 0, s, a.part_w, a.part_b, a.part_nll, parts, a.V, dW, db, nll_sum);
}

}  // namespace cpg
