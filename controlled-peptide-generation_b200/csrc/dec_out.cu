// Decoder output layer: out-dropout -> Linear(102 -> V) -> softmax cross-entropy, forward and
// backward fused over 64-row tiles of the [B*L, 104] hidden-state matrix.
//
// Replaces nn.Dropout + nn.Linear at models/decoder.py:43-45,83 and
// F.cross_entropy(..., reduction='mean', ignore_index=PAD) at losses.py:27-30:
//   nll_row = logsumexp(logits) - logits[tgt]   (0 where tgt == <pad>)
//   dlogits = (softmax - onehot(tgt)) / N_tok    (0 where tgt == <pad>)
// with N_tok the number of non-<pad> targets of the WHOLE (global) batch.
//
// Per tile, three small contractions run out of shared memory with float4 operands:
//   logits = hd W^T (64x32x104),  dhd = dlogits W (64x104x32),  dW += dlogits^T hd (32x104x64)
// dW / db / nll accumulate in registers across the tiles of a persistent CTA and leave as one
// partial per CTA (ordered reduction afterwards, no float atomics).
#include "kernels.h"
#include "dec_out.h"
#include "../../include/cpg_b200.h"

namespace cpg {

constexpr int DT = 64;                 // rows per tile
constexpr int DSTR = 108;              // padded row stride (floats) of the hd and W tiles
constexpr int LSTR = 36;               // padded row stride of the dlogits tile
constexpr int DO_THREADS = 256;
constexpr int NF4 = DEC_HP / 4;        // 26 float4 per hidden row

__global__ void __launch_bounds__(DO_THREADS, 2)
k_dec_out(DecOutArgs a) {
    CPG_DYN_SMEM(float, smem);                              // 57 KB: above the 48 KB static limit
    float* Ws = smem;                                       // [VMAX][DSTR]
    float* hd_s = Ws + VMAX * DSTR;                         // [DT][DSTR]
    float* dl_s = hd_s + DT * DSTR;                         // [DT][LSTR]
    unsigned char* keep_s = reinterpret_cast<unsigned char*>(dl_s + DT * LSTR);   // [DT][DEC_HP]
    __shared__ float red_nll[DO_THREADS / 32];
    const int tid = threadIdx.x;
    const int V = a.V;
    const int nrows = a.B * a.L;
    const bool want_grad = a.dh_out != nullptr;
    const bool has_mask = a.out_keep != nullptr;
    const float ntok_v = a.ntok_i != nullptr ? (float)*a.ntok_i : (a.ntok != nullptr ? *a.ntok : 0.f);
    const float inv_ntok = ntok_v > 0.f ? 1.0f / ntok_v : 0.f;

    for (int i = tid; i < VMAX * DEC_HP; i += DO_THREADS) Ws[(i / DEC_HP) * DSTR + (i % DEC_HP)] = a.fc_w[i];
    // phase B / C mapping: 2 rows x 4 classes per thread
    const int rp = tid >> 3, vq = tid & 7;
    float bias4[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) bias4[c] = (vq * 4 + c < V) ? a.fc_b[vq * 4 + c] : 0.f;
    // phase E mapping: class v, float4 columns jq + 8 i
    const int ev = tid >> 3, ejq = tid & 7;
    float4 accW[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) accW[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    float accb = 0.f, accn = 0.f;

    const int ntiles = ceil_div(nrows, DT);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int row0 = tile * DT;
        __syncthreads();                                   // previous tile fully consumed (and Ws visible)
        // ---- A) stage hd = h * keep * scale (and the keep bytes) for 64 rows
        for (int idx = tid; idx < DT * NF4; idx += DO_THREADS) {
            const int r = idx / NF4, f = idx % NF4;
            const int row = row0 + r;
            float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
            unsigned char k4[4] = {0, 0, 0, 0};
            if (row < nrows) {
                h = ld4(a.hs + (size_t)row * DEC_HP + f * 4);
                float hv[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int j = f * 4 + c;
                    unsigned char k = j < DEC_H ? (has_mask ? a.out_keep[(size_t)row * DEC_H + j] : 1) : 0;
                    k4[c] = k;
                    hv[c] = k ? (has_mask ? hv[c] * a.keep_scale : hv[c]) : 0.f;
                }
                h = make_float4(hv[0], hv[1], hv[2], hv[3]);
            }
            st4(hd_s + r * DSTR + f * 4, h);
#pragma unroll
            for (int c = 0; c < 4; ++c) keep_s[r * DEC_HP + f * 4 + c] = k4[c];
        }
        __syncthreads();
        // ---- B) logits for rows (2 rp, 2 rp + 1), classes 4 vq .. 4 vq + 3
        float lg[2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int c = 0; c < 4; ++c) lg[i][c] = 0.f;
#pragma unroll 2
        for (int f = 0; f < NF4; ++f) {
            const float4 h0 = ld4(hd_s + (2 * rp) * DSTR + f * 4);
            const float4 h1 = ld4(hd_s + (2 * rp + 1) * DSTR + f * 4);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float4 w = ld4(Ws + (vq * 4 + c) * DSTR + f * 4);
                lg[0][c] = fmaf(h0.x, w.x, fmaf(h0.y, w.y, fmaf(h0.z, w.z, fmaf(h0.w, w.w, lg[0][c]))));
                lg[1][c] = fmaf(h1.x, w.x, fmaf(h1.y, w.y, fmaf(h1.z, w.z, fmaf(h1.w, w.w, lg[1][c]))));
            }
        }
        // ---- C) softmax / CE per row (8 consecutive lanes hold the 32 classes of a row)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = 2 * rp + i, row = row0 + r;
            float dl[4] = {0.f, 0.f, 0.f, 0.f};
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                lg[i][c] = (vq * 4 + c < V) ? lg[i][c] + bias4[c] : -INFINITY;
                mx = fmaxf(mx, lg[i][c]);
            }
            if (row < nrows && a.logits_out != nullptr) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (vq * 4 + c < V) a.logits_out[(size_t)row * V + vq * 4 + c] = lg[i][c];
            }
            if (a.fused_ce) {
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
                float e[4], se = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) { e[c] = (vq * 4 + c < V) ? expf(lg[i][c] - mx) : 0.f; se += e[c]; }
                se += __shfl_xor_sync(0xffffffffu, se, 1);
                se += __shfl_xor_sync(0xffffffffu, se, 2);
                se += __shfl_xor_sync(0xffffffffu, se, 4);
                const int tg = row < nrows ? a.tgt[row] : PAD;
                float lt = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) if (vq * 4 + c == tg) lt = lg[i][c];
                lt += __shfl_xor_sync(0xffffffffu, lt, 1);
                lt += __shfl_xor_sync(0xffffffffu, lt, 2);
                lt += __shfl_xor_sync(0xffffffffu, lt, 4);
                if (tg != PAD) {
                    if (vq == 0) accn += (mx + logf(se)) - lt;
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        dl[c] = (vq * 4 + c < V) ? (e[c] / se - (vq * 4 + c == tg ? 1.f : 0.f)) * inv_ntok : 0.f;
                }
            } else if (a.dlogits_in != nullptr && row < nrows) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (vq * 4 + c < V) dl[c] = a.dlogits_in[(size_t)row * V + vq * 4 + c];
            }
            st4(dl_s + r * LSTR + vq * 4, make_float4(dl[0], dl[1], dl[2], dl[3]));
        }
        if (!want_grad) continue;
        __syncthreads();
        // ---- D) dh = (dlogits W) * keep * scale : row r = tid / 4, float4 columns jq + 4 i
        {
            const int r = tid >> 2, jq = tid & 3, row = row0 + r;
            float4 acc[7];
#pragma unroll
            for (int i = 0; i < 7; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int v = 0; v < V; ++v) {
                const float d = dl_s[r * LSTR + v];
#pragma unroll
                for (int i = 0; i < 7; ++i) {
                    const int f = jq + 4 * i;
                    if (f < NF4) {
                        const float4 w = ld4(Ws + v * DSTR + f * 4);
                        acc[i].x = fmaf(d, w.x, acc[i].x); acc[i].y = fmaf(d, w.y, acc[i].y);
                        acc[i].z = fmaf(d, w.z, acc[i].z); acc[i].w = fmaf(d, w.w, acc[i].w);
                    }
                }
            }
            if (row < nrows) {
                const float sc = has_mask ? a.keep_scale : 1.0f;
#pragma unroll
                for (int i = 0; i < 7; ++i) {
                    const int f = jq + 4 * i;
                    if (f < NF4) {
                        const unsigned char* k = keep_s + r * DEC_HP + f * 4;
                        st4(a.dh_out + (size_t)row * DEC_HP + f * 4,
                            make_float4(k[0] ? acc[i].x * sc : 0.f, k[1] ? acc[i].y * sc : 0.f,
                                        k[2] ? acc[i].z * sc : 0.f, k[3] ? acc[i].w * sc : 0.f));
                    }
                }
            }
        }
        // ---- E) dW[v][:] += sum_r dlogits[r][v] hd[r][:],  db[v] += sum_r dlogits[r][v]
        for (int r = 0; r < DT; ++r) {
            const float d = dl_s[r * LSTR + ev];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int f = ejq + 8 * i;
                if (f < NF4) {
                    const float4 h = ld4(hd_s + r * DSTR + f * 4);
                    accW[i].x = fmaf(d, h.x, accW[i].x); accW[i].y = fmaf(d, h.y, accW[i].y);
                    accW[i].z = fmaf(d, h.z, accW[i].z); accW[i].w = fmaf(d, h.w, accW[i].w);
                }
            }
            if (ejq == 0) accb += d;
        }
    }
    // ---- per-CTA partials
    if (want_grad) {
        float* pw = a.part_w + (size_t)blockIdx.x * VMAX * DEC_HP;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int f = ejq + 8 * i;
            if (f < NF4) st4(pw + ev * DEC_HP + f * 4, accW[i]);
        }
        if (ejq == 0) a.part_b[(size_t)blockIdx.x * VMAX + ev] = accb;
    }
    if (a.part_nll != nullptr) {
        accn = warp_sum(accn);
        if ((tid & 31) == 0) red_nll[tid >> 5] = accn;
        __syncthreads();
        if (tid == 0) {
            float s = 0.f;
            for (int w = 0; w < DO_THREADS / 32; ++w) s += red_nll[w];
            a.part_nll[blockIdx.x] = s;
        }
    }
}

// ordered reduction of the per-CTA partials into dW_fc [V][102], db_fc [V], sum nll
__global__ void k_dec_out_reduce(const float* __restrict__ part_w, const float* __restrict__ part_b,
                                 const float* __restrict__ part_nll, int nparts, int V,
                                 float* __restrict__ dW, float* __restrict__ db, float* __restrict__ nll_sum) {
    const int i = blockIdx.x * RED_X + threadIdx.x;
    const int nW = V * DEC_H;
    const bool isW = i < nW, isB = !isW && i - nW < V;
    const float* src = isW ? part_w : part_b;
    const size_t stride = isW ? (size_t)VMAX * DEC_HP : (size_t)VMAX;
    const size_t off = isW ? (size_t)(i / DEC_H) * DEC_HP + i % DEC_H : (size_t)(isB ? i - nW : 0);
    const float s = block_split_sum(src, stride, nparts, off, isW || isB);      // one uniform call per block
    if (threadIdx.y != 0) return;
    if (isW) dW[i] = s;
    else if (isB) db[i - nW] = s;
    else if (i - nW == V && nll_sum != nullptr) {
        double t = 0.0;
        for (int p = 0; p < nparts; ++p) t += (double)part_nll[p];
        *nll_sum = (float)t;
    }
}

int dec_out_parts(int B, int L, int sm_count) {
    int tiles = ceil_div(B * L, DT);
    return max(1, min(tiles, 2 * max(1, sm_count)));
}

int g_opt_dec_out_tc = 1;
static int g_last_parts = 0;       // partials written by the most recent launch_dec_out (consumed by the reduce that follows it)

void launch_dec_out(cudaStream_t s, const DecOutArgs& a, int sm_count) {
#ifndef CPG_EMU
    if (g_opt_dec_out_tc == 2 || (g_opt_dec_out_tc == 1 && a.B * a.L >= 8192)) {
        if (launch_dec_out_tc(s, a, sm_count, &g_last_parts) == CPG_OK) return;
    }
#endif
    const size_t smem = (size_t)(VMAX * DSTR + DT * DSTR + DT * LSTR) * sizeof(float) + DT * DEC_HP;
    CPG_SET_MAX_SMEM(k_dec_out, smem);
    g_last_parts = dec_out_parts(a.B, a.L, sm_count);
    CPG_LAUNCH(k_dec_out, g_last_parts, DO_THREADS, smem, s, a);
}

void launch_dec_out_reduce(cudaStream_t s, const DecOutArgs& a, int sm_count, float* dW, float* db, float* nll_sum) {
    int parts = g_last_parts > 0 ? g_last_parts : dec_out_parts(a.B, a.L, sm_count);
    int n = a.V * DEC_H + a.V + 1;
    CPG_LAUNCH(k_dec_out_reduce, CPG_RED_GRID(n), CPG_RED_BLOCK, 0, s, a.part_w, a.part_b, a.part_nll, parts, a.V, dW, db, nll_sum);
}

}  // namespace cpg
