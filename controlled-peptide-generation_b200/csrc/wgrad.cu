// Weight gradients of the three GRUs from the (dr_pre, dz_pre, dn_pre, dhn) planes the BPTT
// kernels write (gru.cu):
//   dW_hh = sum_{b,s} [dr_pre, dz_pre, dhn]^T h_{s-1}       (contraction over B*L rows, split-K)
//   dT[v] = sum_{(b,s): token == v} [dr_pre, dz_pre, dn_pre, dhn]   (token-table gradient)
//   dW_ih[:, :150] = dT^T E,  dE = dT W_ih[:, :150],  db_ih = sum_v dT,  db_hh = (same r,z ; dhn for n)
// which is what autograd produces for nn.GRU + nn.Embedding in the reference
// (train_vae.py:40 loss.backward()), restructured around the token table.
#include "kernels.h"
#include "wgrad.h"

namespace cpg {

constexpr int WG_ROWS = 16;

template <int HP>
__global__ void __launch_bounds__((HP / 4) * (HP / 8))
k_wgrad_hh(const float* __restrict__ dg, const float* __restrict__ hs, const float* __restrict__ h0, int B, int L,
           int rows_per_split, float* __restrict__ part) {
    constexpr int NG = HP / 4, NK = HP / 8, NT = NG * NK;
    __shared__ __align__(16) float As[WG_ROWS][HP];
    __shared__ __align__(16) float Hs[WG_ROWS][HP];
    const int pl = blockIdx.x;                 // 0: r, 1: z, 2: n (uses the dhn plane)
    const int src_plane = pl == 2 ? 3 : pl;
    const int split = blockIdx.y;
    const int tid = threadIdx.x, gx = tid % NG, kx = tid / NG;
    const int nrows = B * L;
    const int rbeg = split * rows_per_split, rend = min(nrows, rbeg + rows_per_split);
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int r0 = rbeg; r0 < rend; r0 += WG_ROWS) {
        for (int idx = tid; idx < WG_ROWS * (HP / 4); idx += NT) {
            int rr = idx / (HP / 4), c4 = (idx % (HP / 4)) * 4;
            int row = r0 + rr;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), h = a;
            if (row < rend) {
                a = ld4(dg + ((size_t)row * 4 + src_plane) * HP + c4);
                int b = row / L, s = row % L;
                if (s > 0) h = ld4(hs + (size_t)(row - 1) * HP + c4);
                else if (h0 != nullptr) h = ld4(h0 + (size_t)b * HP + c4);
            }
            st4(&As[rr][c4], a);
            st4(&Hs[rr][c4], h);
        }
        __syncthreads();
#pragma unroll 4
        for (int rr = 0; rr < WG_ROWS; ++rr) {
            const float4 av = ld4(&As[rr][gx * 4]);
            const float4 h0v = ld4(&Hs[rr][kx * 8]);
            const float4 h1v = ld4(&Hs[rr][kx * 8 + 4]);
            const float a4[4] = {av.x, av.y, av.z, av.w};
            const float h8[8] = {h0v.x, h0v.y, h0v.z, h0v.w, h1v.x, h1v.y, h1v.z, h1v.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a4[i], h8[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* out = part + ((size_t)split * 3 * HP + pl * HP) * HP;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float* o = out + (size_t)(gx * 4 + i) * HP + kx * 8;
        st4(o, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
        st4(o + 4, make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]));
    }
}

// dW_hh [3H][H] (unpadded) = ordered sum of the split partials [split][3HP][HP]
__global__ void k_wgrad_hh_reduce(const float* __restrict__ part, int nsplit, int HP, int H, float* __restrict__ dW) {
    const int i = blockIdx.x * RED_X + threadIdx.x;
    const bool ok = i < 3 * H * H;
    const int g = ok ? i / H : 0, k = ok ? i % H : 0, pl = g / H, j = g % H;
    const float s = block_split_sum(part, (size_t)3 * HP * HP, nsplit, (size_t)(pl * HP + j) * HP + k, ok);
    if (ok && threadIdx.y == 0) dW[i] = s;
}

// both ordered reductions of a fused-BPTT partial set in ONE launch: blocks [0, nbw) the W_hh gradient, the rest the token table
__global__ void k_wgrad_partial_reduce(const float* __restrict__ part_w, const float* __restrict__ part_t, int nsplit, int HP, int H,
                                       int nbw, int nt, float* __restrict__ dW, float* __restrict__ dT) {
    if ((int)blockIdx.x < nbw) {
        const int i = blockIdx.x * RED_X + threadIdx.x;
        const bool ok = i < 3 * H * H;
        const int g = ok ? i / H : 0, k = ok ? i % H : 0, pl = g / H, j = g % H;
        const float s = block_split_sum(part_w, (size_t)3 * HP * HP, nsplit, (size_t)(pl * HP + j) * HP + k, ok);
        if (ok && threadIdx.y == 0) dW[i] = s;
    } else {
        const int i = (blockIdx.x - nbw) * RED_X + threadIdx.x;
        const bool ok = i < nt;
        const float s = block_split_sum(part_t, (size_t)nt, nsplit, (size_t)(ok ? i : 0), ok);
        if (ok && threadIdx.y == 0) dT[i] = s;
    }
}

int wgrad_splits(int B, int L, int sm_count) {
    int nrows = B * L;
    int want = max(1, sm_count / 3);
    int rps = ceil_div(ceil_div(nrows, want), WG_ROWS) * WG_ROWS;
    return ceil_div(nrows, rps);
}

int g_opt_wgrad_tc = 1;
#ifndef CPG_EMU
bool wgrad_uses_tc(int nrows) { return (g_opt_wgrad_tc == 1 && nrows >= 8192) || g_opt_wgrad_tc == 2; }
#else
bool wgrad_uses_tc(int) { return false; }
#endif
#ifdef CPG_EMU
int wgrad_tc_splits(int) { return 1; }      // the tcgen05 kernel is compiled out of the CPU emulation
#endif

__global__ void k_dtable_reduce(const float* __restrict__ part, int nsplit, int n, float* __restrict__ dT);

bool launch_wgrad_hh(cudaStream_t s, int HP, int H, const float* dg, const float* hs, const float* h0,
                     const uint8_t* tok, int reverse, int V, int B, int L, int sm_count, float* part, float* dt_part,
                     float* dW, float* dT, cudaStream_t reduce_stream, void* reduce_event, int dg_rounded) {
    int nrows = B * L;
#ifndef CPG_EMU
    // tf32 operands (round-to-nearest, fp32 accumulate): the rounding noise averages out over the
    // B*L-long reduction, so the tensor-core path is used where the reduction is long (>= 8192 rows);
    // short reductions (tiny batches) stay on the exact fp32 SIMT kernels.  g_opt_wgrad_tc = 2 forces it.
    if (wgrad_uses_tc(nrows)) {
        int nsplit = 0;
        if (launch_wgrad_tc(s, HP, dg, hs, h0, tok, reverse, B, L, V, sm_count, part, dt_part, &nsplit, dg_rounded) == 0) {
            cudaStream_t rs = s;
            if (reduce_stream != nullptr && reduce_event != nullptr) {
                cudaEventRecord((cudaEvent_t)reduce_event, s);
                cudaStreamWaitEvent(reduce_stream, (cudaEvent_t)reduce_event, 0);
                rs = reduce_stream;
            }
            CPG_LAUNCH(k_wgrad_hh_reduce, CPG_RED_GRID(3 * H * H), CPG_RED_BLOCK, 0, rs, part, nsplit, HP, H, dW);
            CPG_LAUNCH(k_dtable_reduce, CPG_RED_GRID(V * 4 * HP), CPG_RED_BLOCK, 0, rs, dt_part, nsplit, V * 4 * HP, dT);
            return true;                               // W_hh gradient (h0 rows included) and dT both done
        }
    }
#endif
    launch_wgrad_hh_simt(s, HP, H, dg, hs, h0, B, L, sm_count, part, dW);
    return false;
}

void launch_wgrad_partial_reduce(cudaStream_t s, int HP, int H, int V, const float* part_w, const float* part_t, int nsplit,
                                 float* dW, float* dT) {
    const int nbw = CPG_RED_GRID(3 * H * H), nt = V * 4 * HP;
    CPG_LAUNCH(k_wgrad_partial_reduce, nbw + CPG_RED_GRID(nt), CPG_RED_BLOCK, 0, s, part_w, part_t, nsplit, HP, H, nbw, nt, dW, dT);
}

void launch_wgrad_hh_simt(cudaStream_t s, int HP, int H, const float* dg, const float* hs, const float* h0, int B, int L,
                          int sm_count, float* part, float* dW) {
    int nrows = B * L;
    int want = max(1, sm_count / 3);
    int rps = ceil_div(ceil_div(nrows, want), WG_ROWS) * WG_ROWS;
    int nsplit = ceil_div(nrows, rps);
    if (HP == ENC_H) {
        CPG_LAUNCH_NAMED("k_wgrad_hh_enc", k_wgrad_hh<ENC_H>, dim3(3, nsplit), (ENC_H / 4) * (ENC_H / 8), 0, s, dg, hs, h0, B, L, rps, part);
    } else {
        CPG_LAUNCH_NAMED("k_wgrad_hh_dec", k_wgrad_hh<DEC_HP>, dim3(3, nsplit), (DEC_HP / 4) * (DEC_HP / 8), 0, s, dg, hs, h0, B, L, rps, part);
    }
    CPG_LAUNCH(k_wgrad_hh_reduce, CPG_RED_GRID(3 * H * H), CPG_RED_BLOCK, 0, s, part, nsplit, HP, H, dW);
}

// token-table gradient: thread c owns column c of the 4*HP planes; accumulators [V][4HP] in smem
__global__ void k_dtable(const float* __restrict__ dg, const uint8_t* __restrict__ tok, int B, int L, int reverse,
                         int HP4, int V, int rows_per_split, float* __restrict__ part) {
    CPG_DYN_SMEM(float, acc);                 // [V][HP4]
    const int c = threadIdx.x;
    for (int v = 0; v < V; ++v) acc[v * HP4 + c] = 0.f;
    const int nrows = B * L;
    const int rbeg = blockIdx.x * rows_per_split, rend = min(nrows, rbeg + rows_per_split);
    // 8 rows per trip: the global loads are issued together (independent), then the shared-memory
    // read-modify-writes (which depend on the token) follow -- otherwise every row pays a full DRAM latency
    for (int row = rbeg; row < rend; row += 8) {
        float v[8];
        int tk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = row + i;
            const bool ok = r < rend;
            const int b = (ok ? r : rbeg) / L, s = (ok ? r : rbeg) % L;
            const int t = reverse ? (L - 1 - s) : s;
            tk[i] = tok[b * L + t];
            v[i] = ok ? dg[(size_t)r * HP4 + c] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[tk[i] * HP4 + c] += v[i];
    }
    float* out = part + (size_t)blockIdx.x * V * HP4;
    for (int v = 0; v < V; ++v) out[v * HP4 + c] = acc[v * HP4 + c];
}
__global__ void k_dtable_reduce(const float* __restrict__ part, int nsplit, int n, float* __restrict__ dT) {
    const int i = blockIdx.x * RED_X + threadIdx.x;
    const bool ok = i < n;
    const float s = block_split_sum(part, (size_t)n, nsplit, (size_t)(ok ? i : 0), ok);
    if (ok && threadIdx.y == 0) dT[i] = s;
}
int dtable_splits(int B, int L, int sm_count) { return max(1, min(B * L, 2 * sm_count)); }
void launch_dtable(cudaStream_t s, int HP, const float* dg, const uint8_t* tok, int B, int L, int reverse, int V,
                   int sm_count, float* part, float* dT) {
    int nrows = B * L;
    int nsplit = dtable_splits(B, L, sm_count);
    int rps = ceil_div(nrows, nsplit);
    nsplit = ceil_div(nrows, rps);
    int HP4 = 4 * HP;
    size_t smem = (size_t)V * HP4 * sizeof(float);
    CPG_SET_MAX_SMEM(k_dtable, smem);
    CPG_LAUNCH(k_dtable, nsplit, HP4, smem, s, dg, tok, B, L, reverse, HP4, V, rps, part);
    CPG_LAUNCH(k_dtable_reduce, CPG_RED_GRID(V * HP4), CPG_RED_BLOCK, 0, s, part, nsplit, V * HP4, dT);
}

// input-side parameter gradients from the three table gradients
__global__ void k_input_grads(InputGradArgs a, int task0) {
    const int task = blockIdx.y + task0;
    const int stride = gridDim.x * blockDim.x;
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int V = a.V;
    if (task < 2) {                         // encoder dW_ih [240][150], db_ih, db_hh
        const int d = task, H = ENC_H, G = 3 * H, HP4 = 4 * H;
        const float* dT = a.dT_enc[d];
        for (int i = t0; i < G * EMB; i += stride) {
            int g = i / EMB, e = i % EMB;
            float s = 0.f;
            for (int v = 0; v < V; ++v) s = fmaf(dT[v * HP4 + g], a.emb[v * EMB + e], s);
            a.g_enc_wih[d][i] = s;
        }
        for (int g = t0; g < G; g += stride) {
            float si = 0.f, sh = 0.f;
            for (int v = 0; v < V; ++v) {
                si += dT[v * HP4 + g];
                sh += (g < 2 * H) ? dT[v * HP4 + g] : dT[v * HP4 + 3 * H + (g - 2 * H)];
            }
            a.g_enc_bih[d][g] = si;
            a.g_enc_bhh[d][g] = sh;
        }
    } else if (task == 2) {                 // decoder dW_ih [306][252] (embedding cols from dT, zc cols from dwizc)
        const int H = DEC_H, HP = DEC_HP, G = 3 * H, HP4 = 4 * HP;
        const float* dT = a.dT_dec;
        for (int i = t0; i < G * DEC_IN; i += stride) {
            int g = i / DEC_IN, e = i % DEC_IN, gate = g / H, j = g % H;
            float s = 0.f;
            if (e < EMB) {
                for (int v = 0; v < V; ++v) s = fmaf(dT[v * HP4 + gate * HP + j], a.emb[v * EMB + e], s);
            } else {
                s = a.dwizc[(size_t)(gate * HP + j) * HP + (e - EMB)];
            }
            a.g_dec_wih[i] = s;
        }
        for (int g = t0; g < G; g += stride) {
            int gate = g / H, j = g % H;
            float si = 0.f, sh = 0.f;
            for (int v = 0; v < V; ++v) {
                si += dT[v * HP4 + gate * HP + j];
                sh += dT[v * HP4 + (gate < 2 ? gate : 3) * HP + j];
            }
            a.g_dec_bih[g] = si;
            a.g_dec_bhh[g] = sh;
        }
    }
}
// embedding gradient [V][150] = sum over the three tables' gate rows of dT[v][g] * W_ih[g][e]; <pad> row stays 0
// (model.py:47).  786-long contraction per output: EG_Y slices of g per output (short dependent chains), summed in
// slice order.
constexpr int EG_Y = 32;
// dec_part: the decoder table's share only (-> emb_dec; its inputs are complete long before the encoder's); else the two
// encoder tables' shares + emb_dec -> g_emb
__global__ void __launch_bounds__(RED_X * EG_Y)
k_emb_grad(InputGradArgs a, int dec_part) {
    __shared__ float red_s[EG_Y][RED_X + 1];
    const int i = blockIdx.x * RED_X + threadIdx.x, y = threadIdx.y;
    const bool ok = i < a.V * EMB;
    const int v = ok ? i / EMB : 0, e = ok ? i % EMB : 0;
    float s0 = 0.f;
    if (ok && v != PAD) {
        if (dec_part) {
            const float* dT = a.dT_dec + v * 4 * DEC_HP;
            for (int g = y; g < 3 * DEC_H; g += EG_Y) {
                const int gate = g / DEC_H, j = g % DEC_H;
                s0 = fmaf(dT[gate * DEC_HP + j], a.dec_wih[(size_t)g * DEC_IN + e], s0);
            }
        } else {
            for (int d = 0; d < 2; ++d) {
                const float* dT = a.dT_enc[d] + v * 4 * ENC_H;
                const float* w = a.enc_wih[d];
                for (int g = y; g < 3 * ENC_H; g += EG_Y) s0 = fmaf(dT[g], w[g * EMB + e], s0);
            }
        }
    }
    red_s[y][threadIdx.x] = s0;
    __syncthreads();
    if (y == 0 && ok) {
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < EG_Y; ++q) s += red_s[q][threadIdx.x];
        if (dec_part) a.emb_dec[i] = s;
        else a.g_emb[i] = s + a.emb_dec[i];
    }
}
// parts: 1 = encoder tasks, 2 = decoder task, 8 = decoder share of the embedding gradient, 4 = embedding gradient (after 8)
void launch_input_grads(cudaStream_t s, const InputGradArgs& a, cudaStream_t s_emb, int parts) {
    if ((parts & 3) == 3) CPG_LAUNCH(k_input_grads, dim3(148, 3), 256, 0, s, a, 0);
    else if (parts & 1) CPG_LAUNCH(k_input_grads, dim3(148, 2), 256, 0, s, a, 0);
    else if (parts & 2) CPG_LAUNCH(k_input_grads, dim3(148, 1), 256, 0, s, a, 2);
    if (parts & 8) CPG_LAUNCH(k_emb_grad, CPG_RED_GRID(a.V * EMB), dim3(RED_X, EG_Y), 0, s_emb ? s_emb : s, a, 1);
    if (parts & 4) CPG_LAUNCH(k_emb_grad, CPG_RED_GRID(a.V * EMB), dim3(RED_X, EG_Y), 0, s_emb ? s_emb : s, a, 0);   // independent of k_input_grads
}

}  // namespace cpg
