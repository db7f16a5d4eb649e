// Weight gradients of the dense layers around the latent code -- contractions over the BATCH:
//     dW_ih[:,150:] [312 x 104]  = drow^T [312 x B] . [z;c] [B x 104]               (models/decoder.py:70-77)
//     dW_mu | db_mu [100 x 161]  = dmu^T  [100 x B] . [hfin | 1] [B x 161]           (models/encoder.py:50-51)
//     dW_lv | db_lv              = dlv^T . [hfin | 1]
// on the 5th-gen tensor cores (split bf16, three products, fp32 accumulation in TMEM across all row tiles of a CTA).
// Both operands are row-major [batch][feature] in HBM, i.e. MN-major for this contraction: a 16-byte chunk of a tile is 8
// consecutive features of one batch row -- the conversion is a straight copy + split, no transpose.
// Nothing on the dependent chain of the iteration waits for these gradients, so the kernel runs on the reduction lane as a
// NARROW persistent grid (WD_GRID CTAs, 64-row tiles in turn, one partial per CTA; see rf_tc.cu for why narrow) with a
// shared-memory footprint that keeps it from sharing an SM -- and its tensor memory -- with a chain kernel.
#include <string.h>
#include "ctx.h"
#ifndef CPG_EMU
#include "tc_dense.cuh"

namespace cpg {
int wgrad_dense_ctas(int B);
namespace {
constexpr int WD_ROWS = 64;                  // batch rows per tile = K of one MMA batch
constexpr int WD_GRID = 16;
constexpr int WD_MAXT = 3;                   // M tiles per problem
constexpr int WD_CHUNK = WD_ROWS * 16;       // bytes of one 8-feature chunk column of a tile (one term)
constexpr size_t WD_MIN_SMEM = 132 * 1024;   // (see above)

struct WdTile {
    const float* a;          // [B][lda]
    int lda, m0, m_valid;    // features m0 .. m0 + M of A; those >= m_valid read as 0
    int M;                   // 128 | 64
};
struct WdArgs {
    WdTile t[WD_MAXT];
    int ntiles;
    const float* b;          // [B][ldb]
    int ldb, n_valid, ones_col;      // features >= n_valid read as 0, except column ones_col = 1 (bias gradient); -1: none
    int N;                   // padded, multiple of 16
    int B;
    float* part;             // [CTAs][sum_t M_t][N]
};

__global__ void __launch_bounds__(LT_THREADS, 1)
k_wgrad_dense_tc(WdArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar_mma;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int mtot = 0;
    for (int t = 0; t < a.ntiles; ++t) mtot += a.t[t].M;
    const int a_term = (mtot / 8) * WD_CHUNK, b_term = (a.N / 8) * WD_CHUNK;
    unsigned char* At = smem;                                  // [term][tile-major chunks]
    unsigned char* Bt = smem + 2 * a_term;
    if (warp == 0) {
        if (lane == 0) { tc::mbar_init(&bar_mma, 1); tc::fence_barrier_init(); }
        __syncwarp();
        tc::tmem_alloc<512>(&tmem_slot);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int n_rt = ceil_div(a.B, WD_ROWS);
    uint32_t it = 0;
    for (int rt = blockIdx.x; rt < n_rt; rt += gridDim.x, ++it) {
        const int row0 = rt * WD_ROWS;
        // A: element (m, k = batch row b) of tile t at base_t + (m/8) * WD_CHUNK + (b/8) * 128 + (b%8) * 16 + (m%8) * 2
        int cbase = 0;
        for (int t = 0; t < a.ntiles; ++t) {
            const WdTile T = a.t[t];
#pragma unroll 4
            for (int i = tid; i < (T.M / 8) * WD_ROWS; i += LT_THREADS) {
                const int b = i % WD_ROWS, mc = i / WD_ROWS, m = T.m0 + mc * 8;
                float x[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) x[e] = 0.f;
                if (row0 + b < a.B && m < T.m_valid) {
                    const float* src = T.a + (size_t)(row0 + b) * T.lda + m;
                    const float4 v0 = ld_stream4(src);
                    x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w;
                    if (m + 4 < T.m_valid) {
                        const float4 v1 = ld_stream4(src + 4);
                        x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
                    }
                }
                uint4 h, l;
                split8(x, h, l);
                const int off = (cbase + mc) * WD_CHUNK + (b >> 3) * 128 + (b & 7) * 16;
                *reinterpret_cast<uint4*>(At + off) = h;
                *reinterpret_cast<uint4*>(At + a_term + off) = l;
            }
            cbase += T.M / 8;
        }
#pragma unroll 3
        for (int i = tid; i < (a.N / 8) * WD_ROWS; i += LT_THREADS) {
            const int b = i % WD_ROWS, nc = i / WD_ROWS, n = nc * 8;
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = 0.f;
            if (row0 + b < a.B) {
                if (n < a.n_valid) {
                    const float* src = a.b + (size_t)(row0 + b) * a.ldb + n;
                    const float4 v0 = ld_stream4(src);
                    x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w;
                    if (n + 4 < a.n_valid) {
                        const float4 v1 = ld_stream4(src + 4);
                        x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
                    }
                }
                if (a.ones_col >= n && a.ones_col < n + 8) x[a.ones_col - n] = 1.f;
            }
            uint4 h, l;
            split8(x, h, l);
            const int off = nc * WD_CHUNK + (b >> 3) * 128 + (b & 7) * 16;
            *reinterpret_cast<uint4*>(Bt + off) = h;
            *reinterpret_cast<uint4*>(Bt + b_term + off) = l;
        }
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
        if (tid == 0) {
            const uint32_t a0 = tc::smem_u32(At), b0 = tc::smem_u32(Bt);
            int cb = 0, col = 0;
            for (int t = 0; t < a.ntiles; ++t) {
                const uint32_t at = a0 + cb * WD_CHUNK;
                issue_products(tmem + col, a.t[t].M, at, at + a_term, WD_ROWS, b0, b0 + b_term, true, WD_ROWS, 0, a.N, WD_ROWS,
                               it == 0, false, true, 0);
                cb += a.t[t].M / 8;
                col += a.N;
            }
            tc::umma_commit(&bar_mma);
        }
        tc::mbar_wait(&bar_mma, it & 1);                       // the operand tiles may be overwritten
        tc::tc_fence_after();
    }
    // ---- partial of this CTA: [tile rows m][N]
    {
        const int q = warp & 3, part = warp >> 2;
        float* out = a.part + (size_t)blockIdx.x * mtot * a.N;
        int mrow = 0, col = 0;
        for (int t = 0; t < a.ntiles; ++t) {
            const int M = a.t[t].M;
            const bool has = M == 128 || lane < 16;
            const int r = M == 128 ? q * 32 + lane : q * 16 + (lane & 15);
            for (int c0 = part * 16; c0 < a.N; c0 += 16 * LT_PARTS) {
                float v[16];
                tmem_ld_cols<16>(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(col + c0), v);
                if (has && it > 0) {
#pragma unroll
                    for (int e = 0; e < 16; e += 4) st4(out + (size_t)(mrow + r) * a.N + c0 + e, make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]));
                } else if (has) {                              // a CTA without a row tile: zeros
#pragma unroll
                    for (int e = 0; e < 16; e += 4) st4(out + (size_t)(mrow + r) * a.N + c0 + e, make_float4(0.f, 0.f, 0.f, 0.f));
                }
            }
            mrow += M;
            col += a.N;
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

// out[dst(i)] = sum over the CTAs' partials, in CTA order.  mode 0: dW_ih[:,150:] padded [3*104][104] from rows g of the
// [320][112] partial; mode 1: head gradients -- rows 0..99 of tile 0 -> (W_mu row | b_mu), rows 0..99 of tile 1 (at row 128)
// -> (W_logvar row | b_logvar)
__global__ void k_wgrad_dense_reduce(const float* __restrict__ part, int ncta, int mtot, int N, int mode, float* __restrict__ o0,
                                     float* __restrict__ o1, float* __restrict__ o2, float* __restrict__ o3) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int row, colx;
    float* dst;
    if (mode == 0) {
        if (i >= 3 * DEC_HP * DEC_HP) return;
        row = i / DEC_HP; colx = i % DEC_HP;
        dst = o0 + i;
    } else {
        constexpr int W = 2 * ENC_H + 1;
        if (i >= 2 * ZD * W) return;
        const int head = i / (ZD * W), j = (i / W) % ZD;
        colx = i % W;
        row = head * 128 + j;
        dst = colx < 2 * ENC_H ? (head ? o1 : o0) + (size_t)j * 2 * ENC_H + colx : (head ? o3 : o2) + j;
    }
    const float* p = part + (size_t)row * N + colx;
    float acc = 0.f;
    for (int c = 0; c < ncta; ++c) acc += p[(size_t)c * mtot * N];
    *dst = acc;
}

int launch_wd(cudaStream_t s, const WdArgs& a, int mtot) {
    const size_t need = 2 * (size_t)(mtot / 8 + a.N / 8) * WD_CHUNK;
    const size_t smem = need > WD_MIN_SMEM ? need : WD_MIN_SMEM;
    static size_t set_for = 0;
    if (set_for < smem) {
        if (cudaFuncSetAttribute((const void*)k_wgrad_dense_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            return CPG_ECUDA;
        }
        set_for = smem;
    }
    CPG_LAUNCH(k_wgrad_dense_tc, wgrad_dense_ctas(a.B), LT_THREADS, smem, s, a);
    return CPG_OK;
}
}  // namespace

int g_opt_wgrad_dense_tc = 1;
int g_opt_wd_grid = 0;
int wgrad_dense_ctas(int B) { return std::min(g_opt_wd_grid > 0 ? g_opt_wd_grid : WD_GRID, ceil_div(B, WD_ROWS)); }
size_t wgrad_dense_part_floats(int B) { return (size_t)std::min(64, ceil_div(B, WD_ROWS)) * 320 * 176; }

// dwizc [3*104][104] (padded layout) = drow^T [z;c]
int launch_wgrad_zc_tc(cudaStream_t s, const float* drow, const float* zc, int B, float* part, float* dwizc) {
    WdArgs a;
    memset(&a, 0, sizeof(a));
    a.t[0] = WdTile{drow, 3 * DEC_HP, 0, 3 * DEC_HP, 128};
    a.t[1] = WdTile{drow, 3 * DEC_HP, 128, 3 * DEC_HP, 128};
    a.t[2] = WdTile{drow, 3 * DEC_HP, 256, 3 * DEC_HP, 64};
    a.ntiles = 3;
    a.b = zc; a.ldb = DEC_HP; a.n_valid = DEC_HP; a.ones_col = -1; a.N = 112;
    a.B = B; a.part = part;
    const int rc = launch_wd(s, a, 320);
    if (rc) return rc;
    CPG_LAUNCH(k_wgrad_dense_reduce, ceil_div(3 * DEC_HP * DEC_HP, 256), 256, 0, s, part, wgrad_dense_ctas(B), 320, 112, 0, dwizc,
               nullptr, nullptr, nullptr);
    return CPG_OK;
}

// head weight / bias gradients = (dmu | dlv)^T [hfin | 1]
int launch_wgrad_heads_tc(cudaStream_t s, const float* dmu, const float* dlv, const float* hfin, int B, float* part, float* g_wmu,
                          float* g_wlv, float* g_bmu, float* g_blv) {
    WdArgs a;
    memset(&a, 0, sizeof(a));
    a.t[0] = WdTile{dmu, ZD, 0, ZD, 128};
    a.t[1] = WdTile{dlv, ZD, 0, ZD, 128};
    a.ntiles = 2;
    a.b = hfin; a.ldb = 2 * ENC_H; a.n_valid = 2 * ENC_H; a.ones_col = 2 * ENC_H; a.N = 176;
    a.B = B; a.part = part;
    const int rc = launch_wd(s, a, 256);
    if (rc) return rc;
    CPG_LAUNCH(k_wgrad_dense_reduce, ceil_div(2 * ZD * (2 * ENC_H + 1), 256), 256, 0, s, part, wgrad_dense_ctas(B), 256, 176, 1, g_wmu,
               g_wlv, g_bmu, g_blv);
    return CPG_OK;
}

}  // namespace cpg
#else
namespace cpg {
int g_opt_wgrad_dense_tc = 1;
int wgrad_dense_ctas(int) { return 1; }
size_t wgrad_dense_part_floats(int) { return 16; }
int launch_wgrad_zc_tc(cudaStream_t, const float*, const float*, int, float*, float*) { return CPG_ECUDA; }
int launch_wgrad_heads_tc(cudaStream_t, const float*, const float*, const float*, int, float*, float*, float*, float*, float*) { return CPG_ECUDA; }
}  // namespace cpg
#endif
