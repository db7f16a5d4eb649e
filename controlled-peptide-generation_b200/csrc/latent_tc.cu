// The dense layers around the latent code on the 5th-gen tensor cores, fused with the element-wise work between them
// (two launches instead of six on the critical path of the iteration):
//
//   k_latent_fwd_tc   hfin -> [q_mu | q_logvar] heads (models/encoder.py:50-51) -> z = mu + exp(logvar/2) eps
//                     (models/model.py:107-112) -> [z;c] -> its input projection W_ih[:,150:] [z;c] for the decoder GRU
//                     (models/decoder.py:70-77: the non-embedding columns of the GRU input)
//   k_latent_bwd_tc   drow . W_ih[:,150:] + dh0 (gradient at [z;c]) -> gradients of the latent losses and of the
//                     reparameterisation -> dmu, dlogvar -> d hfin = dmu W_mu + dlogvar W_logvar
//
// One CTA = M batch rows (M = 64: TMEM lanes 0-15 of each quadrant, 64 CTAs at B = 4096; or 128).  Every contraction is
// a split product (x = x1 + x2, three products, fp32 accumulation; a two-term product of this projection failed the 1e-4
// logits bar in round 1).  The weight operands arrive pre-split in their shared-memory tile image (k_prep_weights builds
// them once per step, latent.h: LT_*) as ONE bulk copy per stage, issued while the threads convert the activation
// operand (coalesced 16-byte loads) or run the previous epilogue; one thread issues the MMAs; the epilogue
// (thread = row) does the element-wise part with 16-byte accesses and writes the next stage's A operand straight into
// shared memory.
#include "ctx.h"
#ifndef CPG_EMU
#include "tc_dense.cuh"

namespace cpg {
int check_launch(const char* where);

namespace {
constexpr int KH = 2 * ENC_H;            // 160: K of the heads
constexpr int NHD = LT_F1_ROWS;          // 2 * 100 head outputs (mu_j, logvar_j interleaved) padded to a multiple of 16
constexpr int KZ = LT_F2_K;              // [z;c] (102 -> 104) padded to a multiple of 16
constexpr int NG = LT_F2_ROWS;           // 3 * 104 gate rows padded to a multiple of 16 (two MMAs of N = 160)
constexpr int KG = LT_B1_K;              // K of the backward's first product
constexpr int KD = LT_B2_K;              // K of the backward's second product: (dmu_j, dlv_j) interleaved
constexpr int NHF = LT_B2_N;             // N of the backward's second product (hfin columns)
static_assert(KH == LT_F1_K && KZ == LT_B1_ROWS && NHF == KH, "tile geometry");

struct LatFwdArgs {
    const float* hfin;      // [B][160]
    const float* bmu; const float* blv;     // [100]
    const float* eps;       // [B][100] or null (z = mu)
    const float* c;         // [B][2]
    const unsigned char* tiles;     // LT_F1 | LT_F2 | ... (latent.h)
    float* mu; float* logvar; float* z; float* zc;       // [B][100] x3 (z may be null), [B][104]
    float* rowbias;         // [B][312] or null (encoder-only inference)
    int B;
};

template <int M> struct FwdCfg {
    static constexpr int A1 = M * KH * 2;                    // bytes per term of the hfin tile
    static constexpr int A2 = M * KZ * 2;                    // bytes per term of the [z;c] tile
    // stage 1: A1 (2 terms) | B1 = LT_F1; stage 2 reuses the region: A2 (written by the epilogue of stage 1) | B2 = LT_F2
    static constexpr size_t S1 = 2 * (size_t)A1 + 2 * LT_F1_TERM, S2 = 2 * (size_t)A2 + 2 * LT_F2_TERM;
    static constexpr size_t SMEM = S1 > S2 ? S1 : S2;
    static_assert(SMEM <= 227 * 1024, "shared memory");
};

template <int M>
__global__ void __launch_bounds__(LT_THREADS, 1)
k_latent_fwd_tc(LatFwdArgs a) {
    using C = FwdCfg<M>;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* A1 = smem;
    unsigned char* B1 = smem + 2 * C::A1;
    unsigned char* A2 = smem;
    unsigned char* B2 = smem + 2 * C::A2;
    __shared__ __align__(8) uint64_t bar_mma, bar_w;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = blockIdx.x * M, B = a.B;
    if (warp == 0) {
        if (lane == 0) {
            tc::mbar_init(&bar_mma, 1);
            tc::mbar_init(&bar_w, 1);
            tc::fence_barrier_init();
            tc::mbar_expect_tx(&bar_w, (uint32_t)(2 * LT_F1_TERM));
            bulk_load(B1, a.tiles + LT_F1_OFF, (uint32_t)(2 * LT_F1_TERM), &bar_w);
        }
        __syncwarp();
        tc::tmem_alloc<512>(&tmem_slot);
    }
    fill_rows_kmajor<true, M, KH>(A1, A1 + C::A1, a.hfin + (size_t)row0 * KH, KH, B - row0, KH);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        tc::mbar_wait(&bar_w, 0);
        const uint32_t a1 = tc::smem_u32(A1), b1 = tc::smem_u32(B1);
        issue_products(tmem, M, a1, a1 + C::A1, M, b1, b1 + (uint32_t)LT_F1_TERM, false, NHD, 0, NHD, KH, true, true);
        tc::umma_commit(&bar_mma);
    }
    tc::mbar_wait(&bar_mma, 0);                                // the MMAs are done: the stage-1 operands may be overwritten
    tc::tc_fence_after();
    if (tid == 0 && a.rowbias != nullptr) {                    // next weight tile lands while the epilogue runs
        tc::mbar_expect_tx(&bar_w, (uint32_t)(2 * LT_F2_TERM));
        bulk_load(B2, a.tiles + LT_F2_OFF, (uint32_t)(2 * LT_F2_TERM), &bar_w);
    }
    const int q = warp & 3, part = warp >> 2;                  // LT_PARTS warps per lane quadrant, 16-column slices in turn
    bool has;
    const int r = epi_row<M>(q, lane, has), row = row0 + r;
    const bool live = has && row < B;
    // ---- epilogue 1: heads + bias, reparameterisation, [z;c]; 16 columns = 8 (mu_j, logvar_j) pairs = one K chunk of [z;c]
    for (int c0 = part * 16; c0 < KZ * 2; c0 += 16 * LT_PARTS) {
        float zq[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) zq[e] = 0.f;
        if (c0 < NHD) {
            float v[16];
            tmem_ld_cols<16>(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            const int j0 = c0 >> 1;                            // 8 latent dims j0 .. j0 + 7 (j0 + 8 <= 104)
            if (live) {
                float m[8], lv[8], ep[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int j = j0 + e;
                    m[e] = v[2 * e] + (j < ZD ? __ldg(a.bmu + j) : 0.f);
                    lv[e] = v[2 * e + 1] + (j < ZD ? __ldg(a.blv + j) : 0.f);
                    ep[e] = 0.f;
                }
                const size_t o = (size_t)row * ZD + j0;
                if (a.eps != nullptr) {
                    const float4 e0 = ld_stream4(a.eps + o);
                    ep[0] = e0.x; ep[1] = e0.y; ep[2] = e0.z; ep[3] = e0.w;
                    if (j0 + 4 < ZD) {
                        const float4 e1 = ld_stream4(a.eps + o + 4);
                        ep[4] = e1.x; ep[5] = e1.y; ep[6] = e1.z; ep[7] = e1.w;
                    }
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) zq[e] = a.eps != nullptr ? m[e] + expf(lv[e] / 2) * ep[e] : m[e];
                if (j0 + 4 >= ZD) {                            // last slice: j = 96..99 latent, 100/101 the code c, 102/103 padding
                    zq[4] = a.c != nullptr ? __ldg(a.c + (size_t)row * CD) : 0.f;
                    zq[5] = a.c != nullptr ? __ldg(a.c + (size_t)row * CD + 1) : 0.f;
                    zq[6] = zq[7] = 0.f;
                }
                st4(a.mu + o, make_float4(m[0], m[1], m[2], m[3]));
                st4(a.logvar + o, make_float4(lv[0], lv[1], lv[2], lv[3]));
                if (a.z != nullptr) st4(a.z + o, make_float4(zq[0], zq[1], zq[2], zq[3]));
                if (j0 + 4 < ZD) {
                    st4(a.mu + o + 4, make_float4(m[4], m[5], m[6], m[7]));
                    st4(a.logvar + o + 4, make_float4(lv[4], lv[5], lv[6], lv[7]));
                    if (a.z != nullptr) st4(a.z + o + 4, make_float4(zq[4], zq[5], zq[6], zq[7]));
                }
                float* zc = a.zc + (size_t)row * DEC_HP + j0;
                st4(zc, make_float4(zq[0], zq[1], zq[2], zq[3]));
                st4(zc + 4, make_float4(zq[4], zq[5], zq[6], zq[7]));
            }
        }
        if (has) {
            uint4 h, l;
            split8x<true>(zq, h, l);
            const int off = (c0 >> 4) * (M * 16) + (r >> 3) * 128 + (r & 7) * 16;
            *reinterpret_cast<uint4*>(A2 + off) = h;
            *reinterpret_cast<uint4*>(A2 + C::A2 + off) = l;
        }
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();                                           // accumulator drained, A2 complete
    tc::tc_fence_after();
    if (a.rowbias == nullptr) {                                // encoder-only call
        if (warp == 0) tc::tmem_dealloc<512>(tmem);
        return;
    }
    if (tid == 0) {
        tc::mbar_wait(&bar_w, 1);
        const uint32_t a2 = tc::smem_u32(A2), b2 = tc::smem_u32(B2);
        issue_products(tmem, M, a2, a2 + C::A2, M, b2, b2 + (uint32_t)LT_F2_TERM, false, NG, 0, 160, KZ, true, true);
        issue_products(tmem + 160, M, a2, a2 + C::A2, M, b2, b2 + (uint32_t)LT_F2_TERM, false, NG, 160, 160, KZ, true, true);
        tc::umma_commit(&bar_mma);
    }
    tc::mbar_wait(&bar_mma, 1);
    tc::tc_fence_after();
    for (int c0 = part * 16; c0 < NG; c0 += 16 * LT_PARTS) {
        float v[16];
        tmem_ld_cols<16>(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        if (live) {
#pragma unroll
            for (int e = 0; e < 16; e += 4)
                if (c0 + e < 3 * DEC_HP) st4(a.rowbias + (size_t)row * (3 * DEC_HP) + c0 + e, make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]));
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

struct LatBwdArgs {
    const float* drow;      // [B][312]
    const float* dh0;       // [B][104]
    const unsigned char* tiles;
    LatentBwdArgs lat;      // mu, logvar, eps, dz_rf, dz_ext, dmu_ext, dlv_ext, weights, B, B_global, dmu, dlv
    float* dhfin;           // [B][160]
    const float* hfin;      // [B][160]
    float* hg_part;         // [CTAs][LT_HG_ROWS][LT_HG_COLS] per-CTA partials of the head weight / bias gradients, or null
};
constexpr int BW_M = 64;
constexpr int BW_A1 = BW_M * KG * 2;          // bytes per term of the drow tile (full K = 320)
constexpr int BW_A2 = BW_M * KD * 2;          // bytes per term of the (dmu, dlv) tile
// stage 1: A1 (2 terms) | B1 = LT_B1; stage 2: A2 (written by the epilogue of stage 1) | B2 = LT_B2
// stage 3 (head weight gradients, contracted over the CTA's 64 batch rows): A = the (dmu, dlv) tile read MN-major,
// B3 = [hfin | 1 | 0...] rows as an MN-major tile (n fastest) behind B2
constexpr int BW_B3 = (LT_HG_COLS / 8) * (BW_M * 16);          // bytes per term
constexpr size_t BW_S1 = 2 * (size_t)BW_A1 + 2 * LT_B1_TERM, BW_S2 = 2 * (size_t)BW_A2 + 2 * LT_B2_TERM + 2 * (size_t)BW_B3;
constexpr size_t BW_SMEM = BW_S1 > BW_S2 ? BW_S1 : BW_S2;
static_assert(BW_SMEM + 1024 <= 227 * 1024, "shared memory");
static_assert(LT_HG_ROWS == KD && LT_HG_COLS % 16 == 0 && LT_HG_COLS > NHF, "head-gradient partial geometry");

__device__ __forceinline__ void ld4_to(const float* p, float* x) {
    const float4 v = ld_stream4(p);
    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
}

__global__ void __launch_bounds__(LT_THREADS, 1)
k_latent_bwd_tc(LatBwdArgs a) {
    constexpr int M = BW_M;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* A1 = smem;
    unsigned char* B1 = smem + 2 * BW_A1;
    unsigned char* A2 = smem;
    unsigned char* B2 = smem + 2 * BW_A2;
    unsigned char* B3 = B2 + 2 * LT_B2_TERM;
    __shared__ __align__(8) uint64_t bar_mma, bar_w;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = blockIdx.x * M, B = a.lat.B;
    if (warp == 0) {
        if (lane == 0) {
            tc::mbar_init(&bar_mma, 1);
            tc::mbar_init(&bar_w, 1);
            tc::fence_barrier_init();
            tc::mbar_expect_tx(&bar_w, (uint32_t)(2 * LT_B1_TERM));
            bulk_load(B1, a.tiles + LT_B1_OFF, (uint32_t)(2 * LT_B1_TERM), &bar_w);
        }
        __syncwarp();
        tc::tmem_alloc<512>(&tmem_slot);
    }
    fill_rows_kmajor<false, M, KG>(A1, A1 + BW_A1, a.drow + (size_t)row0 * (3 * DEC_HP), 3 * DEC_HP, B - row0, 3 * DEC_HP);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        tc::mbar_wait(&bar_w, 0);
        const uint32_t a1 = tc::smem_u32(A1), b1 = tc::smem_u32(B1);
        issue_products(tmem, M, a1, a1 + BW_A1, M, b1, b1 + (uint32_t)LT_B1_TERM, false, KZ, 0, KZ, KG, true, false);
        tc::umma_commit(&bar_mma);
    }
    tc::mbar_wait(&bar_mma, 0);
    tc::tc_fence_after();
    if (tid == 0) {
        tc::mbar_expect_tx(&bar_w, (uint32_t)(2 * LT_B2_TERM));
        bulk_load(B2, a.tiles + LT_B2_OFF, (uint32_t)(2 * LT_B2_TERM), &bar_w);
    }
    if (a.hg_part != nullptr) {
        // B3: element (n, k = batch row b) at (n/8) * M * 16 + (b/8) * 128 + (b%8) * 16 + (n%8) * 2; n < 160: hfin[b][n],
        // n = 160: 1 (the bias gradients ride as one more column), rest 0
        for (int i = tid; i < (LT_HG_COLS / 8) * M; i += LT_THREADS) {
            const int b = i % M, nc = i / M;
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = 0.f;
            if (row0 + b < B) {
                if (nc < NHF / 8) {
                    ld4_to(a.hfin + (size_t)(row0 + b) * KH + nc * 8, x);
                    ld4_to(a.hfin + (size_t)(row0 + b) * KH + nc * 8 + 4, x + 4);
                } else if (nc == NHF / 8) {
                    x[0] = 1.f;
                }
            }
            uint4 h, l;
            split8(x, h, l);
            const int off = nc * (M * 16) + (b >> 3) * 128 + (b & 7) * 16;
            *reinterpret_cast<uint4*>(B3 + off) = h;
            *reinterpret_cast<uint4*>(B3 + BW_B3 + off) = l;
        }
    }
    const int q = warp & 3, part = warp >> 2;
    bool has;
    const int r = epi_row<M>(q, lane, has), row = row0 + r;
    const bool live = has && row < B;
    // ---- epilogue 1: gradient at z -> (dmu, dlogvar) (the arithmetic of k_latent_bwd), next A operand (dmu_j, dlv_j) interleaved
    {
        const LatentBwdArgs& L = a.lat;
        const float invB = 1.0f / (float)L.B_global;
        const float w_kl = (L.dyn != nullptr && L.w_kl != 0.f) ? L.dyn->beta : L.w_kl;
        for (int c0 = part * 16; c0 < KZ; c0 += 16 * LT_PARTS) {
            float v[16];
            tmem_ld_cols<16>(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {                   // 4 latent dims -> 8 interleaved K entries = one 16-byte chunk
                const int j0 = c0 + g4 * 4;
                float x[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) x[e] = 0.f;
                if (j0 < ZD && live) {
                    const size_t i = (size_t)row * ZD + j0;
                    float dz[4], m[4], lv[4], t[4], dmu[4], dlv[4];
                    ld4_to(a.dh0 + (size_t)row * DEC_HP + j0, t);
#pragma unroll
                    for (int e = 0; e < 4; ++e) dz[e] = v[g4 * 4 + e] + t[e];
                    if (L.dz_rf != nullptr) {
                        ld4_to(L.dz_rf + i, t);
#pragma unroll
                        for (int e = 0; e < 4; ++e) dz[e] += t[e];
                    }
                    if (L.dz_ext != nullptr) {
                        ld4_to(L.dz_ext + i, t);
#pragma unroll
                        for (int e = 0; e < 4; ++e) dz[e] += t[e];
                    }
                    ld4_to(L.mu + i, m);
                    ld4_to(L.logvar + i, lv);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float ex = expf(lv[e]);
                        dmu[e] = dz[e] + w_kl * m[e] * invB;
                        dlv[e] = (w_kl + L.w_klsm) * 0.5f * (ex - 1.0f) * invB;
                        dlv[e] += L.w_l1 * (lv[e] > 0.f ? 1.f : (lv[e] < 0.f ? -1.f : 0.f)) * invB;
                    }
                    if (L.eps != nullptr) {
                        ld4_to(L.eps + i, t);
#pragma unroll
                        for (int e = 0; e < 4; ++e) dlv[e] += dz[e] * t[e] * 0.5f * expf(lv[e] / 2);
                    }
                    if (L.dmu_ext != nullptr) {
                        ld4_to(L.dmu_ext + i, t);
#pragma unroll
                        for (int e = 0; e < 4; ++e) dmu[e] += t[e];
                    }
                    if (L.dlv_ext != nullptr) {
                        ld4_to(L.dlv_ext + i, t);
#pragma unroll
                        for (int e = 0; e < 4; ++e) dlv[e] += t[e];
                    }
                    st4(L.dmu + i, make_float4(dmu[0], dmu[1], dmu[2], dmu[3]));
                    st4(L.dlv + i, make_float4(dlv[0], dlv[1], dlv[2], dlv[3]));
#pragma unroll
                    for (int e = 0; e < 4; ++e) { x[2 * e] = dmu[e]; x[2 * e + 1] = dlv[e]; }
                }
                const int kc = j0 >> 2;                        // K index 2 j -> chunk j / 4
                if (has && kc < KD / 8) {
                    uint4 h, l;
                    split8(x, h, l);
                    const int off = kc * (M * 16) + (r >> 3) * 128 + (r & 7) * 16;
                    *reinterpret_cast<uint4*>(A2 + off) = h;
                    *reinterpret_cast<uint4*>(A2 + BW_A2 + off) = l;
                }
            }
        }
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    // stage 2: d hfin = [dmu | dlv] . [W_mu ; W_logvar]; B (MN-major): element (n = hfin column, k = 2 j | 2 j + 1) = W_mu[j][n] | W_logvar[j][n]
    if (tid == 0) {
        tc::mbar_wait(&bar_w, 1);
        const uint32_t a2 = tc::smem_u32(A2), b2 = tc::smem_u32(B2);
        issue_products(tmem, M, a2, a2 + BW_A2, M, b2, b2 + (uint32_t)LT_B2_TERM, true, KD, 0, NHF, KD, true, false);
        if (a.hg_part != nullptr) {
            // stage 3: [dW_mu ; dW_logvar | db] (rows j' = 2 j | 2 j + 1) = (dmu, dlv)^T . [hfin | 1]: the A2 tile read MN-major
            // (m = j'), two M = 128 tiles (rows 0..127 and 80..207), K = the CTA's 64 batch rows
            const uint32_t b3 = tc::smem_u32(B3);
            issue_products(tmem + 160, 128, a2, a2 + BW_A2, M, b3, b3 + BW_B3, true, M, 0, LT_HG_COLS, M, true, false, true, 0);
            issue_products(tmem + 160 + LT_HG_COLS, 128, a2, a2 + BW_A2, M, b3, b3 + BW_B3, true, M, 0, LT_HG_COLS, M, true, false, true, KD - 128);
        }
        tc::umma_commit(&bar_mma);
    }
    tc::mbar_wait(&bar_mma, 1);
    tc::tc_fence_after();
    for (int c0 = part * 16; c0 < NHF; c0 += 16 * LT_PARTS) {
        float v[16];
        tmem_ld_cols<16>(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        if (live) {
#pragma unroll
            for (int e = 0; e < 16; e += 4) st4(a.dhfin + (size_t)row * KH + c0 + e, make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]));
        }
    }
    if (a.hg_part != nullptr) {
        // partial of this CTA: tile 0 holds rows j' = TMEM lane, tile 1 rows j' = 80 + lane (only j' >= 128 taken from it)
        float* out = a.hg_part + (size_t)blockIdx.x * LT_HG_ROWS * LT_HG_COLS;
        constexpr int NCH = LT_HG_COLS / 16;
        for (int it = part; it < 2 * NCH; it += LT_PARTS) {
            const int tile = it / NCH, c0 = (it % NCH) * 16;
            const int jr = tile == 0 ? q * 32 + lane : (KD - 128) + q * 32 + lane;
            if (tile == 1 && q == 0) continue;                 // rows 80..111: already covered by tile 0
            float v[16];
            tmem_ld_cols<16>(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(160 + tile * LT_HG_COLS + c0), v);
            if (tile == 0 || jr >= 128) {
#pragma unroll
                for (int e = 0; e < 16; e += 4) st4(out + (size_t)jr * LT_HG_COLS + c0 + e, make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]));
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

// head weight / bias gradients from the per-CTA partials, summed in CTA order: row j' = 2 j (q_mu) | 2 j + 1 (q_logvar),
// columns 0..159 the weight row, column 160 the bias
__global__ void k_head_grad_reduce(const float* __restrict__ part, int n_cta, float* __restrict__ g_wmu, float* __restrict__ g_wlv,
                                   float* __restrict__ g_bmu, float* __restrict__ g_blv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * ZD * (NHF + 1)) return;
    const int jr = i / (NHF + 1), n = i % (NHF + 1);
    const float* p = part + (size_t)jr * LT_HG_COLS + n;
    float acc = 0.f;
    for (int c = 0; c < n_cta; ++c) acc += p[(size_t)c * LT_HG_ROWS * LT_HG_COLS];
    const int j = jr >> 1;
    if (n < NHF) ((jr & 1) ? g_wlv : g_wmu)[(size_t)j * NHF + n] = acc;
    else ((jr & 1) ? g_blv : g_bmu)[j] = acc;
}

template <int M>
int launch_fwd(cudaStream_t s, const LatFwdArgs& a) {
    static bool set = false;
    if (!set) {
        if (cudaFuncSetAttribute((const void*)k_latent_fwd_tc<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FwdCfg<M>::SMEM) != cudaSuccess) {
            cudaGetLastError();
            return CPG_ECUDA;
        }
        set = true;
    }
    CPG_LAUNCH(k_latent_fwd_tc<M>, ceil_div(a.B, M), LT_THREADS, FwdCfg<M>::SMEM, s, a);
    return CPG_OK;
}
}  // namespace

int g_opt_latent_tc = 1;      // 1: tcgen05 dense layers when B >= 512, 2: always, 0: fp32 SIMT GEMMs + element-wise kernels
int g_opt_latent_rows = 64;   // batch rows per CTA of the forward kernel (64 | 128)
bool latent_uses_tc(int B) { return g_opt_latent_tc == 2 || (g_opt_latent_tc == 1 && B >= 512); }

int launch_latent_fwd_tc(cudaStream_t s, const float* hfin, const float* bmu, const float* blv, const float* eps, const float* c,
                         const unsigned char* tiles, int B, float* mu, float* logvar, float* z, float* zc, float* rowbias) {
    LatFwdArgs a{hfin, bmu, blv, eps, c, tiles, mu, logvar, z, zc, rowbias, B};
    return g_opt_latent_rows == 128 ? launch_fwd<128>(s, a) : launch_fwd<64>(s, a);
}

int latent_bwd_tc_ctas(int B) { return ceil_div(B, BW_M); }

int launch_latent_bwd_tc(cudaStream_t s, const float* drow, const float* dh0, const unsigned char* tiles, const LatentBwdArgs& lat,
                         float* dhfin, const float* hfin, float* hg_part) {
    LatBwdArgs a;
    a.drow = drow; a.dh0 = dh0; a.tiles = tiles; a.lat = lat; a.lat.dyn = g_dyn; a.dhfin = dhfin; a.hfin = hfin; a.hg_part = hg_part;
    static bool set = false;
    if (!set) {
        if (cudaFuncSetAttribute((const void*)k_latent_bwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BW_SMEM) != cudaSuccess) { cudaGetLastError(); return CPG_ECUDA; }
        set = true;
    }
    CPG_LAUNCH(k_latent_bwd_tc, ceil_div(lat.B, BW_M), LT_THREADS, BW_SMEM, s, a);
    return CPG_OK;
}

void launch_head_grad_reduce(cudaStream_t s, const float* hg_part, int B, float* g_wmu, float* g_wlv, float* g_bmu, float* g_blv) {
    CPG_LAUNCH(k_head_grad_reduce, ceil_div(2 * ZD * (NHF + 1), 256), 256, 0, s, hg_part, ceil_div(B, BW_M), g_wmu, g_wlv, g_bmu, g_blv);
}

}  // namespace cpg
#else
namespace cpg {
int g_opt_latent_tc = 0, g_opt_latent_rows = 64;
bool latent_uses_tc(int) { return false; }
int launch_latent_fwd_tc(cudaStream_t, const float*, const float*, const float*, const float*, const float*, const unsigned char*, int,
                         float*, float*, float*, float*, float*) { return CPG_ECUDA; }
int latent_bwd_tc_ctas(int) { return 1; }
int launch_latent_bwd_tc(cudaStream_t, const float*, const float*, const unsigned char*, const LatentBwdArgs&, float*, const float*, float*) { return CPG_ECUDA; }
void launch_head_grad_reduce(cudaStream_t, const float*, int, float*, float*, float*, float*) {}
}  // namespace cpg
#endif
