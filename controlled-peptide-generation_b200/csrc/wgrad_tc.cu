// Recurrent weight gradients on the 5th-gen tensor cores:
//     dW_hh[g][k] = sum_{r=(b,s)} dgh[r][g] * h_{s-1}[r][k]         (K = B*L = 102,400 at B=4096)
// a pure contraction over the batch*time axis, split across one CTA per SM, each CTA streaming its
// row range through a TMA -> (tf32 rounding) -> tcgen05.mma pipeline with the three gate tiles
// accumulating in TMEM for the whole range.
//
// Both operands are "MN-major" for UMMA: a row r of `dg` ([B*L][4*HP]) holds the M index (gate column)
// contiguously, a row of `hs` ([B*L][HP]) holds the N index (hidden unit) contiguously; TMA boxes of
// {32 floats, 32 rows} with the 128-byte / 32-byte-atom swizzle (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) land
// exactly in the canonical MN-major SWIZZLE_128B_BASE32B layout -- the only one tcgen05 accepts for MN-major tf32.
// h_{s-1} is the hs row ABOVE (TMA row coordinate r-1, row -1 zero-filled); rows with s == 0 are
// zeroed by the conversion pass (their h0 term is added by the caller for the decoder; the encoder
// starts from h0 = 0).  Operands are rounded to tf32 (round-to-nearest) in shared memory by four
// conversion warps, which keeps the products unbiased (the tensor core itself truncates).
//
// Warp roles (192 threads): warp 4 = TMA producer, warp 5 = MMA issuer (+ TMEM owner),
// warps 0-3 = tf32 conversion per stage, then the TMEM -> global epilogue.
#include "ctx.h"
#ifndef CPG_EMU
#include "tc_common.cuh"
#include <stdio.h>

namespace cpg {
int check_launch(const char* where);

constexpr int WT_RK = 32;                 // reduction rows per pipeline stage
constexpr int WT_CHUNK = WT_RK * 128;     // bytes of one {32 floats x 32 rows} box
constexpr int WT_STAGES = 3;

template <int HP>
struct WgTc {
    static constexpr int NCH = (HP + 31) / 32;                   // 32-float chunks that hold valid columns
    static constexpr int N_MMA = (HP + 15) / 16 * 16;            // UMMA N (multiple of 16 for M = 128)
    static constexpr int A_BYTES = 3 * 4 * WT_CHUNK;             // 3 gates x 4 chunks (M = 128 lanes)
    static constexpr int B_BYTES = 4 * WT_CHUNK;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr size_t SMEM = (size_t)WT_STAGES * STAGE_BYTES + 1024;
};

template <int HP>
__global__ void __launch_bounds__(192, 1)
k_wgrad_hh_tc(const __grid_constant__ CUtensorMap tmap_dg, const __grid_constant__ CUtensorMap tmap_hs, int nrows,
              int L, int rows_per_cta, float* __restrict__ part, int dbg) {
    using C = WgTc<HP>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar_full[WT_STAGES], bar_conv[WT_STAGES], bar_empty[WT_STAGES], bar_done;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r_begin = blockIdx.x * rows_per_cta;
    const int r_end = min(nrows, r_begin + rows_per_cta);
    const int n_stage_iters = r_end > r_begin ? (r_end - r_begin + WT_RK - 1) / WT_RK : 0;

    if (warp == 5) {
        if (lane == 0) {
            for (int s = 0; s < WT_STAGES; ++s) {
                tc::mbar_init(&bar_full[s], 1);
                tc::mbar_init(&bar_conv[s], 128);
                tc::mbar_init(&bar_empty[s], 1);
            }
            tc::mbar_init(&bar_done, 1);
            tc::fence_barrier_init();
        }
        __syncwarp();
        tc::tmem_alloc<512>(&tmem_slot);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = tmem_slot;

    if (warp == 4) {
        // ---------------- TMA producer
        if (lane == 0) {
            tc::tma_prefetch_desc(&tmap_dg);
            tc::tma_prefetch_desc(&tmap_hs);
            for (int it = 0; it < n_stage_iters; ++it) {
                const int st = it % WT_STAGES, ph = (it / WT_STAGES) & 1;
                if (it >= WT_STAGES) tc::mbar_wait(&bar_empty[st], ph ^ 1);
                unsigned char* sa = smem + (size_t)st * C::STAGE_BYTES;
                unsigned char* sb = sa + C::A_BYTES;
                const int r0 = r_begin + it * WT_RK;
                tc::mbar_expect_tx(&bar_full[st], (3 * C::NCH + C::NCH) * WT_CHUNK);
                for (int g = 0; g < 3; ++g) {
                    const int plane = g == 2 ? 3 : g;                 // (dr_pre, dz_pre, dhn)
                    for (int c = 0; c < C::NCH; ++c)
                        tc::tma_load_2d(sa + (g * 4 + c) * WT_CHUNK, &tmap_dg, &bar_full[st], plane * HP + c * 32, r0);
                }
                for (int c = 0; c < C::NCH; ++c)
                    tc::tma_load_2d(sb + c * WT_CHUNK, &tmap_hs, &bar_full[st], c * 32, r0 - 1);
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        // ---------------- MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = tc::make_idesc_tf32(128, C::N_MMA, 1, 1);
            for (int it = 0; it < n_stage_iters; ++it) {
                const int st = it % WT_STAGES, ph = (it / WT_STAGES) & 1;
                tc::mbar_wait(&bar_conv[st], ph);
                tc::tc_fence_after();
                const uint32_t sa = tc::smem_u32(smem + (size_t)st * C::STAGE_BYTES);
                const uint32_t sb = sa + C::A_BYTES;
#pragma unroll
                for (int ks = 0; ks < WT_RK / 8; ++ks) {
                    // MN-major SWIZZLE_128B_BASE32B: next 32 floats of M/N one 4 KB box further (LBO),
                    // 4-row K groups 512 B apart (SBO); one K=8 MMA spans two such groups
                    const uint32_t lbo = (uint32_t)WT_CHUNK, sbo = 512u;
                    const uint64_t db = tc::make_smem_desc(sb + ks * 1024, lbo, sbo, 1);
#pragma unroll
                    for (int g = 0; g < 3; ++g) {
                        const uint64_t da = tc::make_smem_desc(sa + g * 4 * WT_CHUNK + ks * 1024, lbo, sbo, 1);
                        tc::umma_tf32(tmem_d + g * 128, da, db, idesc, (it > 0 || ks > 0) ? 1u : 0u);
                    }
                }
                tc::umma_commit(&bar_empty[st]);          // stage reusable once these MMAs have read it
            }
            tc::umma_commit(&bar_done);
        }
        __syncwarp();
    } else {
        // ---------------- tf32 conversion (in place), 128 threads
        for (int it = 0; it < n_stage_iters; ++it) {
            const int st = it % WT_STAGES, ph = (it / WT_STAGES) & 1;
            tc::mbar_wait(&bar_full[st], ph);
            unsigned char* sa = smem + (size_t)st * C::STAGE_BYTES;
            const int r0 = r_begin + it * WT_RK;
            // A: 3 gates x NCH chunks; each chunk = 32 rows x 8 float4 (row i = bytes [128 i, 128 i + 128))
            // thread -> float4 slot (tid & 7) of rows (tid >> 3) and (tid >> 3) + 16 of every 4 KB box
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int row = (tid >> 3) + 16 * half;
                const int r = r0 + row;
                const bool zero_row = (r >= r_end) || (r % L) == 0;
                const int slot = row * 8 + (tid & 7);
#pragma unroll
                for (int g = 0; g < 3; ++g)
#pragma unroll
                    for (int c = 0; c < C::NCH; ++c) {
                        float4* p = reinterpret_cast<float4*>(sa + (g * 4 + c) * WT_CHUNK) + slot;
                        float4 v = *p;
                        if (zero_row) {
                            v = make_float4(0.f, 0.f, 0.f, 0.f);
                        } else {
                            v.x = tc::to_tf32_rn(v.x); v.y = tc::to_tf32_rn(v.y); v.z = tc::to_tf32_rn(v.z); v.w = tc::to_tf32_rn(v.w);
                        }
                        *p = v;
                    }
#pragma unroll
                for (int c = 0; c < C::NCH; ++c) {
                    float4* p = reinterpret_cast<float4*>(sa + C::A_BYTES + c * WT_CHUNK) + slot;
                    float4 v = *p;
                    v.x = tc::to_tf32_rn(v.x); v.y = tc::to_tf32_rn(v.y); v.z = tc::to_tf32_rn(v.z); v.w = tc::to_tf32_rn(v.w);
                    *p = v;
                }
            }
            tc::fence_proxy_async();                      // generic-proxy writes -> visible to the tensor core
            tc::mbar_arrive(&bar_conv[st]);
        }
        // ---------------- epilogue: TMEM -> per-CTA partial [3*HP][HP]
        if (n_stage_iters > 0) {
            tc::mbar_wait(&bar_done, 0);
            tc::tc_fence_after();
        }
        float* out = part + (size_t)blockIdx.x * 3 * HP * HP;
        for (int g = 0; g < 3; ++g) {
#pragma unroll 1
            for (int c = 0; c < C::NCH; ++c) {
                float v[32];
                if (n_stage_iters > 0) {
                    tc::tmem_ld_32x32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(g * 128 + c * 32), v);
                } else {
#pragma unroll
                    for (int q = 0; q < 32; ++q) v[q] = 0.f;
                }
                if (tid < HP) {
                    float* o = out + ((size_t)g * HP + tid) * HP + c * 32;
#pragma unroll
                    for (int q = 0; q < 32; ++q)
                        if (c * 32 + q < HP) o[q] = v[q];
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 5) tc::tmem_dealloc<512>(tmem_d);
}

int wgrad_tc_splits(int sm_count) { return sm_count > 0 ? sm_count : 1; }

int g_dbg_wgrad = 0;
int launch_wgrad_hh_tc(cudaStream_t s, int HP, const float* dg, const float* hs, int B, int L, int sm_count,
                       float* part, int* nsplit_out) {
    const int nrows = B * L;
    int nsplit = std::min(wgrad_tc_splits(sm_count), ceil_div(nrows, WT_RK));
    int rpc = ceil_div(ceil_div(nrows, nsplit), WT_RK) * WT_RK;
    nsplit = ceil_div(nrows, rpc);
    *nsplit_out = nsplit;
    CUtensorMap tm_dg, tm_hs;
    int rc = make_tmap_2d_f32_sw128(&tm_dg, dg, (uint64_t)4 * HP, (uint64_t)nrows, (uint64_t)4 * HP * sizeof(float), WT_RK, true);
    if (rc) return rc;
    rc = make_tmap_2d_f32_sw128(&tm_hs, hs, (uint64_t)HP, (uint64_t)nrows, (uint64_t)HP * sizeof(float), WT_RK, true);
    if (rc) return rc;
    if (HP == ENC_H) {
        auto kfn = k_wgrad_hh_tc<ENC_H>;
        static bool once = false;
        if (!once) { cudaFuncSetAttribute((const void*)kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WgTc<ENC_H>::SMEM); once = true; }
        CPG_LAUNCH_NAMED("k_wgrad_hh_tc_enc", kfn, nsplit, 192, WgTc<ENC_H>::SMEM, s, tm_dg, tm_hs, nrows, L, rpc, part, g_dbg_wgrad);
    } else {
        auto kfn = k_wgrad_hh_tc<DEC_HP>;
        static bool once = false;
        if (!once) { cudaFuncSetAttribute((const void*)kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WgTc<DEC_HP>::SMEM); once = true; }
        CPG_LAUNCH_NAMED("k_wgrad_hh_tc_dec", kfn, nsplit, 192, WgTc<DEC_HP>::SMEM, s, tm_dg, tm_hs, nrows, L, rpc, part, g_dbg_wgrad);
    }
    return CPG_OK;
}

}  // namespace cpg
#endif  // CPG_EMU

#ifndef CPG_EMU
// developer probe: run the tensor-core contraction alone and return the raw per-CTA partials
extern "C" int cpg_debug_wgrad_tc(cpg_ctx* ctx, void* stream, int HP, const float* dg, const float* hs, int B, int L,
                                  float* part, int* nsplit, int dbg) {
    cpg::g_dbg_wgrad = dbg;
    int rc = cpg::launch_wgrad_hh_tc((cudaStream_t)stream, HP, dg, hs, B, L, ctx->sm_count, part, nsplit);
    if (rc) return rc;
    return cpg::check_launch("cpg_debug_wgrad_tc");
}
#endif
