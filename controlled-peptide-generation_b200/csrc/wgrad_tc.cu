// Recurrent weight gradients AND token-table gradients on the 5th-gen tensor cores.
//
// Both are contractions over the batch*time axis r = (b, s)  (K = B*L = 102,400 at B = 4096):
//     dW_hh[g][k]   = sum_r dgh[r][g] * h_prev[r][k]            (3 gate planes, N = H)
//     dT[v][pl][j]  = sum_r [tok_r == v] * dg[r][pl][j]          (4 planes, N = 32 one-hot columns)
// The range is split over one persistent CTA per SM; each CTA streams its rows through a
// TMA -> convert -> tcgen05.mma pipeline (4 stages of 16 rows) with all seven accumulator tiles
// (3 x [128 x H] + 4 x [128 x 32]) resident in TMEM for the whole range, so dg is read from HBM once.
//
// Operands are "MN-major" for UMMA: a row of `dg` ([B*L][4*HP]) holds the M index (gate column)
// contiguously, a row of `hs` ([B*L][HP]) the N index.  TMA boxes of {32 floats, 32 rows} with the
// 128-byte / 32-byte-atom swizzle (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) land exactly in the canonical
// MN-major SWIZZLE_128B_BASE32B layout -- the only layout tcgen05 accepts for MN-major tf32.
// h_prev is the hs row ABOVE (TMA row coordinate r-1); for rows with s == 0 the conversion warps
// overwrite it with h0 (decoder: [z;c]) or zeros (encoder).  The conversion warps also round every
// operand to tf32 (round-to-nearest: unbiased products; the tensor core itself truncates) and build the
// one-hot token tile in the same swizzled layout.
//
// Warp roles (320 threads): warp 4 = TMA producer, warp 5 = MMA issuer (+ TMEM owner), warps 0-3 and 6-9 =
// conversion per stage; warps 0-3 then run the TMEM -> global epilogue.
#include "ctx.h"
#ifndef CPG_EMU
#include "tc_common.cuh"

#ifdef CPG_GRU_TIMELINE
// developer-only probe (tools/wgrad_timeline.py): cycles each warp role of CTA 0 spends waiting / working
__device__ long long g_wg_tl[32];
extern "C" int cpg_debug_wgrad_timeline(long long* out) { return (int)cudaMemcpyFromSymbol(out, g_wg_tl, sizeof(g_wg_tl)); }
#define WG_T0() const long long _t0 = clock64()
#define WG_ACC(var) var += clock64() - _t0
#define WG_OUT(slot, v) do { if (blockIdx.x == 0) g_wg_tl[slot] = (v); } while (0)
#else
#define WG_T0() do { } while (0)
#define WG_ACC(var) do { } while (0)
#define WG_OUT(slot, v) do { } while (0)
#endif

namespace cpg {
int check_launch(const char* where);

constexpr int WT_RK = 16;                 // reduction rows per pipeline stage (two K = 8 MMAs per tile)
constexpr int WT_CHUNK = WT_RK * 128;     // bytes of one {32 floats x WT_RK rows} box
constexpr int WT_STAGES = 4;              // 4 x 42 KB: the TMA latency of three stages hides under the fourth's conversion + MMAs
constexpr int WT_A_BYTES = 4 * 4 * WT_CHUNK;      // 4 planes x 4 chunks (M = 128 lanes per plane)
constexpr int WT_B_BYTES = 4 * WT_CHUNK;          // h_prev: up to 128 columns
constexpr int WT_O_BYTES = WT_CHUNK;              // one-hot tokens: 32 columns
constexpr int WT_STAGE_BYTES = WT_A_BYTES + WT_B_BYTES + WT_O_BYTES;
constexpr size_t WT_SMEM = (size_t)WT_STAGES * WT_STAGE_BYTES + 1024;
constexpr int WT_TCOL = 384;              // TMEM column of the first dT accumulator (4 x 32 columns)

struct WgTcArgs {
    const uint8_t* tok;      // [B][L] tokens that fed this GRU
    const float* h0;         // [B][HP] initial hidden state, or null (zeros)
    float* part_w;           // [nsplit][3*HP][HP]
    float* part_t;           // [nsplit][V][4*HP]
    int nrows, L, V, reverse, rows_per_cta;
    int dg_rounded;          // dg arrives already rounded to tf32 (k_gru_bwd_tc): its boxes need no conversion pass
};

// one lane of a converged warp (keeps the single-thread TMA / MMA issue loops on the uniform datapath)
__device__ __forceinline__ bool wt_elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// byte offset of 16-byte unit u (columns 4u..4u+3) of row i inside one swizzled 4 KB box
__device__ __forceinline__ int wt_unit_off(int i, int u) { return i * 128 + ((((u >> 1) ^ (i & 3))) << 5) + ((u & 1) << 4); }

template <int HP>
__global__ void __launch_bounds__(320, 1)
k_wgrad_tc(const __grid_constant__ CUtensorMap tmap_dg, const __grid_constant__ CUtensorMap tmap_hs, WgTcArgs a) {
    constexpr int NCH = (HP + 31) / 32;                 // 32-float chunks holding valid columns
    constexpr int N_MMA = (HP + 15) / 16 * 16;          // UMMA N for the W_hh tiles
    extern __shared__ __align__(1024) unsigned char smem[];   // 1024-byte aligned window (SWIZZLE_128B atoms); used
                                                              // directly so that every access stays an LDS/STS
    __shared__ __align__(8) uint64_t bar_full[WT_STAGES], bar_conv[WT_STAGES], bar_empty[WT_STAGES], bar_done;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = a.L;
    const int r_begin = blockIdx.x * a.rows_per_cta;
    const int r_end = min(a.nrows, r_begin + a.rows_per_cta);
    const int n_iters = r_end > r_begin ? (r_end - r_begin + WT_RK - 1) / WT_RK : 0;

    if (warp == 5) {
        if (lane == 0) {
            for (int s = 0; s < WT_STAGES; ++s) {
                tc::mbar_init(&bar_full[s], 1);
                tc::mbar_init(&bar_conv[s], 256);
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
        // ---------------- TMA producer (whole warp converged, one elected lane issues)
        if (lane == 0) {
            tc::tma_prefetch_desc(&tmap_dg);
            tc::tma_prefetch_desc(&tmap_hs);
        }
        __syncwarp();
        long long w_empty = 0, w_issue = 0;
        const long long tstart = clock64();
        for (int it = 0; it < n_iters; ++it) {
            const int st = it % WT_STAGES, u = it / WT_STAGES;
            { WG_T0(); if (u >= 1) tc::mbar_wait(&bar_empty[st], (u - 1) & 1); WG_ACC(w_empty); }
            WG_T0();
            if (wt_elect_one()) {
                unsigned char* sa = smem + (size_t)st * WT_STAGE_BYTES;
                unsigned char* sb = sa + WT_A_BYTES;
                const int r0 = r_begin + it * WT_RK;
                tc::mbar_expect_tx(&bar_full[st], (4 * NCH + NCH) * WT_CHUNK);
#pragma unroll
                for (int pl = 0; pl < 4; ++pl)
#pragma unroll
                    for (int c = 0; c < NCH; ++c)
                        tc::tma_load_2d(sa + (pl * 4 + c) * WT_CHUNK, &tmap_dg, &bar_full[st], pl * HP + c * 32, r0);
#pragma unroll
                for (int c = 0; c < NCH; ++c)
                    tc::tma_load_2d(sb + c * WT_CHUNK, &tmap_hs, &bar_full[st], c * 32, r0 - 1);
            }
            __syncwarp();
            WG_ACC(w_issue);
        }
        if (lane == 0) { WG_OUT(0, w_empty); WG_OUT(1, w_issue); WG_OUT(2, clock64() - tstart); WG_OUT(3, (long long)n_iters); }
    } else if (warp == 5) {
        // ---------------- MMA issuer (whole warp converged, one elected lane issues)
        constexpr uint32_t idesc_w = tc::make_idesc_tf32(128, N_MMA, 1, 1);
        constexpr uint32_t idesc_t = tc::make_idesc_tf32(128, 32, 1, 1);
        // MN-major SWIZZLE_128B_BASE32B: next 32 floats of M/N one box further (LBO),
        // 4-row K groups 512 B apart (SBO); one K = 8 MMA spans two such groups
        constexpr uint32_t lbo = WT_CHUNK, sbo = 512u;
        long long w_conv = 0, w_mma = 0;
        const long long tstart = clock64();
        const uint32_t s0 = tc::smem_u32(smem);
        for (int it = 0; it < n_iters; ++it) {
            const int st = it % WT_STAGES, u = it / WT_STAGES;
            { WG_T0(); tc::mbar_wait(&bar_conv[st], u & 1); WG_ACC(w_conv); }
            tc::tc_fence_after();
            WG_T0();
            if (wt_elect_one()) {
                const uint32_t sa = s0 + (uint32_t)st * WT_STAGE_BYTES;
                const uint32_t sb = sa + WT_A_BYTES, so = sb + WT_B_BYTES;
#pragma unroll
                for (int ks = 0; ks < WT_RK / 8; ++ks) {
                    const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
                    const uint64_t db = tc::make_smem_desc(sb + ks * 1024, lbo, sbo, 1);
                    const uint64_t dt_ = tc::make_smem_desc(so + ks * 1024, lbo, sbo, 1);
#pragma unroll
                    for (int pl = 0; pl < 4; ++pl) {
                        const uint64_t da = tc::make_smem_desc(sa + pl * 4 * WT_CHUNK + ks * 1024, lbo, sbo, 1);
                        if (pl != 2) tc::umma_tf32(tmem_d + (pl == 3 ? 2 : pl) * 128, da, db, idesc_w, acc);   // dr, dz, dhn
                        tc::umma_tf32(tmem_d + WT_TCOL + pl * 32, da, dt_, idesc_t, acc);
                    }
                }
                tc::umma_commit(&bar_empty[st]);          // stage reusable once these MMAs have read it
            }
            __syncwarp();
            WG_ACC(w_mma);
        }
        if (wt_elect_one()) tc::umma_commit(&bar_done);
        __syncwarp();
        if (lane == 0) { WG_OUT(4, w_conv); WG_OUT(5, w_mma); WG_OUT(6, clock64() - tstart); }
    } else {
        // ---------------- conversion (in place), 256 threads = warps 0-3 and 6-9 (two warps per scheduler: the
        // single-warp version was latency-bound and paced the whole pipeline).  Converter ci -> 16-byte unit
        // (ci & 7) of row (ci >> 3) & 15 of every box; group ci >> 7 takes the boxes of its parity.  Group 0 also
        // rounds h_prev (and rewrites it for first steps), group 1 builds the one-hot token tile.
        const int ci = warp < 4 ? tid : tid - 64;
        const int grp = ci >> 7, slot = ci & 127, row = slot >> 3, un = slot & 7;
        auto rnd4 = [](float4 v) {          // tf32 round-to-nearest = add half an ulp; the tensor core drops the low bits
            v.x = __uint_as_float(__float_as_uint(v.x) + 0x1000u); v.y = __uint_as_float(__float_as_uint(v.y) + 0x1000u);
            v.z = __uint_as_float(__float_as_uint(v.z) + 0x1000u); v.w = __uint_as_float(__float_as_uint(v.w) + 0x1000u);
            return v;
        };
        static_assert(WT_RK == 16, "one row per converter slot");
        long long w_full = 0, w_work = 0;
        const long long tstart = clock64();
        for (int it = 0; it < n_iters; ++it) {
            const int st = it % WT_STAGES, u_it = it / WT_STAGES;
            const int r0 = r_begin + it * WT_RK;
            const int r = r0 + row;
            const bool dead = r >= r_end;                     // rows of the next CTA's range (or past the end)
            const int rr = dead ? r_begin : r;
            const int b = rr / L, s = rr % L;
            // token of this row: fetched before the wait so that its latency hides under the TMA's
            int tk = 0;
            if (grp == 1) {
                const int t = a.reverse ? (L - 1 - s) : s;
                tk = a.tok[(size_t)b * L + t];
            }
            { WG_T0(); tc::mbar_wait(&bar_full[st], u_it & 1); WG_ACC(w_full); }
            WG_T0();
            unsigned char* sa = smem + (size_t)st * WT_STAGE_BYTES;
            unsigned char* sb = sa + WT_A_BYTES;
            unsigned char* so = sb + WT_B_BYTES;
            // (rows past the end of the reduction are zero-filled by the TMA itself; rows past this CTA's range
            //  only exist when the range is not a multiple of the tile, i.e. never with pre-rounded dg)
            if (!a.dg_rounded || dead)
#pragma unroll
            for (int pl = 0; pl < 4; ++pl)
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    if (((pl * NCH + c) & 1) != grp) continue;
                    float4* p = reinterpret_cast<float4*>(sa + (pl * 4 + c) * WT_CHUNK) + slot;   // box-linear: swizzle agnostic
                    *p = dead ? make_float4(0.f, 0.f, 0.f, 0.f) : rnd4(*p);
                }
            if (grp == 0) {
                if (s == 0) {
                    // h_prev of a first step is h0, not the row above: rewrite it (swizzle-aware addressing)
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        const int col = c * 32 + un * 4;
                        if (a.h0 != nullptr && col < HP) v = ld4(a.h0 + (size_t)b * HP + col);
                        *reinterpret_cast<float4*>(sb + c * WT_CHUNK + wt_unit_off(row, un)) = rnd4(v);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        float4* p = reinterpret_cast<float4*>(sb + c * WT_CHUNK) + slot;
                        *p = rnd4(*p);
                    }
                }
            } else {
                // one-hot row of the token that fed step s (time L-1-s for the reverse direction)
                float4 oh = make_float4(0.f, 0.f, 0.f, 0.f);
                if ((tk >> 2) == un) {
                    const int k = tk & 3;
                    oh.x = k == 0 ? 1.f : 0.f; oh.y = k == 1 ? 1.f : 0.f; oh.z = k == 2 ? 1.f : 0.f; oh.w = k == 3 ? 1.f : 0.f;
                }
                *reinterpret_cast<float4*>(so + wt_unit_off(row, un)) = oh;
            }
            tc::fence_proxy_async();                      // generic-proxy writes -> visible to the tensor core
            tc::mbar_arrive(&bar_conv[st]);
            WG_ACC(w_work);
        }
        if (tid == 0) { WG_OUT(8, w_full); WG_OUT(9, w_work); WG_OUT(10, clock64() - tstart); }
        const long long tepi = clock64();
        // ---------------- epilogue (warps 0-3): TMEM -> per-CTA partials
        if (warp < 4) {
            if (n_iters > 0) {
                tc::mbar_wait(&bar_done, 0);
                tc::tc_fence_after();
            }
            float* out_w = a.part_w + (size_t)blockIdx.x * 3 * HP * HP;
            float* out_t = a.part_t + (size_t)blockIdx.x * a.V * 4 * HP;
            const uint32_t lane_addr = tmem_d + ((uint32_t)(warp * 32) << 16);
            for (int g = 0; g < 3; ++g) {
#pragma unroll 1
                for (int c = 0; c < NCH; ++c) {
                    float v[32];
                    if (n_iters > 0) {
                        tc::tmem_ld_32x32(lane_addr + (uint32_t)(g * 128 + c * 32), v);
                    } else {
#pragma unroll
                        for (int q = 0; q < 32; ++q) v[q] = 0.f;
                    }
                    if (tid < HP) {
                        float* o = out_w + ((size_t)g * HP + tid) * HP + c * 32;      // 16-byte aligned: HP % 4 == 0
#pragma unroll
                        for (int q = 0; q < 32; q += 4)
                            if (c * 32 + q < HP) st4(o + q, make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]));
                    }
                }
            }
#pragma unroll 1
            for (int pl = 0; pl < 4; ++pl) {
                float v[32];
                if (n_iters > 0) {
                    tc::tmem_ld_32x32(lane_addr + (uint32_t)(WT_TCOL + pl * 32), v);
                } else {
#pragma unroll
                    for (int q = 0; q < 32; ++q) v[q] = 0.f;
                }
                if (tid < HP) {
#pragma unroll
                    for (int q = 0; q < 32; ++q)
                        if (q < a.V) out_t[(size_t)q * 4 * HP + pl * HP + tid] = v[q];
                }
            }
        }
#ifdef CPG_GRU_TIMELINE
        if (tid == 0) WG_OUT(11, clock64() - tepi);
#else
        (void)tepi;
#endif
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 5) tc::tmem_dealloc<512>(tmem_d);
}

int wgrad_tc_splits(int sm_count) { return sm_count > 0 ? sm_count : 1; }

// Fills part_w ([nsplit][3*HP][HP]) and part_t ([nsplit][V][4*HP]); the caller reduces them.
int launch_wgrad_tc(cudaStream_t s, int HP, const float* dg, const float* hs, const float* h0, const uint8_t* tok,
                    int reverse, int B, int L, int V, int sm_count, float* part_w, float* part_t, int* nsplit_out, int dg_rounded) {
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
    WgTcArgs a;
    a.tok = tok; a.h0 = h0; a.part_w = part_w; a.part_t = part_t;
    a.nrows = nrows; a.L = L; a.V = V; a.reverse = reverse; a.rows_per_cta = rpc; a.dg_rounded = dg_rounded;
    if (HP == ENC_H) {
        auto kfn = k_wgrad_tc<ENC_H>;
        static bool once = false;
        if (!once) { cudaFuncSetAttribute((const void*)kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WT_SMEM); once = true; }
        CPG_LAUNCH_NAMED("k_wgrad_tc_enc", kfn, nsplit, 320, WT_SMEM, s, tm_dg, tm_hs, a);
    } else {
        auto kfn = k_wgrad_tc<DEC_HP>;
        static bool once = false;
        if (!once) { cudaFuncSetAttribute((const void*)kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WT_SMEM); once = true; }
        CPG_LAUNCH_NAMED("k_wgrad_tc_dec", kfn, nsplit, 320, WT_SMEM, s, tm_dg, tm_hs, a);
    }
    return CPG_OK;
}

}  // namespace cpg
#endif  // CPG_EMU
