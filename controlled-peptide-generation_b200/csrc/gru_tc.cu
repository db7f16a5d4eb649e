// GRU recurrences (forward with activation stash, and BPTT) on the 5th-gen tensor cores, with
// fp32-grade accuracy.  Same math and the same HBM stash layouts as the fp32 SIMT kernels of gru.cu
// (models/encoder.py:25-30,42 and models/decoder.py:40,77; gate order r,z,n; h' = (1-z) n + z h),
// so either flavour of forward pairs with either flavour of backward.
//
// Formulation.  The batch is the N side of the MMA and the weights are the M side ("transposed"):
//     forward   G^T[3H x nb] = W_hh[3H x H]   . h^T[H x nb]          (M tiles of 128 gate rows)
//     backward  dh^T[H x nb] = W_hh^T[H x 3H] . (dr,dz,dhn)^T[3H x nb]
// N = 32 batch rows per MMA, so B = 4096 gives 128 CTAs whatever the hidden size (the row-major
// variant is stuck at M = 128 rows per tile, i.e. 32..64 CTAs on 148 SMs).  tcgen05.mma kind::f16 with
// bf16 operands and fp32 accumulation in TMEM; every fp32 operand x is split x = x1 + x2 (two bf16
// terms, 16 mantissa bits) and the three products x1 w1 + x1 w2 + x2 w1 are accumulated: relative
// error 2^-16 per term, well inside the 1e-4 parity bar of the path (a single bf16/tf32 pass is not).
//
// One CTA = NSUB sub-tiles of 32 batch rows of one direction, all L steps:
//   warp NW (last) : TMEM owner; one elected thread issues the MMAs of a sub-tile as soon as that
//                    sub-tile's operand tile of the previous step is complete (mbarrier bar_x)
//   warps 0..NW-1  : per sub-tile and step
//        phase 1  tcgen05.ld (lane = gate row, 32 batch columns) -> P[batch][gate] in shared memory
//        phase 2  thread = (batch row, 4 consecutive hidden units): gate math, fp32 stash to HBM with
//                 row-contiguous float4 stores, next operand tile (bf16 split, K-major core matrices)
//   With NSUB = 2 the MMAs of one sub-tile run under the epilogue of the other (ping-pong).
// W_hh (both bf16 terms) stays in shared memory for all L steps; operand tiles use the no-swizzle
// K-major core-matrix layout with a 16-byte pad on the K stride (bank-conflict-free 8-byte stores).
#include "ctx.h"
#ifndef CPG_EMU
#include <cuda_bf16.h>
#include "tc_common.cuh"

#ifdef CPG_GRU_TIMELINE
// developer-only: clock64() stamps of one CTA / one step (tools/gru_timeline.py); never in the product build
__device__ long long g_gru_tl[64];
#define CPG_TL(slot) do { if (blockIdx.x == 0 && blockIdx.y == 0 && s == 12) g_gru_tl[slot] = clock64(); } while (0)
#define CPG_TL0(slot) do { if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) g_gru_tl[slot] = clock64(); } while (0)
extern "C" int cpg_debug_gru_timeline(long long* out) { return (int)cudaMemcpyFromSymbol(out, g_gru_tl, sizeof(g_gru_tl)); }
#else
#define CPG_TL(slot) do { } while (0)
#define CPG_TL0(slot) do { } while (0)
#endif

namespace cpg {
int check_launch(const char* where);

namespace {
constexpr int NBS = 32;                               // batch rows per sub-tile = MMA N
constexpr int X_LBO = (NBS / 8) * 128 + 16;           // K-adjacent core matrices of an operand tile (padded)
constexpr int X_SBO = 128;                            // N-adjacent core matrices
constexpr int W_SBO = 128;                            // M-adjacent core matrices of the weight tile

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// same, A operand resident in tensor memory (lane = M row, one 32-bit column = two consecutive K elements)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// kind::f16 with bf16 operands, fp32 accumulate, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// one lane of a converged warp (the form the compiler keeps on the uniform datapath)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
template <int N>
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory"); }

// (x0, x1) -> packed bf16 leading terms (x0 in the low half) and packed bf16 remainders
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const float f0 = __uint_as_float(hi << 16), f1 = __uint_as_float(hi & 0xffff0000u);
    const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - f0, x1 - f1);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// 4 fp32 -> 4 bf16 leading terms + 4 bf16 remainders
__device__ __forceinline__ void split4(const float (&x)[4], uint2& hi, uint2& lo) {
    split2(x[0], x[1], hi.x, lo.x);
    split2(x[2], x[3], hi.y, lo.y);
}
__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& lo) {
    split2(x[0], x[1], hi.x, lo.x);
    split2(x[2], x[3], hi.y, lo.y);
    split2(x[4], x[5], hi.z, lo.z);
    split2(x[6], x[7], hi.w, lo.w);
}
// Gate non-linearities straight on the SFU (ex2.approx + rcp.approx, flush-to-zero forms: no denormal
// fix-up code around them); absolute error <= ~6e-7 like the SIMT kernels' versions.
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_ftz(1.0f + ex2_ftz(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_fast(float x) { return fmaf(-2.0f, rcp_ftz(1.0f + ex2_ftz(2.8853900817779268f * x)), 1.0f); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// (operand-tile term, weight term) of the three accumulated products
__device__ constexpr int XS[3] = {0, 0, 1};
__device__ constexpr int WS[3] = {0, 1, 0};

// ------------------------------------------------------------------------------------- forward
template <int HP_, int KP_, int NSUB_, bool DEC_>
struct FwdCfg {
    static constexpr int HP = HP_, KP = KP_, NSUB = NSUB_;
    static constexpr bool DEC = DEC_;
    static constexpr int G3 = 3 * HP;                       // gate rows (M)
    static constexpr int MT = (G3 + 127) / 128;             // M tiles
    static constexpr int NQ = HP / 4;                       // unit quads per row
    static constexpr int ITEMS = 2;                         // (row, quad) items per thread and sub-tile
    static constexpr int NT_EPI = NBS * NQ / ITEMS;         // 320 | 416
    static constexpr int NW_EPI = NT_EPI / 32;
    static constexpr int NTHREADS = NT_EPI + 32;
    static constexpr int KC = KP / 8, KSTEPS = KP / 16;
    static constexpr int X_SPLIT = KC * X_LBO;
    static constexpr int P_FLOATS = NBS * G3;
    // tensor memory: accumulators [NSUB][MT] x 32 columns, then W_hh as the A operand: [MT][2 terms][KP/2] columns
    static constexpr int WCOL0 = NSUB * MT * NBS;
    static constexpr int WCOLS = KP / 2;
    static constexpr uint32_t TMEM_COLS = 512;
    static_assert(WCOL0 + MT * 2 * WCOLS <= 512, "TMEM columns");
    static_assert(NT_EPI % 32 == 0 && NT_EPI % NQ == 0 && MT * 4 <= NW_EPI, "thread mapping");
    static_assert(HP % 8 == 0 && KP % 16 == 0 && G3 % 8 == 0, "core-matrix geometry");
    static size_t smem_bytes(int V, int L) {
        return (size_t)NSUB * 2 * X_SPLIT + (size_t)NSUB * P_FLOATS * 4 + (DEC ? 0 : (size_t)V * G3 * 4) +
               (size_t)NSUB * NBS * L + 128;
    }
};

struct FwdArgs {
    const uint8_t* tok;        // [B][L]
    const float* table[2];     // [V][3*HP] per direction
    const float* whh[2];       // [3*HP][HP] natural (zero padded for the decoder)
    const float* bhn[2];       // [HP]
    const float* rowbias;      // decoder: [B][3*HP]
    const float* h0;           // decoder: [B][HP]
    float* hs[2];              // [B][L][HP] by step (nullable)
    float* gates[2];           // [B][L][4][HP] (nullable)
    float* hfin;               // encoder: [B][2*HP]
    int B, L, V;
};

template <class C>
__global__ void __launch_bounds__(C::NTHREADS, 1)
k_gru_fwd_tc(FwdArgs a) {
    constexpr int HP = C::HP, G3 = C::G3, MT = C::MT, NQ = C::NQ, NSUB = C::NSUB, KC = C::KC;
    extern __shared__ __align__(1024) unsigned char smem[];   // used directly: keeps every access an LDS/STS
    unsigned char* Xb = smem;                                            // [NSUB][2 terms][KC][X_LBO]
    float* Pb = reinterpret_cast<float*>(Xb + NSUB * 2 * C::X_SPLIT);    // [NSUB][NBS][G3]
    float* tab = Pb + NSUB * C::P_FLOATS;                                // encoder: [V][G3]
    uint8_t* toks = reinterpret_cast<uint8_t*>(tab + (C::DEC ? 0 : a.V * G3));   // [NSUB*NBS][L]
    __shared__ __align__(8) uint64_t bar_x[NSUB], bar_d[NSUB];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dir = blockIdx.y;
    const int row0 = blockIdx.x * (NSUB * NBS);
    const int B = a.B, L = a.L;
    CPG_TL0(50);

    // ---- one-time setup
    {
        for (int idx = tid; idx < NSUB * NBS * KC; idx += C::NTHREADS) {
            const int bb = idx / KC, kc = idx % KC, sub = bb / NBS, b = bb % NBS;
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = 0.f;
            if (C::DEC && kc * 8 + 8 <= HP) {
                const int row = min(row0 + bb, B - 1);
                const float4 f0 = ldg4(a.h0 + (size_t)row * HP + kc * 8), f1 = ldg4(a.h0 + (size_t)row * HP + kc * 8 + 4);
                x[0] = f0.x; x[1] = f0.y; x[2] = f0.z; x[3] = f0.w; x[4] = f1.x; x[5] = f1.y; x[6] = f1.z; x[7] = f1.w;
            }
            uint4 hi, lo;
            split8(x, hi, lo);
            const int off = kc * X_LBO + (b >> 3) * X_SBO + (b & 7) * 16;
            *reinterpret_cast<uint4*>(Xb + (sub * 2 + 0) * C::X_SPLIT + off) = hi;
            *reinterpret_cast<uint4*>(Xb + (sub * 2 + 1) * C::X_SPLIT + off) = lo;
        }
        if (!C::DEC)
            for (int i = tid; i < a.V * G3; i += C::NTHREADS) tab[i] = (dir ? a.table[1] : a.table[0])[i];
        for (int i = tid; i < NSUB * NBS * L; i += C::NTHREADS) {
            const int row = min(row0 + i / L, B - 1);
            toks[i] = a.tok[(size_t)row * L + i % L];
        }
    }
    if (warp == C::NW_EPI) {
        if (lane == 0) {
            for (int i = 0; i < NSUB; ++i) {
                tc::mbar_init(&bar_x[i], C::NT_EPI);
                tc::mbar_init(&bar_d[i], 1);
            }
            tc::fence_barrier_init();
        }
        __syncwarp();
        tc::tmem_alloc<C::TMEM_COLS>(&tmem_slot);
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = tmem_slot;

    // W_hh -> tensor memory (A operand of every MMA of this CTA): warp q fills lane quadrant q of each M tile,
    // lane = gate row, 16 K elements (8 packed columns) per store, leading and remainder bf16 terms side by side
    if (warp < 4) {
        const float* whh = (dir ? a.whh[1] : a.whh[0]);
#pragma unroll 1
        for (int t = 0; t < MT; ++t) {
            const int m = t * 128 + warp * 32 + lane;
            const uint32_t lane_addr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(C::WCOL0 + t * 2 * C::WCOLS);
#pragma unroll 1
            for (int ks = 0; ks < C::KSTEPS; ++ks) {
                float x[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) x[e] = 0.f;
                if (m < G3) {
#pragma unroll
                    for (int e4 = 0; e4 < 4; ++e4) {
                        const int k = ks * 16 + e4 * 4;
                        if (k + 4 <= HP) {
                            const float4 f = ldg4(whh + (size_t)m * HP + k);
                            x[e4 * 4 + 0] = f.x; x[e4 * 4 + 1] = f.y; x[e4 * 4 + 2] = f.z; x[e4 * 4 + 3] = f.w;
                        }
                    }
                }
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) split2(x[2 * e], x[2 * e + 1], hi[e], lo[e]);
                tmem_st_32x8(lane_addr + (uint32_t)(ks * 8), hi);
                tmem_st_32x8(lane_addr + (uint32_t)(C::WCOLS + ks * 8), lo);
            }
        }
        tmem_st_wait();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    CPG_TL0(51);

    if (warp == C::NW_EPI) {
        // ---------------- MMA issuer (whole warp converged; one elected lane issues)
        constexpr uint32_t idesc = make_idesc_bf16(128, NBS);
        const uint32_t x0 = tc::smem_u32(Xb);
        for (int s = 0; s < L; ++s) {
#pragma unroll
            for (int sub = 0; sub < NSUB; ++sub) {
                if (s > 0) {
                    tc::mbar_wait(&bar_x[sub], (s - 1) & 1);
                    tc::tc_fence_after();
                }
                if (elect_one()) {
                    CPG_TL(0 + sub);
#pragma unroll
                    for (int t = 0; t < MT; ++t) {
                        uint32_t acc = 0;
#pragma unroll
                        for (int p = 0; p < 3; ++p) {
#pragma unroll
                            for (int ks = 0; ks < C::KSTEPS; ++ks) {
                                const uint32_t ta = tmem_d + (uint32_t)(C::WCOL0 + (t * 2 + WS[p]) * C::WCOLS + ks * 8);
                                const uint64_t db = tc::make_smem_desc(x0 + (sub * 2 + XS[p]) * C::X_SPLIT + ks * 2 * X_LBO,
                                                                       X_LBO, X_SBO, 0);
                                umma_bf16_ts(tmem_d + (uint32_t)((sub * MT + t) * NBS), ta, db, idesc, acc);
                                acc = 1;
                            }
                        }
                    }
                    tc::umma_commit(&bar_d[sub]);
                }
                __syncwarp();
            }
        }
        __syncwarp();
    } else {
        // ---------------- epilogue
        const int quad = tid % NQ, j0 = quad * 4, bq = tid / NQ;        // item it -> batch row bq + 16 it
        const float4 bhn4 = ldg4((dir ? a.bhn[1] : a.bhn[0]) + j0);
        const float bhn[4] = {bhn4.x, bhn4.y, bhn4.z, bhn4.w};
        float hprev[NSUB][C::ITEMS][4];
        float rb[C::DEC ? C::ITEMS : 1][3][4];
#pragma unroll
        for (int sub = 0; sub < NSUB; ++sub)
#pragma unroll
            for (int it = 0; it < C::ITEMS; ++it) {
                const int row = min(row0 + sub * NBS + bq + 16 * it, B - 1);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (C::DEC) {
                    v = ldg4(a.h0 + (size_t)row * HP + j0);
#pragma unroll
                    for (int g = 0; g < 3; ++g) {
                        const float4 r4 = ldg4(a.rowbias + (size_t)row * G3 + g * HP + j0);
                        rb[C::DEC ? it : 0][g][0] = r4.x; rb[C::DEC ? it : 0][g][1] = r4.y;
                        rb[C::DEC ? it : 0][g][2] = r4.z; rb[C::DEC ? it : 0][g][3] = r4.w;
                    }
                }
                hprev[sub][it][0] = v.x; hprev[sub][it][1] = v.y; hprev[sub][it][2] = v.z; hprev[sub][it][3] = v.w;
            }
        const float* tabg = C::DEC ? (dir ? a.table[1] : a.table[0]) : tab;
        float* hs_g = (dir ? a.hs[1] : a.hs[0]);
        float* gates_g = (dir ? a.gates[1] : a.gates[0]);

        for (int s = 0; s < L; ++s) {
            const int t = dir ? (L - 1 - s) : s;
#pragma unroll
            for (int sub = 0; sub < NSUB; ++sub) {
                float* P = Pb + sub * C::P_FLOATS;
                // input-side pre-activations do not depend on the MMA: fetch them before waiting on it
                float4 tin[C::ITEMS][3];
#pragma unroll
                for (int it = 0; it < C::ITEMS; ++it) {
                    const int tk = toks[(sub * NBS + bq + 16 * it) * L + t];
                    const float* trow = tabg + tk * G3 + j0;
                    if (C::DEC) { tin[it][0] = ldg4(trow); tin[it][1] = ldg4(trow + HP); tin[it][2] = ldg4(trow + 2 * HP); }
                    else { tin[it][0] = ld4(trow); tin[it][1] = ld4(trow + HP); tin[it][2] = ld4(trow + 2 * HP); }
                }
                if (lane == 0 && (warp == 0 || warp == C::NW_EPI - 1)) CPG_TL(8 + 16 * sub + (warp ? 8 : 0));
                tc::mbar_wait(&bar_d[sub], s & 1);
                tc::tc_fence_after();
                if (lane == 0 && (warp == 0 || warp == C::NW_EPI - 1)) CPG_TL(9 + 16 * sub + (warp ? 8 : 0));
                // phase 1: accumulator (lane = gate row, 32 batch columns) -> P[batch][gate]
                if (warp < MT * 4) {
                    const int tq = warp >> 2, q = warp & 3;
                    float v[32];
                    tc::tmem_ld_32x32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)((sub * MT + tq) * NBS), v);
                    const int m = tq * 128 + q * 32 + lane;
                    if (m < G3) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) P[c * G3 + m] = v[c];
                    }
                }
                tc::tc_fence_before();
                if (lane == 0 && (warp == 0 || warp == C::NW_EPI - 1)) CPG_TL(10 + 16 * sub + (warp ? 8 : 0));
                epi_bar_sync<C::NT_EPI>();
                if (lane == 0 && (warp == 0 || warp == C::NW_EPI - 1)) CPG_TL(11 + 16 * sub + (warp ? 8 : 0));
                // phase 2: gates for (row, 4 units)
#pragma unroll
                for (int it = 0; it < C::ITEMS; ++it) {
                    const int b = bq + 16 * it;
                    const int row = row0 + sub * NBS + b;
                    const float4 pr = ld4(P + b * G3 + j0), pz = ld4(P + b * G3 + HP + j0), pn = ld4(P + b * G3 + 2 * HP + j0);
                    const float4 tr = tin[it][0], tz = tin[it][1], tn = tin[it][2];
                    float gr[4] = {tr.x + pr.x, tr.y + pr.y, tr.z + pr.z, tr.w + pr.w};
                    float gz[4] = {tz.x + pz.x, tz.y + pz.y, tz.z + pz.z, tz.w + pz.w};
                    float gn[4] = {tn.x, tn.y, tn.z, tn.w};
                    const float pnv[4] = {pn.x, pn.y, pn.z, pn.w};
                    float rr[4], zz[4], nn[4], hh[4], hn[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (C::DEC) {
                            gr[e] += rb[C::DEC ? it : 0][0][e];
                            gz[e] += rb[C::DEC ? it : 0][1][e];
                            gn[e] += rb[C::DEC ? it : 0][2][e];
                        }
                        rr[e] = sigmoid_fast(gr[e]);
                        zz[e] = sigmoid_fast(gz[e]);
                        hh[e] = pnv[e] + bhn[e];
                        nn[e] = tanh_fast(gn[e] + rr[e] * hh[e]);
                        hn[e] = (1.0f - zz[e]) * nn[e] + zz[e] * hprev[sub][it][e];
                        hprev[sub][it][e] = hn[e];
                    }
                    uint2 hi, lo;
                    split4(hn, hi, lo);
                    const int off = (j0 >> 3) * X_LBO + (b >> 3) * X_SBO + (b & 7) * 16 + (j0 & 7) * 2;
                    *reinterpret_cast<uint2*>(Xb + (sub * 2 + 0) * C::X_SPLIT + off) = hi;
                    *reinterpret_cast<uint2*>(Xb + (sub * 2 + 1) * C::X_SPLIT + off) = lo;
                    if (tid == 0 && sub == 0) CPG_TL(40 + 2 * it);
                    if (row < B) {
                        const size_t bs = (size_t)row * L + s;
                        if (hs_g != nullptr) st4(hs_g + bs * HP + j0, make_float4(hn[0], hn[1], hn[2], hn[3]));
                        if (gates_g != nullptr) {
                            float* g = gates_g + bs * 4 * HP + j0;
                            st4(g, make_float4(rr[0], rr[1], rr[2], rr[3]));
                            st4(g + HP, make_float4(zz[0], zz[1], zz[2], zz[3]));
                            st4(g + 2 * HP, make_float4(nn[0], nn[1], nn[2], nn[3]));
                            st4(g + 3 * HP, make_float4(hh[0], hh[1], hh[2], hh[3]));
                        }
                        if (!C::DEC && s == L - 1)
                            st4(a.hfin + (size_t)row * (2 * HP) + dir * HP + j0, make_float4(hn[0], hn[1], hn[2], hn[3]));
                    }
                    if (tid == 0 && sub == 0) CPG_TL(41 + 2 * it);
                }
                if (lane == 0 && (warp == 0 || warp == C::NW_EPI - 1)) CPG_TL(12 + 16 * sub + (warp ? 8 : 0));
                tc::fence_proxy_async();                         // operand-tile stores -> visible to the tensor core
                tc::mbar_arrive(&bar_x[sub]);
                if (lane == 0 && (warp == 0 || warp == C::NW_EPI - 1)) CPG_TL(13 + 16 * sub + (warp ? 8 : 0));
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    CPG_TL0(52);
    if (warp == C::NW_EPI) tc::tmem_dealloc<C::TMEM_COLS>(tmem_d);
}

// ------------------------------------------------------------------------------------ backward
// Per step (reverse step order), as in gru.cu:
//   dht = dh + dh_out[s];  dn = dht (1-z); dz = dht (h_prev - n); carry = dht z
//   dn_pre = dn (1-n^2); dr = dn_pre hn; dhn = dn_pre r; dr_pre = dr r (1-r); dz_pre = dz z (1-z)
//   dh = carry + W_hh^T [dr_pre, dz_pre, dhn]
template <int HP_, int NSUB_, bool DEC_>
struct BwdCfg {
    static constexpr int HP = HP_, NSUB = NSUB_;
    static constexpr bool DEC = DEC_;
    static constexpr int K3 = 3 * HP;
    static constexpr int KPAD = (K3 + 15) / 16 * 16;        // 240 | 320
    static constexpr int KC = KPAD / 8, KSTEPS = KPAD / 16;
    static constexpr int NQ = HP / 4;
    static constexpr int ITEMS = 2;
    static constexpr int NT_EPI = NBS * NQ / ITEMS;
    static constexpr int NW_EPI = NT_EPI / 32;
    static constexpr int NTHREADS = NT_EPI + 32;
    static constexpr int X_SPLIT = KC * X_LBO;
    static constexpr int P_FLOATS = NBS * HP;
    // tensor memory: accumulators [NSUB] x 32 columns, then W_hh^T as the A operand: [2 terms][KPAD/2] columns
    static constexpr int WCOL0 = NSUB * NBS;
    static constexpr int WCOLS = KPAD / 2;
    static constexpr uint32_t TMEM_COLS = 512;
    static_assert(WCOL0 + 2 * WCOLS <= 512, "TMEM columns");
    static_assert(HP <= 128 && NT_EPI % NQ == 0 && NW_EPI >= 4, "thread mapping");
    static size_t smem_bytes() { return (size_t)NSUB * 2 * X_SPLIT + (size_t)NSUB * P_FLOATS * 4 + 128; }
};

struct BwdArgs {
    const float* whh[2];       // [3*HP][HP] natural
    const float* hs[2];        // [B][L][HP]
    const float* gates[2];     // [B][L][4][HP]
    const float* h0;           // decoder: [B][HP] (null = zeros)
    const float* dh_out;       // decoder: [B][L][HP]
    const float* dh_fin;       // encoder: [B][2*HP]
    float* dg[2];              // [B][L][4][HP]
    float* dh0;                // decoder: [B][HP]
    float* drow;               // decoder: [B][3*HP]
    int B, L;
};

template <class C>
__global__ void __launch_bounds__(C::NTHREADS, 1)
k_gru_bwd_tc(BwdArgs a) {
    constexpr int HP = C::HP, NQ = C::NQ, NSUB = C::NSUB, KC = C::KC, K3 = C::K3;
    extern __shared__ __align__(1024) unsigned char smem[];   // used directly: keeps every access an LDS/STS
    unsigned char* Xb = smem;                                            // [NSUB][2 terms][KC][X_LBO]
    float* Pb = reinterpret_cast<float*>(Xb + NSUB * 2 * C::X_SPLIT);    // [NSUB][NBS][HP]
    __shared__ __align__(8) uint64_t bar_x[NSUB], bar_d[NSUB];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dir = blockIdx.y;
    const int row0 = blockIdx.x * (NSUB * NBS);
    const int B = a.B, L = a.L;

    // ---- one-time setup: operand tiles zeroed (K padding stays 0)
    {
        for (int i = tid; i < NSUB * 2 * C::X_SPLIT / 16; i += C::NTHREADS) reinterpret_cast<uint4*>(Xb)[i] = make_uint4(0, 0, 0, 0);
    }
    if (warp == C::NW_EPI) {
        if (lane == 0) {
            for (int i = 0; i < NSUB; ++i) {
                tc::mbar_init(&bar_x[i], C::NT_EPI);
                tc::mbar_init(&bar_d[i], 1);
            }
            tc::fence_barrier_init();
        }
        __syncwarp();
        tc::tmem_alloc<C::TMEM_COLS>(&tmem_slot);
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = tmem_slot;

    // W_hh^T -> tensor memory: lane = hidden unit j, K index = gate row k, A[j][k] = W_hh[k][j]
    if (warp < 4) {
        const float* whh = (dir ? a.whh[1] : a.whh[0]);
        const int j = warp * 32 + lane;
        const uint32_t lane_addr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)C::WCOL0;
#pragma unroll 1
        for (int ks = 0; ks < C::KSTEPS; ++ks) {
            float x[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int k = ks * 16 + e;
                x[e] = (j < HP && k < K3) ? __ldg(whh + (size_t)k * HP + j) : 0.f;
            }
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) split2(x[2 * e], x[2 * e + 1], hi[e], lo[e]);
            tmem_st_32x8(lane_addr + (uint32_t)(ks * 8), hi);
            tmem_st_32x8(lane_addr + (uint32_t)(C::WCOLS + ks * 8), lo);
        }
        tmem_st_wait();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();

    if (warp == C::NW_EPI) {
        // ---------------- MMA issuer: iteration i handles step s = L-1-i
        constexpr uint32_t idesc = make_idesc_bf16(128, NBS);
        const uint32_t x0 = tc::smem_u32(Xb);
        for (int i = 0; i < L; ++i) {
#pragma unroll
            for (int sub = 0; sub < NSUB; ++sub) {
                tc::mbar_wait(&bar_x[sub], i & 1);
                tc::tc_fence_after();
                if (elect_one()) {
                    uint32_t acc = 0;
#pragma unroll
                    for (int p = 0; p < 3; ++p) {
#pragma unroll
                        for (int ks = 0; ks < C::KSTEPS; ++ks) {
                            const uint32_t ta = tmem_d + (uint32_t)(C::WCOL0 + WS[p] * C::WCOLS + ks * 8);
                            const uint64_t db = tc::make_smem_desc(x0 + (sub * 2 + XS[p]) * C::X_SPLIT + ks * 2 * X_LBO, X_LBO, X_SBO, 0);
                            umma_bf16_ts(tmem_d + (uint32_t)(sub * NBS), ta, db, idesc, acc);
                            acc = 1;
                        }
                    }
                    tc::umma_commit(&bar_d[sub]);
                }
                __syncwarp();
            }
        }
        __syncwarp();
    } else {
        // ---------------- epilogue
        const int quad = tid % NQ, j0 = quad * 4, bq = tid / NQ;
        const float* hs_g = (dir ? a.hs[1] : a.hs[0]);
        const float* gates_g = (dir ? a.gates[1] : a.gates[0]);
        float* dg_g = (dir ? a.dg[1] : a.dg[0]);
        float carry[NSUB][C::ITEMS][4];
        float rs[C::DEC ? C::ITEMS : 1][3][4];
#pragma unroll
        for (int sub = 0; sub < NSUB; ++sub)
#pragma unroll
            for (int it = 0; it < C::ITEMS; ++it)
#pragma unroll
                for (int e = 0; e < 4; ++e) carry[sub][it][e] = 0.f;
        if (C::DEC) {
#pragma unroll
            for (int it = 0; it < C::ITEMS; ++it)
#pragma unroll
                for (int g = 0; g < 3; ++g)
#pragma unroll
                    for (int e = 0; e < 4; ++e) rs[C::DEC ? it : 0][g][e] = 0.f;
        }
        // prefetch registers for the step about to be processed: gates (r,z,n,hn), h_prev, dh_out
        float4 pg[NSUB][C::ITEMS][4], ph[NSUB][C::ITEMS], pd[NSUB][C::ITEMS];
        auto prefetch = [&](int sub, int s) {
#pragma unroll
            for (int it = 0; it < C::ITEMS; ++it) {
                const int row = min(row0 + sub * NBS + bq + 16 * it, B - 1);
                const size_t bs = (size_t)row * L + s;
                const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int pl = 0; pl < 4; ++pl) pg[sub][it][pl] = ldg4(gates_g + (bs * 4 + pl) * HP + j0);
                if (s > 0) ph[sub][it] = ldg4(hs_g + (bs - 1) * HP + j0);
                else ph[sub][it] = (C::DEC && a.h0 != nullptr) ? ldg4(a.h0 + (size_t)row * HP + j0) : zero;
                if (C::DEC) pd[sub][it] = ldg4(a.dh_out + bs * HP + j0);
                else pd[sub][it] = (s == L - 1) ? ldg4(a.dh_fin + (size_t)row * (2 * HP) + dir * HP + j0) : zero;
            }
        };
#pragma unroll
        for (int sub = 0; sub < NSUB; ++sub) prefetch(sub, L - 1);

        for (int i = 0; i <= L; ++i) {
            const int s = L - 1 - i;                             // i == L: only collects the last contraction (dh0)
#pragma unroll
            for (int sub = 0; sub < NSUB; ++sub) {
                float* P = Pb + sub * C::P_FLOATS;
                if (i > 0) {
                    tc::mbar_wait(&bar_d[sub], (i - 1) & 1);
                    tc::tc_fence_after();
                    if (warp < 4) {                              // lane = hidden unit j (rows >= HP unused)
                        float v[32];
                        tc::tmem_ld_32x32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(sub * NBS), v);
                        const int j = warp * 32 + lane;
                        if (j < HP) {
#pragma unroll
                            for (int c = 0; c < 32; ++c) P[c * HP + j] = v[c];
                        }
                    }
                    tc::tc_fence_before();
                    epi_bar_sync<C::NT_EPI>();
                }
#pragma unroll
                for (int it = 0; it < C::ITEMS; ++it) {
                    const int b = bq + 16 * it;
                    const int row = row0 + sub * NBS + b;
                    float dh[4] = {carry[sub][it][0], carry[sub][it][1], carry[sub][it][2], carry[sub][it][3]};
                    if (i > 0) {
                        const float4 p4 = ld4(P + b * HP + j0);
                        dh[0] += p4.x; dh[1] += p4.y; dh[2] += p4.z; dh[3] += p4.w;
                    }
                    if (i == L) {
                        if (C::DEC && row < B) {
                            if (a.dh0 != nullptr) st4(a.dh0 + (size_t)row * HP + j0, make_float4(dh[0], dh[1], dh[2], dh[3]));
                            if (a.drow != nullptr) {
#pragma unroll
                                for (int g = 0; g < 3; ++g)
                                    st4(a.drow + (size_t)row * K3 + g * HP + j0,
                                        make_float4(rs[C::DEC ? it : 0][g][0], rs[C::DEC ? it : 0][g][1],
                                                    rs[C::DEC ? it : 0][g][2], rs[C::DEC ? it : 0][g][3]));
                            }
                        }
                        continue;
                    }
                    const float r4[4] = {pg[sub][it][0].x, pg[sub][it][0].y, pg[sub][it][0].z, pg[sub][it][0].w};
                    const float z4[4] = {pg[sub][it][1].x, pg[sub][it][1].y, pg[sub][it][1].z, pg[sub][it][1].w};
                    const float n4[4] = {pg[sub][it][2].x, pg[sub][it][2].y, pg[sub][it][2].z, pg[sub][it][2].w};
                    const float hn4[4] = {pg[sub][it][3].x, pg[sub][it][3].y, pg[sub][it][3].z, pg[sub][it][3].w};
                    const float hp4[4] = {ph[sub][it].x, ph[sub][it].y, ph[sub][it].z, ph[sub][it].w};
                    const float do4[4] = {pd[sub][it].x, pd[sub][it].y, pd[sub][it].z, pd[sub][it].w};
                    float o_r[4], o_z[4], o_n[4], o_hn[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float dht = dh[e] + do4[e];
                        const float dn = dht * (1.0f - z4[e]);
                        const float dz = dht * (hp4[e] - n4[e]);
                        carry[sub][it][e] = dht * z4[e];
                        const float dn_pre = dn * (1.0f - n4[e] * n4[e]);
                        const float dr = dn_pre * hn4[e];
                        o_hn[e] = dn_pre * r4[e];
                        o_r[e] = dr * r4[e] * (1.0f - r4[e]);
                        o_z[e] = dz * z4[e] * (1.0f - z4[e]);
                        o_n[e] = dn_pre;
                        if (C::DEC) {
                            rs[C::DEC ? it : 0][0][e] += o_r[e];
                            rs[C::DEC ? it : 0][1][e] += o_z[e];
                            rs[C::DEC ? it : 0][2][e] += o_n[e];
                        }
                    }
                    // operand tile: K index = g*HP + j for (dr_pre, dz_pre, dhn)
                    const int boff = (b >> 3) * X_SBO + (b & 7) * 16;
                    uint2 hi, lo;
                    unsigned char* X0 = Xb + (sub * 2 + 0) * C::X_SPLIT;
                    unsigned char* X1 = Xb + (sub * 2 + 1) * C::X_SPLIT;
                    split4(o_r, hi, lo);
                    int k = j0;
                    *reinterpret_cast<uint2*>(X0 + (k >> 3) * X_LBO + boff + (k & 7) * 2) = hi;
                    *reinterpret_cast<uint2*>(X1 + (k >> 3) * X_LBO + boff + (k & 7) * 2) = lo;
                    split4(o_z, hi, lo);
                    k = HP + j0;
                    *reinterpret_cast<uint2*>(X0 + (k >> 3) * X_LBO + boff + (k & 7) * 2) = hi;
                    *reinterpret_cast<uint2*>(X1 + (k >> 3) * X_LBO + boff + (k & 7) * 2) = lo;
                    split4(o_hn, hi, lo);
                    k = 2 * HP + j0;
                    *reinterpret_cast<uint2*>(X0 + (k >> 3) * X_LBO + boff + (k & 7) * 2) = hi;
                    *reinterpret_cast<uint2*>(X1 + (k >> 3) * X_LBO + boff + (k & 7) * 2) = lo;
                    if (row < B) {
                        float* gp = dg_g + ((size_t)row * L + s) * 4 * HP + j0;
                        st4(gp, make_float4(o_r[0], o_r[1], o_r[2], o_r[3]));
                        st4(gp + HP, make_float4(o_z[0], o_z[1], o_z[2], o_z[3]));
                        st4(gp + 2 * HP, make_float4(o_n[0], o_n[1], o_n[2], o_n[3]));
                        st4(gp + 3 * HP, make_float4(o_hn[0], o_hn[1], o_hn[2], o_hn[3]));
                    }
                }
                if (i < L) {
                    tc::fence_proxy_async();
                    tc::mbar_arrive(&bar_x[sub]);
                    if (s > 0) prefetch(sub, s - 1);             // in flight under the other sub-tile / the MMAs
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == C::NW_EPI) tc::tmem_dealloc<C::TMEM_COLS>(tmem_d);
}

using EncFwd = FwdCfg<ENC_H, ENC_H, 2, false>;
using DecFwd = FwdCfg<DEC_HP, 112, 1, true>;
using EncBwd = BwdCfg<ENC_H, 2, false>;
using DecBwd = BwdCfg<DEC_HP, 1, true>;

template <class K>
int set_smem(K kfn, size_t bytes, size_t& set_for) {
    if (set_for < bytes) {
        if (cudaFuncSetAttribute((const void*)kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
            cudaGetLastError();
            return CPG_ECUDA;
        }
        set_for = bytes;
    }
    return CPG_OK;
}
}  // namespace

int launch_gru_fwd_enc_tc(cudaStream_t s, const GruSeq* two, int B, int L, int V) {
    FwdArgs a;
    memset(&a, 0, sizeof(a));
    a.tok = two[0].tok;
    for (int d = 0; d < 2; ++d) {
        a.table[d] = two[d].table; a.whh[d] = two[d].whh; a.bhn[d] = two[d].bhn;
        a.hs[d] = two[d].hs; a.gates[d] = two[d].gates;
    }
    a.hfin = two[0].hfin;
    a.B = B; a.L = L; a.V = V;
    const size_t smem = EncFwd::smem_bytes(V, L);
    static size_t set_for = 0;
    if (set_smem(k_gru_fwd_tc<EncFwd>, smem, set_for)) return CPG_ECUDA;
    CPG_LAUNCH_NAMED("k_gru_fwd_enc_tc", k_gru_fwd_tc<EncFwd>, dim3(ceil_div(B, EncFwd::NSUB * NBS), 2), EncFwd::NTHREADS, smem, s, a);
    return CPG_OK;
}

int launch_gru_fwd_dec_tc(cudaStream_t s, const GruSeq& q, int B, int L, int V) {
    FwdArgs a;
    memset(&a, 0, sizeof(a));
    a.tok = q.tok; a.table[0] = q.table; a.whh[0] = q.whh; a.bhn[0] = q.bhn;
    a.rowbias = q.rowbias; a.h0 = q.h0; a.hs[0] = q.hs; a.gates[0] = q.gates;
    a.B = B; a.L = L; a.V = V;
    const size_t smem = DecFwd::smem_bytes(V, L);
    static size_t set_for = 0;
    if (set_smem(k_gru_fwd_tc<DecFwd>, smem, set_for)) return CPG_ECUDA;
    CPG_LAUNCH_NAMED("k_gru_fwd_dec_tc", k_gru_fwd_tc<DecFwd>, dim3(ceil_div(B, DecFwd::NSUB * NBS), 1), DecFwd::NTHREADS, smem, s, a);
    return CPG_OK;
}

int launch_gru_bwd_enc_tc(cudaStream_t s, const GruSeq* two, int B, int L) {
    BwdArgs a;
    memset(&a, 0, sizeof(a));
    for (int d = 0; d < 2; ++d) {
        a.whh[d] = two[d].whh; a.hs[d] = two[d].hs; a.gates[d] = two[d].gates; a.dg[d] = two[d].dg;
    }
    a.dh_fin = two[0].dh_fin;
    a.B = B; a.L = L;
    const size_t smem = EncBwd::smem_bytes();
    static size_t set_for = 0;
    if (set_smem(k_gru_bwd_tc<EncBwd>, smem, set_for)) return CPG_ECUDA;
    CPG_LAUNCH_NAMED("k_gru_bwd_enc_tc", k_gru_bwd_tc<EncBwd>, dim3(ceil_div(B, EncBwd::NSUB * NBS), 2), EncBwd::NTHREADS, smem, s, a);
    return CPG_OK;
}

int launch_gru_bwd_dec_tc(cudaStream_t s, const GruSeq& q, int B, int L) {
    BwdArgs a;
    memset(&a, 0, sizeof(a));
    a.whh[0] = q.whh; a.hs[0] = q.hs; a.gates[0] = q.gates; a.dg[0] = q.dg;
    a.h0 = q.h0; a.dh_out = q.dh_out; a.dh0 = q.dh0; a.drow = q.drow;
    a.B = B; a.L = L;
    const size_t smem = DecBwd::smem_bytes();
    static size_t set_for = 0;
    if (set_smem(k_gru_bwd_tc<DecBwd>, smem, set_for)) return CPG_ECUDA;
    CPG_LAUNCH_NAMED("k_gru_bwd_dec_tc", k_gru_bwd_tc<DecBwd>, dim3(ceil_div(B, DecBwd::NSUB * NBS), 1), DecBwd::NTHREADS, smem, s, a);
    return CPG_OK;
}

}  // namespace cpg
#endif  // CPG_EMU
