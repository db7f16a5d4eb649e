// Encoder bi-GRU forward recurrence on the 5th-gen tensor cores with fp32-grade accuracy.
//
// Per step the hidden-state contraction  G[128 x 240] = h[128 x 80] . W_hh^T  runs as tcgen05.mma
// (kind::f16, bf16 operands, fp32 accumulation in TMEM).  To keep the fp32 parity the path is held to
// (1e-4 on mu / logvar / logits), both operands are split into three bf16 terms,
//     x = x1 + x2 + x3   (8 + 8 + 8 mantissa bits),
// and the six products whose weight is >= 2^-16 are accumulated:  h1W1 + h1W2 + h2W1 + h2W2 + h3W1 + h1W3
// (the dropped terms are <= 2^-24 relative, i.e. fp32 rounding level).  30 MMAs of M=128, N=240, K=16
// per step replace 128*80*240 fp32 FMAs.
//
// One CTA owns 128 batch rows of one direction for all L steps:
//   warp 8      : TMEM owner + MMA issuer
//   warps 0..7  : gate epilogue; warp w serves TMEM lanes 32 (w % 4) .. +31 (thread = batch row) and the
//                 hidden units 40 (w / 4) .. +39.  Per step: tcgen05.ld the three gate pre-activations,
//                 add the token-table row, sigmoid / tanh, h' = (1-z) n + z h, stash (h, r, z, n, hn)
//                 for BPTT, split h' into 3 bf16 terms and write them as the next step's A operand.
// W_hh splits (115 KB), the h splits (60 KB) and the token table (23 KB) stay in shared memory.
// Operand layout: K-major, no swizzle ("interleaved" 8x16-byte core matrices), so a thread's 16-byte
// store of 8 consecutive units of its row is exactly one core-matrix row.
#include "ctx.h"
#ifndef CPG_EMU
#include <cuda_bf16.h>
#include "tc_common.cuh"

namespace cpg {
int check_launch(const char* where);

namespace {
constexpr int TH = ENC_H;                 // 80
constexpr int TG = 3 * TH;                // 240 gate columns
constexpr int TM = 128;                   // rows per CTA
constexpr int KC = TH / 8;                // 10 K core matrices (8 bf16 each)
constexpr int KSTEPS = TH / 16;           // 5 MMAs of K = 16 per product
constexpr int W_SPLIT_BYTES = TG * TH * 2;            // 38,400
constexpr int A_SPLIT_BYTES = TM * TH * 2;            // 20,480
constexpr int W_LBO = (TG / 8) * 128, W_SBO = 128;    // K-adjacent / N-adjacent core-matrix strides
constexpr int A_LBO = (TM / 8) * 128, A_SBO = 128;
constexpr int EPI_WARPS = 8;
constexpr int NTHREADS = (EPI_WARPS + 1) * 32;

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// kind::f16 with bf16 operands, fp32 accumulate, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// x -> three bf16 terms (round-to-nearest each), packed pairwise by the caller
__device__ __forceinline__ void split3(float x, __nv_bfloat16& a, __nv_bfloat16& b, __nv_bfloat16& c) {
    a = __float2bfloat16_rn(x);
    float r1 = x - __bfloat162float(a);
    b = __float2bfloat16_rn(r1);
    float r2 = r1 - __bfloat162float(b);
    c = __float2bfloat16_rn(r2);
}
__device__ __forceinline__ uint32_t pack2(__nv_bfloat16 lo, __nv_bfloat16 hi) {
    return (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
}
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }

struct EncTcArgs {
    const uint8_t* tok;        // [B][L]
    const float* table[2];     // [V][240] per direction
    const float* whh[2];       // [240][80] natural (N x K, K contiguous)
    const float* bhn[2];       // [80]
    float* hs[2];              // [B][L][80] by step (nullable)
    float* gates[2];           // [B][L][4][80] (nullable)
    float* hfin;               // [B][160]
    int B, L, V;
};
}  // namespace

__global__ void __launch_bounds__(NTHREADS, 1)
k_gru_fwd_enc_tc(EncTcArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    unsigned char* Wb = smem;                                   // [3][KC][TG/8][128 B]
    unsigned char* Ab = Wb + 3 * W_SPLIT_BYTES;                 // [3][KC][TM/8][128 B]
    float* tab = reinterpret_cast<float*>(Ab + 3 * A_SPLIT_BYTES);   // [V][240]
    __shared__ __align__(8) uint64_t bar_a, bar_d;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dir = blockIdx.y;
    const int row0 = blockIdx.x * TM;
    const int B = a.B, L = a.L, V = a.V;

    // ---- one-time setup: W_hh -> 3 bf16 splits in core-matrix layout, table -> smem, A tiles = 0 (h0 = 0)
    for (int idx = tid; idx < TG * KC; idx += NTHREADS) {
        const int n = idx / KC, kc = idx % KC;
        const float* src = a.whh[dir] + (size_t)n * TH + kc * 8;
        const float4 f0 = ld4(src), f1 = ld4(src + 4);
        const float x[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
        __nv_bfloat16 s1[8], s2[8], s3[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) split3(x[e], s1[e], s2[e], s3[e]);
        const int off = kc * W_LBO + (n >> 3) * W_SBO + (n & 7) * 16;
        *reinterpret_cast<uint4*>(Wb + 0 * W_SPLIT_BYTES + off) = make_uint4(pack2(s1[0], s1[1]), pack2(s1[2], s1[3]), pack2(s1[4], s1[5]), pack2(s1[6], s1[7]));
        *reinterpret_cast<uint4*>(Wb + 1 * W_SPLIT_BYTES + off) = make_uint4(pack2(s2[0], s2[1]), pack2(s2[2], s2[3]), pack2(s2[4], s2[5]), pack2(s2[6], s2[7]));
        *reinterpret_cast<uint4*>(Wb + 2 * W_SPLIT_BYTES + off) = make_uint4(pack2(s3[0], s3[1]), pack2(s3[2], s3[3]), pack2(s3[4], s3[5]), pack2(s3[6], s3[7]));
    }
    for (int i = tid; i < 3 * A_SPLIT_BYTES / 16; i += NTHREADS) reinterpret_cast<uint4*>(Ab)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < V * TG; i += NTHREADS) tab[i] = a.table[dir][i];
    if (warp == EPI_WARPS) {
        if (lane == 0) {
            tc::mbar_init(&bar_a, EPI_WARPS * 32);
            tc::mbar_init(&bar_d, 1);
            tc::fence_barrier_init();
        }
        __syncwarp();
        tc::tmem_alloc<256>(&tmem_slot);
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = tmem_slot;

    if (warp == EPI_WARPS) {
        // ---------------- MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(TM, TG);
            const uint32_t a0 = tc::smem_u32(Ab), w0 = tc::smem_u32(Wb);
            // (h split, W split) products with weight >= 2^-16
            const int ha[6] = {0, 0, 1, 1, 2, 0};
            const int wb[6] = {0, 1, 0, 1, 0, 2};
            for (int s = 0; s < L; ++s) {
                if (s > 0) {
                    tc::mbar_wait(&bar_a, (s - 1) & 1);          // h of step s-1 written by all epilogue threads
                    tc::tc_fence_after();
                }
                uint32_t acc = 0;
#pragma unroll
                for (int p = 0; p < 6; ++p) {
#pragma unroll
                    for (int ks = 0; ks < KSTEPS; ++ks) {
                        const uint64_t da = tc::make_smem_desc(a0 + ha[p] * A_SPLIT_BYTES + ks * 2 * A_LBO, A_LBO, A_SBO, 0);
                        const uint64_t db = tc::make_smem_desc(w0 + wb[p] * W_SPLIT_BYTES + ks * 2 * W_LBO, W_LBO, W_SBO, 0);
                        umma_bf16(tmem_d, da, db, idesc, acc);
                        acc = 1;
                    }
                }
                tc::umma_commit(&bar_d);
            }
        }
        __syncwarp();
    } else {
        // ---------------- gate epilogue: thread = batch row (TMEM lane), 40 hidden units
        const int q = warp & 3, half = warp >> 2;
        const int rl = q * 32 + lane;                           // row within the tile = TMEM lane
        const int row = row0 + rl;
        const bool ok = row < B;
        const int rowc = ok ? row : (B - 1);
        const int u0 = half * (TH / 2);
        const uint32_t lane_addr = tmem_d + ((uint32_t)(q * 32) << 16);
        float hprev[TH / 2];
#pragma unroll
        for (int i = 0; i < TH / 2; ++i) hprev[i] = 0.f;
        const float* bhn = a.bhn[dir];
        for (int s = 0; s < L; ++s) {
            const int t = dir ? (L - 1 - s) : s;
            const int tk = a.tok[(size_t)rowc * L + t];
            const float* trow = tab + tk * TG;
            tc::mbar_wait(&bar_d, s & 1);
            tc::tc_fence_after();
            const size_t bs = (size_t)rowc * L + s;
#pragma unroll
            for (int blk = 0; blk < TH / 16; ++blk) {            // 5 blocks of 8 units
                const int j0 = u0 + blk * 8;
                float gr[8], gz[8], gn[8];
                tmem_ld_32x8(lane_addr + (uint32_t)(j0), gr);
                tmem_ld_32x8(lane_addr + (uint32_t)(TH + j0), gz);
                tmem_ld_32x8(lane_addr + (uint32_t)(2 * TH + j0), gn);
                tmem_ld_wait();
                float hn[8], rr[8], zz[8], nn[8], hh[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int j = j0 + e;
                    rr[e] = sigmoid_fast(trow[j] + gr[e]);
                    zz[e] = sigmoid_fast(trow[TH + j] + gz[e]);
                    hh[e] = gn[e] + bhn[j];
                    nn[e] = tanh_fast(trow[2 * TH + j] + rr[e] * hh[e]);
                    hn[e] = (1.0f - zz[e]) * nn[e] + zz[e] * hprev[blk * 8 + e];
                    hprev[blk * 8 + e] = hn[e];
                }
                __nv_bfloat16 s1[8], s2[8], s3[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) split3(hn[e], s1[e], s2[e], s3[e]);
                const int off = (j0 >> 3) * A_LBO + (rl >> 3) * A_SBO + (rl & 7) * 16;
                *reinterpret_cast<uint4*>(Ab + 0 * A_SPLIT_BYTES + off) = make_uint4(pack2(s1[0], s1[1]), pack2(s1[2], s1[3]), pack2(s1[4], s1[5]), pack2(s1[6], s1[7]));
                *reinterpret_cast<uint4*>(Ab + 1 * A_SPLIT_BYTES + off) = make_uint4(pack2(s2[0], s2[1]), pack2(s2[2], s2[3]), pack2(s2[4], s2[5]), pack2(s2[6], s2[7]));
                *reinterpret_cast<uint4*>(Ab + 2 * A_SPLIT_BYTES + off) = make_uint4(pack2(s3[0], s3[1]), pack2(s3[2], s3[3]), pack2(s3[4], s3[5]), pack2(s3[6], s3[7]));
                if (ok) {
                    if (a.hs[dir] != nullptr) {
                        float* o = a.hs[dir] + bs * TH + j0;
                        st4(o, make_float4(hn[0], hn[1], hn[2], hn[3]));
                        st4(o + 4, make_float4(hn[4], hn[5], hn[6], hn[7]));
                    }
                    if (a.gates[dir] != nullptr) {
                        float* g = a.gates[dir] + bs * 4 * TH + j0;
                        st4(g, make_float4(rr[0], rr[1], rr[2], rr[3])); st4(g + 4, make_float4(rr[4], rr[5], rr[6], rr[7]));
                        st4(g + TH, make_float4(zz[0], zz[1], zz[2], zz[3])); st4(g + TH + 4, make_float4(zz[4], zz[5], zz[6], zz[7]));
                        st4(g + 2 * TH, make_float4(nn[0], nn[1], nn[2], nn[3])); st4(g + 2 * TH + 4, make_float4(nn[4], nn[5], nn[6], nn[7]));
                        st4(g + 3 * TH, make_float4(hh[0], hh[1], hh[2], hh[3])); st4(g + 3 * TH + 4, make_float4(hh[4], hh[5], hh[6], hh[7]));
                    }
                    if (s == L - 1) {
                        float* f = a.hfin + (size_t)row * (2 * TH) + dir * TH + j0;
                        st4(f, make_float4(hn[0], hn[1], hn[2], hn[3]));
                        st4(f + 4, make_float4(hn[4], hn[5], hn[6], hn[7]));
                    }
                }
            }
            tc::fence_proxy_async();                             // A-tile stores -> visible to the tensor core
            tc::tc_fence_before();                               // TMEM reads ordered before the next MMA overwrites D
            tc::mbar_arrive(&bar_a);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == EPI_WARPS) tc::tmem_dealloc<256>(tmem_d);
}

size_t gru_tc_enc_smem(int V) { return (size_t)3 * W_SPLIT_BYTES + 3 * A_SPLIT_BYTES + (size_t)V * TG * 4 + 256; }

int launch_gru_fwd_enc_tc(cudaStream_t s, const GruSeq* two, int B, int L, int V) {
    EncTcArgs a;
    a.tok = two[0].tok;
    for (int d = 0; d < 2; ++d) {
        a.table[d] = two[d].table; a.whh[d] = two[d].whh; a.bhn[d] = two[d].bhn;
        a.hs[d] = two[d].hs; a.gates[d] = two[d].gates;
    }
    a.hfin = two[0].hfin;
    a.B = B; a.L = L; a.V = V;
    const size_t smem = gru_tc_enc_smem(V);
    static size_t set_for = 0;
    if (set_for < smem) {
        cudaFuncSetAttribute((const void*)k_gru_fwd_enc_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        set_for = smem;
    }
    CPG_LAUNCH_NAMED("k_gru_fwd_enc_tc", k_gru_fwd_enc_tc, dim3(ceil_div(B, TM), 2), NTHREADS, smem, s, a);
    return CPG_OK;
}

}  // namespace cpg
#endif  // CPG_EMU
