// Decoder output layer on the 5th-gen tensor cores: out-dropout -> Linear(102 -> V) -> softmax cross-entropy,
// forward and backward fused over 128-row tiles of the [B*L, 104] hidden-state matrix (same contract as
// dec_out.cu, which stays the path for small batches; models/decoder.py:43-45,83 and losses.py:27-30).
//
// Per tile three contractions run as tcgen05.mma (kind::f16, bf16 operand terms, fp32 accumulation in TMEM).
// The logits are compared element-wise at 1e-4 relative, also where they pass through zero, so that product
// uses three bf16 terms per operand (x = x1 + x2 + x3, the six products of weight >= 2^-16: fp32-level);
// dh and dW use two terms (three products, 2^-16):
//     logits[128 x 32]   = hd[128 x 104]  . W^T              (M = rows,  N = classes, K = hidden)
//     dh[128 x 104]      = dl[128 x 32]   . W                (M = rows,  N = hidden,  K = classes)
//     dW^T[104 x 32]    += hd^T[104 x 128] . dl[128 x 32]    (M = hidden, N = classes, K = rows; accumulates in
//                                                             TMEM over all tiles of the persistent CTA)
// All operand tiles are K-major no-swizzle core-matrix tiles in shared memory; the dW contraction, which needs hd and dl
// with the tile rows as K, reads the SAME hd / dl tiles through MN-major descriptors (no transposed copies).
// Softmax / CE run with thread = row straight out of TMEM
// (all 32 class logits of a row in one thread's registers: no shuffles).
// One persistent CTA per SM, 256 threads: all 8 warps stage, warps 0-3 own the TMEM epilogues, an
// elected lane of warp 4 issues the MMAs; dh and dW complete on separate mbarriers (the dh read-out runs under the dW contraction).  Per-CTA partials (dW, db, nll) feed the ordered reduction of dec_out.cu.
#include "ctx.h"
#ifndef CPG_EMU
#include <cuda_bf16.h>
#include "tc_common.cuh"

#ifdef CPG_GRU_TIMELINE
// developer-only probe: cycles thread 0 of CTA 0 spends in each phase (includes the waits at the phase's barrier)
__device__ long long g_do_tl[16];
__device__ unsigned long long g_do_cta[3 * 160];               // per CTA: globaltimer at start, after set-up, at exit (ns)
extern "C" int cpg_debug_dec_out_timeline(long long* out) { return (int)cudaMemcpyFromSymbol(out, g_do_tl, sizeof(g_do_tl)); }
extern "C" int cpg_debug_dec_out_ctas(unsigned long long* out) { return (int)cudaMemcpyFromSymbol(out, g_do_cta, sizeof(g_do_cta)); }
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define DO_CTA(slot) do { if (threadIdx.x == 0 && blockIdx.x < 160) g_do_cta[blockIdx.x * 3 + (slot)] = gtime(); } while (0)
#define DO_MARK(slot) do { if (blockIdx.x == 0 && threadIdx.x == 0) { const long long _n = clock64(); g_do_acc[slot] += _n - g_do_last; g_do_last = _n; } } while (0)
#else
#define DO_MARK(slot) do { } while (0)
#define DO_CTA(slot) do { } while (0)
#endif

namespace cpg {

namespace {
constexpr int TR = 128;                    // rows per tile
constexpr int KH = 112;                    // hidden padded to a multiple of 16
constexpr int NF4 = DEC_HP / 4;            // 26 float4 per hidden row
constexpr int NTH = 256;
// K-major no-swizzle tiles: offset(row, k) = (k >> 3) * LBO + (row >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2
constexpr int LBO_R = (TR / 8) * 128 + 16;         // 128-row tiles (padded: conflict-free stores)
constexpr int LBO_V = (VMAX / 8) * 128;            // 32-row weight tile (written once)
constexpr int LBO_VT = (VMAX / 8) * 128 + 16;      // 32-row dl^T tile
constexpr int LBO_H = (KH / 8) * 128;              // 112-row W^T tile (written once)
constexpr int HD_SPLIT = (KH / 8) * LBO_R;         // hd   : 128 rows x K = 112
constexpr int WK_SPLIT = (KH / 8) * LBO_V;         // W    : 32 classes x K = 112
constexpr int DL_SPLIT = (VMAX / 8) * LBO_R;       // dl   : 128 rows x K = 32
constexpr int WT_SPLIT = (VMAX / 8) * LBO_H;       // W^T  : 112 hidden x K = 32
constexpr int STAGE_BYTES = TR * (DEC_HP + 4) * 4;  // fp32 read-out tile of dh (padded rows)
constexpr int OFF_HD = 0;
constexpr int OFF_WK = OFF_HD + 3 * HD_SPLIT;
constexpr int OFF_DL = OFF_WK + 3 * WK_SPLIT;
constexpr int OFF_WT = OFF_DL + 2 * DL_SPLIT;
constexpr int OFF_STAGE = OFF_WT + 2 * WT_SPLIT;
constexpr int OFF_KEEP = OFF_STAGE + STAGE_BYTES;  // [128][26] keep nibbles (one byte per 4 hidden units)
constexpr int SMEM_TOTAL = OFF_KEEP + TR * NF4 + 128;
// tensor-memory columns
constexpr int TC_LG = 0, TC_DH = 32, TC_DW = 160;
constexpr uint32_t TMEM_COLS = 256;

__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn = 0, int b_mn = 0) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
// 16-byte load of data this kernel reads exactly once: no L1 allocation (the 198 KB of shared memory leave L1 ~30 KB)
__device__ __forceinline__ float4 ld_once4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
// 16-byte store that leaves no line in L1 (the next reader is another kernel, through L2)
__device__ __forceinline__ void st_once4(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ unsigned short ld_once_u16(const unsigned short* p) {
    unsigned short v;
    asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ex2_fast(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const float f0 = __uint_as_float(hi << 16), f1 = __uint_as_float(hi & 0xffff0000u);
    const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - f0, x1 - f1);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// (x0, x1) -> three packed bf16 terms
__device__ __forceinline__ void split2x3(float x0, float x1, uint32_t& t1, uint32_t& t2, uint32_t& t3) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(x0, x1);
    t1 = *reinterpret_cast<const uint32_t*>(&a);
    const float r0 = x0 - __uint_as_float(t1 << 16), r1 = x1 - __uint_as_float(t1 & 0xffff0000u);
    const __nv_bfloat162 b = __floats2bfloat162_rn(r0, r1);
    t2 = *reinterpret_cast<const uint32_t*>(&b);
    const __nv_bfloat162 c = __floats2bfloat162_rn(r0 - __uint_as_float(t2 << 16), r1 - __uint_as_float(t2 & 0xffff0000u));
    t3 = *reinterpret_cast<const uint32_t*>(&c);
}
// contraction over `ksteps` K slices of 16 as a sum of `np` (A term, B term) products; the terms of a tile
// lie `a_split` / `b_split` bytes apart
__device__ __forceinline__ void mma_terms(uint32_t tmem_d, uint32_t a0, int a_split, int a_lbo, uint32_t b0, int b_split, int b_lbo,
                                          int ksteps, uint32_t idesc, uint32_t acc, const int* xa, const int* xb, int np) {
    for (int p = 0; p < np; ++p)
        for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t da = tc::make_smem_desc(a0 + xa[p] * a_split + ks * 2 * a_lbo, a_lbo, 128, 0);
            const uint64_t db = tc::make_smem_desc(b0 + xb[p] * b_split + ks * 2 * b_lbo, b_lbo, 128, 0);
            umma_bf16_ss(tmem_d, da, db, idesc, acc);
            acc = 1;
        }
}
__device__ __forceinline__ void tmem_ld_16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// three-product split-bf16 contraction over `ksteps` K slices of 16: A and B tiles with their two terms
// `a_split` / `b_split` bytes apart
__device__ __forceinline__ void mma_split3(uint32_t tmem_d, uint32_t a0, int a_split, int a_lbo, uint32_t b0, int b_split, int b_lbo,
                                           int ksteps, uint32_t idesc, uint32_t acc, int np = 3) {
    const int xs[3] = {0, 0, 1}, ws[3] = {0, 1, 0};
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        if (p >= np) break;
        for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t da = tc::make_smem_desc(a0 + xs[p] * a_split + ks * 2 * a_lbo, a_lbo, 128, 0);
            const uint64_t db = tc::make_smem_desc(b0 + ws[p] * b_split + ks * 2 * b_lbo, b_lbo, 128, 0);
            umma_bf16_ss(tmem_d, da, db, idesc, acc);
            acc = 1;
        }
    }
}
// the same, both operands read MN-major from K-major tiles whose rows are this contraction's K: element (mn, k) of a tile
// with chunk stride `lbo` sits at (mn >> 3) * lbo + (k >> 3) * 128 + (k & 7) * 16 + (mn & 7) * 2
__device__ __forceinline__ void mma_split3_mn(uint32_t tmem_d, uint32_t a0, int a_split, int a_lbo, uint32_t b0, int b_split, int b_lbo,
                                              int ksteps, uint32_t idesc, uint32_t acc, int np = 3) {
    const int xs[3] = {0, 0, 1}, ws[3] = {0, 1, 0};
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        if (p >= np) break;
        for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t da = tc::make_smem_desc(a0 + xs[p] * a_split + ks * 256, 128, a_lbo, 0);
            const uint64_t db = tc::make_smem_desc(b0 + ws[p] * b_split + ks * 256, 128, b_lbo, 0);
            umma_bf16_ss(tmem_d, da, db, idesc, acc);
            acc = 1;
        }
    }
}
}  // namespace

__global__ void __launch_bounds__(NTH, 1)
k_dec_out_tc(DecOutArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* HD = smem + OFF_HD;
    unsigned char* WK = smem + OFF_WK;
    unsigned char* DL = smem + OFF_DL;
    unsigned char* WT = smem + OFF_WT;
    unsigned char* STG = smem + OFF_STAGE;
    unsigned char* keep_s = smem + OFF_KEEP;
    __shared__ __align__(8) uint64_t bar_m, bar_w;
    __shared__ uint32_t tmem_slot;
    __shared__ float red_b[8][VMAX / 2];
    __shared__ float red_n[4];
    __shared__ float ex_s[3][2][TR];                           // row max / sum / target logit of each class half

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    DO_CTA(0);
    const int V = a.V;
    const int nrows = a.B * a.L;
    const bool want_grad = a.dh_out != nullptr;
    const bool has_mask = a.out_keep != nullptr;
    const float sc = has_mask ? a.keep_scale : 1.0f;
    const float ntok_v = a.ntok_i != nullptr ? (float)*a.ntok_i : (a.ntok != nullptr ? *a.ntok : 0.f);
    const float inv_ntok = ntok_v > 0.f ? 1.0f / ntok_v : 0.f;

    // ---- one-time setup: zero what no tile ever rewrites but an MMA reads -- the K padding of hd (hidden 104..111, the last
    // chunk of each term) and the two weight tiles (their padding rows / columns) -- then stage the weight tiles.  (Every
    // other byte of hd / dl / the read-out tile is written by each tile before it is read; measured: set-up 6.6 us per CTA
    // with all 198 KB zeroed and the 13 weight loads of a thread issued one by one.)
    static_assert(DEC_HP == 104 && KH == 112, "K padding of hd = its last 8-column chunk");
    for (int i = tid; i < 3 * (LBO_R / 16); i += NTH) {
        const int term = i / (LBO_R / 16), q16 = i % (LBO_R / 16);
        reinterpret_cast<uint4*>(HD + term * HD_SPLIT + (KH / 8 - 1) * LBO_R)[q16] = make_uint4(0, 0, 0, 0);
    }
    for (int i = tid; i < (OFF_STAGE - OFF_WK) / 16; i += NTH) {
        if (i < 3 * WK_SPLIT / 16 || i >= (OFF_WT - OFF_WK) / 16) reinterpret_cast<uint4*>(WK)[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    static_assert(VMAX * DEC_HP % NTH == 0, "weight staging items");
    float wreg[VMAX * DEC_HP / NTH];
#pragma unroll
    for (int it = 0; it < VMAX * DEC_HP / NTH; ++it) wreg[it] = __ldg(a.fc_w + tid + it * NTH);
#pragma unroll
    for (int it = 0; it < VMAX * DEC_HP / NTH; ++it) {
        const int idx = tid + it * NTH;
        const int v = idx / DEC_HP, j = idx % DEC_HP;
        const float w = wreg[it];
        const __nv_bfloat16 h = __float2bfloat16_rn(w);
        const float wr = w - __bfloat162float(h);
        const __nv_bfloat16 l = __float2bfloat16_rn(wr);
        const __nv_bfloat16 l3 = __float2bfloat16_rn(wr - __bfloat162float(l));
        // W as [class][k = hidden] and W^T as [hidden][k = class]
        const int o1 = (j >> 3) * LBO_V + (v >> 3) * 128 + (v & 7) * 16 + (j & 7) * 2;
        const int o2 = (v >> 3) * LBO_H + (j >> 3) * 128 + (j & 7) * 16 + (v & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(WK + o1) = h;
        *reinterpret_cast<__nv_bfloat16*>(WK + WK_SPLIT + o1) = l;
        *reinterpret_cast<__nv_bfloat16*>(WK + 2 * WK_SPLIT + o1) = l3;
        *reinterpret_cast<__nv_bfloat16*>(WT + o2) = h;
        *reinterpret_cast<__nv_bfloat16*>(WT + WT_SPLIT + o2) = l;
    }
    if (warp == 4) {
        if (lane == 0) {
            tc::mbar_init(&bar_m, 1);
            tc::mbar_init(&bar_w, 1);
            tc::fence_barrier_init();
        }
        __syncwarp();
        tc::tmem_alloc<TMEM_COLS>(&tmem_slot);
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t s_hd = tc::smem_u32(HD), s_wk = tc::smem_u32(WK), s_dl = tc::smem_u32(DL), s_wt = tc::smem_u32(WT);

    // epilogue threads (warps 0-3): thread = row of the tile = TMEM lane
    const int rl = tid;                                        // valid for tid < 128
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float accb[VMAX / 2];                                      // db partial of this thread's half of the classes
#pragma unroll
    for (int v = 0; v < VMAX / 2; ++v) accb[v] = 0.f;
    float accn = 0.f;
    uint32_t mphase = 0, wphase = 0;
    bool dw_started = false;

#ifdef CPG_GRU_TIMELINE
    long long g_do_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long g_do_last = clock64();
#endif
    const int ntiles = ceil_div(nrows, TR);
    DO_CTA(1);
    // Raw inputs of a tile ((row, quad) items: 13 per thread): fetched one tile ahead -- the loads of tile t+1 are issued
    // right after the second MMA batch of tile t and land under its dh epilogue.  Unconditional (clamped) addresses
    // and no use of the loaded values at the issue point, so that the 39 loads of a thread are in flight together.
    constexpr int NIT = TR * NF4 / NTH;
    static_assert(TR * NF4 % NTH == 0, "staging items");
    float4 hq[NIT];
    unsigned short k0[NIT], k1[NIT];                             // raw keep bytes (units 4f, 4f+1 | 4f+2, 4f+3)
    auto load_tile = [&](int tile) {
        const int row0 = tile * TR;
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int idx = tid + it * NTH;
            const int r = idx / NF4, f = idx % NF4;
            const int row = min(row0 + r, nrows - 1);
            hq[it] = ld_once4(a.hs + (size_t)row * DEC_HP + f * 4);
            k0[it] = 0x0101; k1[it] = 0x0101;
            if (has_mask) {
                // row * 102 + 4 f is even: two aligned 2-byte loads (the last quad's second pair is clamped, masked below)
                const unsigned short* kp = reinterpret_cast<const unsigned short*>(a.out_keep + (size_t)row * DEC_H + f * 4);
                k0[it] = ld_once_u16(kp);
                k1[it] = ld_once_u16(kp + (f * 4 + 2 < DEC_H ? 1 : 0));
            }
        }
    };
    if ((int)blockIdx.x < ntiles) load_tile(blockIdx.x);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int row0 = tile * TR;
        DO_MARK(7);
        // ---- S1) hd = h * keep * scale -> HD (K-major, three bf16 terms) + keep nibbles; coalesced over (row, quad)
        {
            if (!want_grad && tile != (int)blockIdx.x) load_tile(tile);     // forward-only: no second MMA batch to hide under
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int idx = tid + it * NTH;
                const int r = idx / NF4, f = idx % NF4;
                const float hh[4] = {hq[it].x, hq[it].y, hq[it].z, hq[it].w};
                const bool in_rows = row0 + r < nrows;
                const uint32_t kq = (uint32_t)k0[it] | ((f * 4 + 2 < DEC_H) ? (uint32_t)k1[it] << 16 : 0u);
                float hv[4];
                unsigned kb = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const bool k = in_rows && ((kq >> (8 * c)) & 0xff) != 0;
                    hv[c] = k ? hh[c] * sc : 0.f;
                    kb |= k ? (1u << c) : 0u;
                }
                uint2 t1, t2, t3;
                split2x3(hv[0], hv[1], t1.x, t2.x, t3.x);
                split2x3(hv[2], hv[3], t1.y, t2.y, t3.y);
                const int j0 = f * 4;
                const int off = (j0 >> 3) * LBO_R + (r >> 3) * 128 + (r & 7) * 16 + (j0 & 7) * 2;
                *reinterpret_cast<uint2*>(HD + off) = t1;
                *reinterpret_cast<uint2*>(HD + HD_SPLIT + off) = t2;
                *reinterpret_cast<uint2*>(HD + 2 * HD_SPLIT + off) = t3;
                keep_s[r * NF4 + f] = (unsigned char)kb;
            }
        }
        DO_MARK(0);
        tc::fence_proxy_async();
        __syncthreads();
        DO_MARK(1);
        // ---- M1) logits
        if (warp == 4) {
            tc::tc_fence_after();
            if (elect_one()) {
                const int xa[6] = {0, 0, 1, 1, 2, 0}, xb[6] = {0, 1, 0, 1, 0, 2};
                mma_terms(tmem + TC_LG, s_hd, HD_SPLIT, LBO_R, s_wk, WK_SPLIT, LBO_V, KH / 16, idesc_bf16(128, VMAX), 0, xa, xb,
                          a.nprod == 1 ? 1 : 6);
                tc::umma_commit(&bar_m);
            }
            __syncwarp();
        }
        // ---- E1) softmax / CE: thread = (row, half of the classes); warps w and w + 4 share TMEM lane quadrant w & 3
        // and exchange their row maxima / sums through shared memory
        {
            constexpr int HV = VMAX / 2;
            const int half = warp >> 2, v0 = half * HV;
            const int r2 = (warp & 3) * 32 + lane;
            const int row = row0 + r2;
            // the row's target and the bias do not depend on the MMAs: their (global) loads run under them
            const int tg = (a.fused_ce && row < nrows) ? a.tgt[row] : PAD;
            float fb[HV];
#pragma unroll
            for (int i = 0; i < HV; ++i) fb[i] = v0 + i < V ? __ldg(a.fc_b + v0 + i) : 0.f;
            tc::mbar_wait(&bar_m, mphase & 1);
            tc::tc_fence_after();
            float lg[HV];
            tmem_ld_16(lane_addr + TC_LG + v0, lg);
            float mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < HV; ++i) {
                lg[i] = v0 + i < V ? lg[i] + fb[i] : -INFINITY;
                mx = fmaxf(mx, lg[i]);
            }
            if (row < nrows && a.logits_out != nullptr) {
#pragma unroll
                for (int i = 0; i < HV; ++i)
                    if (v0 + i < V) a.logits_out[(size_t)row * V + v0 + i] = lg[i];
            }
            float dl[HV];
#pragma unroll
            for (int i = 0; i < HV; ++i) dl[i] = 0.f;
            if (a.fused_ce) {
                ex_s[0][half][r2] = mx;
                __syncthreads();
                mx = fmaxf(ex_s[0][0][r2], ex_s[0][1][r2]);
                float e[HV], se = 0.f, lt = 0.f;
#pragma unroll
                for (int i = 0; i < HV; ++i) {
                    e[i] = v0 + i < V ? ex2_fast((lg[i] - mx) * 1.4426950408889634f) : 0.f;     // SFU exp2: 2^-22 relative
                    se += e[i];
                    if (v0 + i == tg) lt = lg[i];
                }
                ex_s[1][half][r2] = se;
                ex_s[2][half][r2] = lt;
                __syncthreads();
                se = ex_s[1][0][r2] + ex_s[1][1][r2];
                lt = ex_s[2][0][r2] + ex_s[2][1][r2];
                if (tg != PAD) {
                    if (half == 0) accn += (mx + __logf(se)) - lt;
                    const float inv = inv_ntok / se;
#pragma unroll
                    for (int i = 0; i < HV; ++i) dl[i] = v0 + i < V ? e[i] * inv - (v0 + i == tg ? inv_ntok : 0.f) : 0.f;
                }
            } else if (a.dlogits_in != nullptr && row < nrows) {
#pragma unroll
                for (int i = 0; i < HV; ++i)
                    if (v0 + i < V) dl[i] = a.dlogits_in[(size_t)row * V + v0 + i];
            }
            if (want_grad) {
#pragma unroll
                for (int i = 0; i < HV; ++i) accb[i] += dl[i];
                // dl -> DL (K = class, K-major; the dW contraction reads the same tile MN-major)
#pragma unroll
                for (int kq = 0; kq < HV / 8; ++kq) {
                    const int kc = half * (HV / 8) + kq;
                    uint4 hi, lo;
                    split2(dl[kq * 8 + 0], dl[kq * 8 + 1], hi.x, lo.x);
                    split2(dl[kq * 8 + 2], dl[kq * 8 + 3], hi.y, lo.y);
                    split2(dl[kq * 8 + 4], dl[kq * 8 + 5], hi.z, lo.z);
                    split2(dl[kq * 8 + 6], dl[kq * 8 + 7], hi.w, lo.w);
                    const int off = kc * LBO_R + (r2 >> 3) * 128 + (r2 & 7) * 16;
                    *reinterpret_cast<uint4*>(DL + off) = hi;
                    *reinterpret_cast<uint4*>(DL + DL_SPLIT + off) = lo;
                }
            }
            tc::tc_fence_before();
        }
        ++mphase;
        if (!want_grad) { __syncthreads(); continue; }
        tc::fence_proxy_async();
        __syncthreads();
        DO_MARK(2);
        // ---- M2) dh = dl W  and  dW^T += hd^T dl
        if (warp == 4) {
            tc::tc_fence_after();
            if (elect_one()) {
                mma_split3(tmem + TC_DH, s_dl, DL_SPLIT, LBO_R, s_wt, WT_SPLIT, LBO_H, VMAX / 16, idesc_bf16(128, KH), 0, a.nprod == 1 ? 1 : 3);
                tc::umma_commit(&bar_m);                       // dh is complete: its read-out starts under the dW contraction
                mma_split3_mn(tmem + TC_DW, s_hd, HD_SPLIT, LBO_R, s_dl, DL_SPLIT, LBO_R, TR / 16, idesc_bf16(128, VMAX, 1, 1),
                              dw_started ? 1u : 0u, a.nprod == 1 ? 1 : 3);
                tc::umma_commit(&bar_w);
            }
            __syncwarp();
        }
        dw_started = true;
        if (tile + (int)gridDim.x < ntiles) load_tile(tile + gridDim.x);      // next tile's inputs: in flight under E2
        // ---- E2) dh_out = dh * keep * scale.  TMEM read-out with thread = row (warps w and w + 4 share a lane quadrant and
        // split the columns) into a padded fp32 tile, then a row-contiguous masked copy to HBM by all 256 threads.
        {
            constexpr int SST = DEC_HP + 4;                    // 108 floats: conflict-free float4 rows
            float* stage = reinterpret_cast<float*>(STG);
            tc::mbar_wait(&bar_m, mphase & 1);
            tc::tc_fence_after();
            const int r2 = (warp & 3) * 32 + lane;
            const int cbeg = warp < 4 ? 0 : 64, cend = warp < 4 ? 64 : DEC_HP;
#pragma unroll 1
            for (int c0 = cbeg; c0 < cend; c0 += 32) {
                float v[32];
                tc::tmem_ld_32x32(lane_addr + TC_DH + c0, v);     // columns >= 104 of the last chunk: padding, unused
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (c0 + q * 4 < DEC_HP) st4(stage + r2 * SST + c0 + q * 4, make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]));
            }
            tc::tc_fence_before();
            __syncthreads();
#pragma unroll
            for (int it = 0; it < TR * NF4 / NTH; ++it) {
                const int idx = tid + it * NTH;
                const int r = idx / NF4, f = idx % NF4, row = row0 + r;
                if (row < nrows) {
                    const float4 d = ld4(stage + r * SST + f * 4);
                    const unsigned k = keep_s[r * NF4 + f];
                    st_once4(a.dh_out + (size_t)row * DEC_HP + f * 4,
                        make_float4((k & 1) ? d.x * sc : 0.f, (k & 2) ? d.y * sc : 0.f, (k & 4) ? d.z * sc : 0.f, (k & 8) ? d.w * sc : 0.f));
                }
            }
        }
        ++mphase;
        tc::mbar_wait(&bar_w, wphase & 1);                     // the dW contraction is done with HD / DL
        ++wphase;
        __syncthreads();                                       // ... and the stage tile is read out
        DO_MARK(3);
    }

#ifdef CPG_GRU_TIMELINE
    if (blockIdx.x == 0 && threadIdx.x == 0) for (int i = 0; i < 8; ++i) g_do_tl[i] = g_do_acc[i];
#endif
    // ---- per-CTA partials
    tc::tc_fence_after();
    if (want_grad) {
        float* pw = a.part_w + (size_t)blockIdx.x * VMAX * DEC_HP;
        if (warp < 4) {
            float dw[VMAX];
            if (dw_started) tc::tmem_ld_32x32(lane_addr + TC_DW, dw);      // lane = hidden unit j, columns = classes
            const int j = tid;
            if (j < DEC_HP) {
#pragma unroll
                for (int v = 0; v < VMAX; ++v) pw[v * DEC_HP + j] = dw_started ? dw[v] : 0.f;
            }
        }
#pragma unroll
        for (int v = 0; v < VMAX / 2; ++v) {
            const float sv = warp_sum(accb[v]);
            if (lane == 0) red_b[warp][v] = sv;
        }
        __syncthreads();
        if (tid < VMAX) {
            const int hf = tid / (VMAX / 2), v = tid % (VMAX / 2);
            a.part_b[(size_t)blockIdx.x * VMAX + tid] =
                (red_b[hf * 4 + 0][v] + red_b[hf * 4 + 1][v]) + (red_b[hf * 4 + 2][v] + red_b[hf * 4 + 3][v]);
        }
    }
    if (a.part_nll != nullptr) {
        if (warp < 4) {
            const float s = warp_sum(accn);
            if (lane == 0) red_n[warp] = s;
        }
        __syncthreads();
        if (tid == 0) a.part_nll[blockIdx.x] = (red_n[0] + red_n[1]) + (red_n[2] + red_n[3]);
    }
    tc::tc_fence_before();
    __syncthreads();
    DO_CTA(2);
    if (warp == 4) tc::tmem_dealloc<TMEM_COLS>(tmem);
}

int launch_dec_out_tc(cudaStream_t s, const DecOutArgs& a, int sm_count, int* parts_out) {
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute((const void*)k_dec_out_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL) != cudaSuccess) {
            cudaGetLastError();
            return CPG_ECUDA;
        }
        attr_set = true;
    }
    const int ntiles = ceil_div(a.B * a.L, TR);
    // the smallest grid with the same number of tiles on the busiest CTA (800 tiles: 134 CTAs x 6 instead of 148): the SMs
    // left over run the kernels of the other lanes, whose CTAs would otherwise hold back CTAs of this kernel
    const int per_cta = ceil_div(ntiles, max(1, min(ntiles, sm_count)));
    const int grid = max(1, ceil_div(ntiles, per_cta));
    CPG_LAUNCH_NAMED("k_dec_out_tc", k_dec_out_tc, grid, NTH, SMEM_TOTAL, s, a);
    *parts_out = grid;
    return CPG_OK;
}

}  // namespace cpg
#endif  // CPG_EMU
