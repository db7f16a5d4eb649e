// Building blocks of the split-precision tcgen05 dense kernels (latent_tc.cu, rf_tc.cu): operand tiles in the no-swizzle
// core-matrix layouts, conversion of fp32 rows into them, the three-product MMA issue loop, epilogue row mapping.
#pragma once
#ifndef CPG_EMU
#include <cuda_fp16.h>
#include "tc_gru.cuh"

namespace cpg {
namespace {
constexpr int LT_THREADS = 512;
constexpr int LT_PARTS = LT_THREADS / 128;      // warps per TMEM lane quadrant: they take 16-column slices in turn

__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__host__ __device__ constexpr uint32_t idesc_16(int M, int N, int a_mn, int b_mn, bool half) {
    return (1u << 4) | (half ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// fp16 split (x = x1 + x2, 11 + 11 mantissa bits): for operands of bounded magnitude (|x| < 6e4; what falls below the fp16
// subnormal range is below 6e-8 absolute) the three products are fp32-grade in ABSOLUTE terms -- needed for the forward
// projections, whose result enters all 25 decoder steps (the bf16 split's 2^-16 left the logits 3e-6 off at zero crossings);
// gradients (tiny magnitudes) keep the bf16 split, whose exponent range is fp32's.
template <bool HALF>
__device__ __forceinline__ void split4x(const float (&x)[4], uint2& hi, uint2& lo) {
    if (HALF) { split2h(x[0], x[1], hi.x, lo.x); split2h(x[2], x[3], hi.y, lo.y); }
    else split4(x, hi, lo);
}
template <bool HALF>
__device__ __forceinline__ void split8x(const float (&x)[8], uint4& hi, uint4& lo) {
    if (HALF) {
        split2h(x[0], x[1], hi.x, lo.x); split2h(x[2], x[3], hi.y, lo.y);
        split2h(x[4], x[5], hi.z, lo.z); split2h(x[6], x[7], hi.w, lo.w);
    } else {
        split8(x, hi, lo);
    }
}

// K-major tile of ROWS rows x K_PAD columns (two 16-bit terms): element (r, k) at (k/8) * ROWS * 16 + (r/8) * 128 +
// (r%8) * 16 + (k%8) * 2.  Filled from a row-major fp32 matrix `src` (leading dimension ld, k_valid columns, rows past
// n_rows read as 0): a warp-task = 8 rows x 16 columns, lane = (row in block, float4 of the 64-byte piece) -- full
// 32-byte sectors on the global side, 8-byte stores that tile 128 contiguous bytes per half-warp on the shared side.
template <bool HALF, int ROWS, int K_PAD>
__device__ __forceinline__ void fill_rows_kmajor(unsigned char* hi, unsigned char* lo, const float* __restrict__ src, int ld,
                                                 int n_rows, int k_valid) {
    constexpr int NT = (ROWS / 8) * (K_PAD / 16);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rl = lane & 7, fl = lane >> 3;
#pragma unroll 4
    for (int t = warp; t < NT; t += LT_THREADS / 32) {
        const int rb = t % (ROWS / 8), kg = t / (ROWS / 8);
        const int r = rb * 8 + rl, k0 = kg * 16 + fl * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < n_rows && k0 < k_valid) v = ld_stream4(src + (size_t)r * ld + k0);
        const float x[4] = {v.x, v.y, v.z, v.w};
        uint2 h, l;
        split4x<HALF>(x, h, l);
        const int off = (k0 >> 3) * (ROWS * 16) + rb * 128 + rl * 16 + (k0 & 7) * 2;
        *reinterpret_cast<uint2*>(hi + off) = h;
        *reinterpret_cast<uint2*>(lo + off) = l;
    }
}

// three split products over K = k_len: A = K-major tile of a_rows rows, or (a_mn) an MN-major tile (m fastest) with a_rows K
// rows, read from row m0; B = K-major rows n0.. of a b_rows-row tile, or (b_mn) an MN-major tile (n fastest) with b_rows K
// rows; `first` = overwrite the accumulator
__device__ __forceinline__ void issue_products(uint32_t tmem_d, int M, uint32_t a_hi, uint32_t a_lo, int a_rows, uint32_t b_hi,
                                               uint32_t b_lo, bool b_mn, int b_rows, int n0, int N, int k_len, bool first, bool half,
                                               bool a_mn = false, int m0 = 0) {
    const uint32_t idesc = idesc_16(M, N, a_mn ? 1 : 0, b_mn ? 1 : 0, half);
    uint32_t acc = first ? 0u : 1u;
#pragma unroll 1
    for (int p = 0; p < 3; ++p) {
        const uint32_t a0 = XS[p] ? a_lo : a_hi, b0 = WS[p] ? b_lo : b_hi;
#pragma unroll 1
        for (int ks = 0; ks < k_len / 16; ++ks) {
            uint64_t da, db;
            if (!a_mn) da = tc::make_smem_desc(a0 + ks * 2 * (a_rows * 16), a_rows * 16, 128, 0);
            else da = tc::make_smem_desc(a0 + (m0 >> 3) * (a_rows * 16) + ks * 256, 128, a_rows * 16, 0);
            if (!b_mn) db = tc::make_smem_desc(b0 + ks * 2 * (b_rows * 16) + (n0 >> 3) * 128, b_rows * 16, 128, 0);
            else db = tc::make_smem_desc(b0 + (n0 >> 3) * (b_rows * 16) + ks * 256, 128, b_rows * 16, 0);
            umma_ss(tmem_d, da, db, idesc, acc);
            acc = 1;
        }
    }
}

// row of the tile held by this thread in the epilogue and whether the thread holds one: M = 128 -> TMEM lane = row,
// M = 64 -> rows 16 q .. 16 q + 15 sit in lanes 0-15 of quadrant q
template <int M>
__device__ __forceinline__ int epi_row(int q, int lane, bool& has) {
    if (M == 128) { has = true; return q * 32 + lane; }
    has = lane < 16;
    return q * 16 + (lane & 15);
}

}  // namespace
}  // namespace cpg
#endif
