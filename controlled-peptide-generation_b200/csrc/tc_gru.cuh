// Device helpers shared by the tcgen05 GRU kernels (gru_tc.cu: recurrences with W_hh resident in tensor memory;
// gru_bwd_fused.cu: BPTT with the weight-gradient contractions fused in).
#pragma once
#ifndef CPG_EMU
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "tc_common.cuh"

namespace cpg {
namespace {
constexpr int X_SBO = 128;                            // N-adjacent core matrices of an operand tile

// D[tmem] (+)= A[tmem] * B[smem desc]: kind::f16 (bf16 operands), fp32 accumulate; the A operand is resident
// in tensor memory (lane = M row, one 32-bit column = two consecutive K elements, even K in the low half)
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
// 32 lanes x N consecutive fp32 columns (N = 16 | 32): thread = TMEM lane of this warp's quadrant
template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float (&v)[N]) {
    static_assert(N == 16 || N == 32, "column count");
    uint32_t r[N];
    if constexpr (N == 32) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr) : "memory");
    } else {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr) : "memory");
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = __uint_as_float(r[i]);
}
// kind::f16 with bf16 operands, fp32 accumulate, B operand K-major
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
// named barrier of one chain's epilogue warps
__device__ __forceinline__ void group_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// (x0, x1) -> packed bf16 leading terms (x0 in the low half) and packed bf16 remainders
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const float f0 = __uint_as_float(hi << 16), f1 = __uint_as_float(hi & 0xffff0000u);
    const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - f0, x1 - f1);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
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
// fp16 flavour of the split (11 + 11 mantissa bits; for operands of bounded magnitude)
__device__ __forceinline__ void split2h(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(x0, x1);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(x0 - f.x, x1 - f.y);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// Gate non-linearities straight on the SFU (ex2.approx + rcp.approx, flush-to-zero forms: no denormal
// fix-up code around them); absolute error <= ~6e-7 like the SIMT kernels' versions.
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_ftz(1.0f + ex2_ftz(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_fast(float x) { return fmaf(-2.0f, rcp_ftz(1.0f + ex2_ftz(2.8853900817779268f * x)), 1.0f); }
// r = sigmoid(a), z = sigmoid(b) with ONE reciprocal: 1/((1+e^-a)(1+e^-b)) feeds both (the SFU, at 16 ops/clk/SM,
// is what bounds the forward gate math: 5 instead of 6 SFU ops per hidden unit).  The exponents are clamped
// so that the product of the two denominators stays finite.
__device__ __forceinline__ void sigmoid_pair(float a, float b, float& r, float& z) {
    const float ea = ex2_ftz(fminf(-1.4426950408889634f * a, 60.0f));
    const float eb = ex2_ftz(fminf(-1.4426950408889634f * b, 60.0f));
    const float pa = 1.0f + ea, pb = 1.0f + eb;
    const float inv = rcp_ftz(pa * pb);
    r = pb * inv;
    z = pa * inv;
}
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// streaming 16-byte load of stash data read exactly once: no L1 allocation
__device__ __forceinline__ float4 ld_stream4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// streaming 16-byte store (evict-first): activation-stash data that is read again only much later -- written with the
// default policy it sits dirty in L2 and its write-back taxes the kernels that follow
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// 16-byte store that leaves no line in L1 (the next reader is another kernel, through L2)
__device__ __forceinline__ void st_once4(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// (operand-tile term, weight term) of the three accumulated products
__device__ constexpr int XS[3] = {0, 0, 1};
__device__ constexpr int WS[3] = {0, 1, 0};

// 1-D bulk copy global -> shared through the TMA engine, completion counted on an mbarrier
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tc::smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}

// 1-D bulk copy shared -> global through the TMA engine (bulk async-group completion)
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gdst), "r"(tc::smem_u32(smem_src)), "r"(bytes) : "memory");
}
// the same with the evict-first L2 policy (stash data that is read again only after much other traffic)
__device__ __forceinline__ void bulk_store_stream(void* gdst, const void* smem_src, uint32_t bytes) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                 ::"l"(gdst), "r"(tc::smem_u32(smem_src)), "r"(bytes), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Gate stash (r, z, n, hn) private to this file's forward/backward pair: 32-row tiles, step-major inside a
// tile, [tile][step][plane][row in tile][HP] -- a chain streams through one contiguous region (whole DRAM
// pages per step) instead of 320-byte pieces 8 KB apart as in the row-major [B][L][4][HP] of the SIMT kernels.
template <int HP>
__device__ __forceinline__ size_t gate_stash_offset(int row, int s, int L) {
    return (((size_t)(row >> 5) * L + s) * 4 * 32 + (row & 31)) * HP;
}

// Which warp of a chain's group serves TMEM lane quadrant (warp % 4) for task number `task` of that quadrant:
// the group's warps with equal quadrant are wl, wl+4, ...; tasks are dealt round-robin over them.
__device__ __forceinline__ bool quadrant_task_is_mine(int wl, int nwg, int task) {
    const int cnt = (nwg - (wl & 3) + 3) >> 2;
    return task % cnt == (wl >> 2);
}

}  // namespace
}  // namespace cpg
#endif
