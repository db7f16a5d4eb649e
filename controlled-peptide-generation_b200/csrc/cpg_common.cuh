// Shared device/host helpers for the cpg_b200 kernels (sm_100a).
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef CPG_EMU
#include "cuda_emu.h"                       // tools/cuda_emu (developer tool, g++ build)
#define CPG_LAUNCH_NAMED(label, kernel, grid, block, smem, stream, ...) \
    do { ++cpg::g_launch_count; emu::launch(dim3(grid), dim3(block), (size_t)(smem), [=]() { kernel(__VA_ARGS__); }); } while (0)
#define CPG_LAUNCH(kernel, grid, block, smem, stream, ...) \
    CPG_LAUNCH_NAMED(#kernel, kernel, grid, block, smem, stream, __VA_ARGS__)
#define CPG_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::cur_block()->dyn_smem)
#define CPG_SET_MAX_SMEM(kernel, bytes) 0
#else
#include <cuda_runtime.h>
#define CPG_LAUNCH_NAMED(label, kernel, grid, block, smem, stream, ...)                      \
    do {                                                                                     \
        ++cpg::g_launch_count;                                                               \
        if (cpg::g_profile_on) cpg::prof_begin(label, (stream));                             \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                          \
        if (cpg::g_profile_on) cpg::prof_end((stream));                                      \
    } while (0)
#define CPG_LAUNCH(kernel, grid, block, smem, stream, ...) \
    CPG_LAUNCH_NAMED(#kernel, kernel, grid, block, smem, stream, __VA_ARGS__)
#define CPG_DYN_SMEM(type, name)                                   \
    extern __shared__ __align__(128) unsigned char name##_raw_[];  \
    type* name = reinterpret_cast<type*>(name##_raw_)
#define CPG_SET_MAX_SMEM(kernel, bytes) \
    cudaFuncSetAttribute((const void*)(kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))
#endif

namespace cpg {

extern long long g_launch_count;      // kernels enqueued by this library (reported by cpg_launch_count)
// optional per-kernel CUDA-event timing (cpg_profile_*): off by default, one branch per launch
extern bool g_profile_on;
const char* shape_label(const char* name, int M, int N, int K, int nprod);
#ifndef CPG_EMU
void prof_begin(const char* label, cudaStream_t s);
void prof_end(cudaStream_t s);
#endif

// ---- token ids (models/mutils.py:5-8)
constexpr int UNK = 0, PAD = 1, START = 2, EOS = 3;

// ---- model geometry of the reference configuration (cfg.py:258-281)
constexpr int EMB = 150;          // emb_dim
constexpr int ENC_H = 80;         // E_args.h_dim
constexpr int ZD = 100;           // z_dim
constexpr int CD = 2;             // c_dim
constexpr int DEC_H = ZD + CD;    // 102
constexpr int DEC_HP = 104;       // decoder hidden padded to a multiple of 4 (zero weights)
constexpr int DEC_IN = EMB + DEC_H;   // 252
constexpr int VMAX = 32;          // vocabulary fits one warp lane per class
constexpr int LMAX = 32;          // max sequence length supported (reference: 25)

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// ---- ordered reduction of split partials:  sum_p part[p * pstride + off]
// Block = (32 outputs) x (RED_Y split lanes).  Lane y accumulates partials y, y + RED_Y, ... with four
// independent loads in flight, then the RED_Y lane sums are added in lane order from shared memory: the
// result is bit-reproducible run to run (no atomics) and the serial chain is nsplit / RED_Y long instead of nsplit.
constexpr int RED_X = 32, RED_Y = 8;
__device__ __forceinline__ float block_split_sum(const float* __restrict__ part, size_t pstride, int nsplit, size_t off, bool active) {
    __shared__ float red_s[RED_Y][RED_X + 1];
    const int y = threadIdx.y;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (active) {
        int p = y;
        for (; p + 3 * RED_Y < nsplit; p += 4 * RED_Y) {
            s0 += part[(size_t)p * pstride + off];
            s1 += part[(size_t)(p + RED_Y) * pstride + off];
            s2 += part[(size_t)(p + 2 * RED_Y) * pstride + off];
            s3 += part[(size_t)(p + 3 * RED_Y) * pstride + off];
        }
        for (; p < nsplit; p += RED_Y) s0 += part[(size_t)p * pstride + off];
    }
    red_s[y][threadIdx.x] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    float s = 0.f;
    if (y == 0) {
#pragma unroll
        for (int q = 0; q < RED_Y; ++q) s += red_s[q][threadIdx.x];
    }
    return s;                                                  // valid on threadIdx.y == 0
}
#define CPG_RED_GRID(n) cpg::ceil_div((n), cpg::RED_X)
#define CPG_RED_BLOCK dim3(cpg::RED_X, cpg::RED_Y)

// ---- Philox4x32-10 counter RNG (Salmon et al. 2011), used by the perf-mode noise
// generators; key = (seed lo, seed hi), counter = (index lo, index hi, stream, 0).
struct Philox {
    static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    __host__ __device__ static inline void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
        uint64_t p = (uint64_t)a * b;
        hi = (uint32_t)(p >> 32);
        lo = (uint32_t)p;
    }
    __host__ __device__ static inline void gen(uint64_t seed, uint64_t index, uint32_t stream, uint32_t out[4]) {
        uint32_t c0 = (uint32_t)index, c1 = (uint32_t)(index >> 32), c2 = stream, c3 = 0;
        uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            uint32_t h0, l0, h1, l1;
            mulhilo(M0, c0, h0, l0);
            mulhilo(M1, c2, h1, l1);
            uint32_t n0 = h1 ^ c1 ^ k0, n1 = l1, n2 = h0 ^ c3 ^ k1, n3 = l0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            k0 += W0; k1 += W1;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
};
// uniform in (0,1): never 0 so that log() is finite
__host__ __device__ inline float u32_to_unit_open(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }
__host__ __device__ inline double u64_to_unit(uint32_t hi, uint32_t lo) {
    // 53-bit uniform in [0,1) like numpy's random_sample
    return (double)((((uint64_t)(hi >> 5)) << 26) | (uint64_t)(lo >> 6)) * (1.0 / 9007199254740992.0);
}

}  // namespace cpg
