// Random-feature MMD (losses.py:59-93) on the 5th-gen tensor cores.  The fp32 SIMT products of the feature map
// (z W, z_prior W: [B x 100] x [100 x R]) and of its gradient (G W^T: [B x R] x [R x 100]) sit on the lane that feeds the
// latent backward; they share the GPU with the recurrence kernels of the dependent chain, which leave ~20 SMs free --
// where SIMT FMA throughput made that lane the critical path of the iteration.  Here:
//
//   k_prep_rf_tiles   rf_w -> pre-split operand tiles in their shared-memory image (once per step; R in halves of 256)
//   k_rf_feat_tc      pre = x W (split fp16, three products) -> optional store of pre (the gradient needs it) ->
//                     phi = cos(pre / sigma + b) sqrt(2/R) -> column sums over the CTA's 64 rows (per TMEM-lane quadrant,
//                     fixed shuffle tree) -> partial [4 * CTAs][R]; k_rf_colsum_final (latent.cu) sums them in order
//   k_rf_grad_tc      G = -coef sin(pre / sigma + b) formed while the A operand is converted (split bf16) -> dz = G W^T
//
// NARROW persistent grids (RF_GRID CTAs, each looping over row tiles; 170+ KB of shared memory each): the kernels of the
// dependent chain occupy 128 of the 148 SMs with one CTA each, and a wide side kernel whose CTAs sit on SMs when the next
// chain kernel starts delays that kernel by the remaining run time of those CTAs (measured: decoder recurrence 69 -> 128 us
// with 64-CTA grids here).  16 CTAs fit beside the chain, and their footprint keeps them from sharing an SM -- and its
// tensor memory, of which the recurrences allocate all 512 columns -- with a chain CTA.
// cos / sin: explicit two-constant reduction to [-pi, pi], then the SFU approximation (abs. error 2^-21.4 there).
#include "ctx.h"
#ifndef CPG_EMU
#include "tc_dense.cuh"

namespace cpg {
int check_launch(const char* where);

namespace {
constexpr int RF_M = 64;                     // rows per tile of the gradient kernel
constexpr int RF_MF = 128;                   // rows per tile of the feature kernel
constexpr int RF_GRID = 16;                  // CTAs per kernel (see above)
constexpr int RF_KZ = 112;                   // latent dim 100 padded to a multiple of 16
constexpr int RF_NH = 256;                   // features per half
constexpr int RF_TERM = (RF_NH / 8) * (RF_KZ * 16);           // bytes per term of a half tile (either layout): 57,344
static_assert(RF_TERM == RF_KZ * RF_NH * 2, "tile size");

// tiles: [half][fwd: MN-major (n = feature fastest, K = latent) hi | lo][bwd: K-major (row = latent, K = feature) hi | lo]
__global__ void k_prep_rf_tiles(const float* __restrict__ rf_w, int R, int n_half, unsigned char* __restrict__ tiles) {
    const int per_half = 2 * (RF_NH / 8) * RF_KZ;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_half * per_half; i += gridDim.x * blockDim.x) {
        const int h = i / per_half, t = i % per_half;
        unsigned char* base = tiles + (size_t)h * 4 * RF_TERM;
        float x[8];
        uint4 hi, lo;
        if (t < per_half / 2) {                        // forward tile: chunk (nc, k) = W[k][h*256 + nc*8 .. +7]
            const int nc = t % (RF_NH / 8), k = t / (RF_NH / 8);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int r = h * RF_NH + nc * 8 + e;
                x[e] = (k < ZD && r < R) ? rf_w[(size_t)k * R + r] : 0.f;
            }
            split8x<true>(x, hi, lo);
            const size_t off = (size_t)nc * (RF_KZ * 16) + (k >> 3) * 128 + (k & 7) * 16;
            *reinterpret_cast<uint4*>(base + off) = hi;
            *reinterpret_cast<uint4*>(base + RF_TERM + off) = lo;
        } else {                                       // backward tile: chunk (n, kc) = W[n][h*256 + kc*8 .. +7]
            const int u = t - per_half / 2;
            const int kc = u % (RF_NH / 8), n = u / (RF_NH / 8);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int r = h * RF_NH + kc * 8 + e;
                x[e] = (n < ZD && r < R) ? rf_w[(size_t)n * R + r] : 0.f;
            }
            split8x<false>(x, hi, lo);
            const size_t off = (size_t)kc * (RF_KZ * 16) + (n >> 3) * 128 + (n & 7) * 16;
            *reinterpret_cast<uint4*>(base + 2 * RF_TERM + off) = hi;
            *reinterpret_cast<uint4*>(base + 3 * RF_TERM + off) = lo;
        }
    }
}

__device__ __forceinline__ float reduce_2pi(float x) {
    const float k = rintf(x * 0.15915494309189535f);
    x = fmaf(-k, 6.2831854820251465f, x);
    return fmaf(-k, -1.7484556e-7f, x);
}
__device__ __forceinline__ float cos_rr(float x) { return __cosf(reduce_2pi(x)); }
__device__ __forceinline__ float sin_rr(float x) { return __sinf(reduce_2pi(x)); }

struct RfFeatArgs {
    const float* x;             // [B][100]
    const unsigned char* tiles;
    const float* rf_b;          // [R]
    float* pre_out;             // [B][R] or null
    float* part;                // [4 * CTAs][R]
    float sigma;
    int B, R, n_half;
};
constexpr int FE_A = RF_MF * RF_KZ * 2;                        // bytes per term of the x tile
constexpr size_t FE_SMEM = 2 * (size_t)FE_A + 2 * (size_t)RF_TERM;

__global__ void __launch_bounds__(LT_THREADS, 1)
k_rf_feat_tc(RfFeatArgs a) {
    constexpr int M = RF_MF;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* A = smem;
    unsigned char* Bt = smem + 2 * FE_A;
    __shared__ __align__(8) uint64_t bar_mma, bar_w;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int B = a.B, R = a.R, n_tiles = ceil_div(B, M);
    if (warp == 0) {
        if (lane == 0) {
            tc::mbar_init(&bar_mma, 1);
            tc::mbar_init(&bar_w, 1);
            tc::fence_barrier_init();
            tc::mbar_expect_tx(&bar_w, 2 * RF_TERM);
            bulk_load(Bt, a.tiles, 2 * RF_TERM, &bar_w);
        }
        __syncwarp();
        tc::tmem_alloc<512>(&tmem_slot);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int q = warp & 3, part = warp >> 2;
    const float scale = sqrtf(2.0f / (float)R), inv_sigma = 1.0f / a.sigma;
    uint32_t it = 0;                                           // (tile, half) pairs processed by this CTA
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row0 = tile * M;
        // (all MMAs that read A are complete: every thread passed the last wait on bar_mma)
        fill_rows_kmajor<true, M, RF_KZ>(A, A + FE_A, a.x + (size_t)row0 * ZD, ZD, B - row0, ZD);
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
        const int row = row0 + q * 32 + lane;
        const bool live = row < B;
        for (int h = 0; h < a.n_half; ++h, ++it) {
            const uint32_t acc = tmem + (uint32_t)((it & 1) * RF_NH);
            if (tid == 0) {
                tc::mbar_wait(&bar_w, it & 1);
                const uint32_t a0 = tc::smem_u32(A), b0 = tc::smem_u32(Bt);
                issue_products(acc, M, a0, a0 + FE_A, M, b0, b0 + RF_TERM, true, RF_KZ, 0, RF_NH, RF_KZ, true, true);
                tc::umma_commit(&bar_mma);
            }
            tc::mbar_wait(&bar_mma, it & 1);
            tc::tc_fence_after();
            const bool more = h + 1 < a.n_half || tile + (int)gridDim.x < n_tiles;
            if (tid == 0 && more) {                            // the next weight half lands while this epilogue runs
                const int hn = h + 1 < a.n_half ? h + 1 : 0;
                tc::mbar_expect_tx(&bar_w, 2 * RF_TERM);
                bulk_load(Bt, a.tiles + (size_t)hn * 4 * RF_TERM, 2 * RF_TERM, &bar_w);
            }
            for (int c0 = part * 16; c0 < RF_NH; c0 += 16 * LT_PARTS) {
                const int r0 = h * RF_NH + c0;
                if (r0 >= R) break;
                float v[16];
                tmem_ld_cols<16>(acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                if (live && a.pre_out != nullptr) {
#pragma unroll
                    for (int e = 0; e < 16; e += 4)
                        if (r0 + e < R) st4(a.pre_out + (size_t)row * R + r0 + e, make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]));
                }
                float mine = 0.f;
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    float p = 0.f;
                    if (live && r0 + e < R) p = cos_rr(fmaf(v[e], inv_sigma, __ldg(a.rf_b + r0 + e))) * scale;
#pragma unroll
                    for (int o = 16; o >= 1; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
                    if (lane == e) mine = p;
                }
                if (lane < 16 && r0 + lane < R) a.part[(size_t)(tile * 4 + q) * R + r0 + lane] = mine;
            }
            tc::tc_fence_before();
            __syncthreads();                                   // this accumulator buffer is drained before it is reused
            tc::tc_fence_after();
        }
    }
    if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

struct RfGradArgs {
    const float* pre;           // [B][R] (z W, as stored by k_rf_feat_tc)
    const unsigned char* tiles;
    const float* rf_b; const float* coef;      // [R]
    float* dz;                  // [B][100]
    float sigma;
    int B, R, n_half;
};
constexpr int GR_A = RF_M * RF_NH * 2;                         // bytes per term of a G half tile
constexpr size_t GR_SMEM = 2 * (size_t)GR_A + 2 * (size_t)RF_TERM;

__global__ void __launch_bounds__(LT_THREADS, 1)
k_rf_grad_tc(RfGradArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* A = smem;
    unsigned char* Bt = smem + 2 * GR_A;
    __shared__ __align__(8) uint64_t bar_mma, bar_w;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int B = a.B, R = a.R, n_tiles = ceil_div(B, RF_M);
    if (warp == 0) {
        if (lane == 0) {
            tc::mbar_init(&bar_mma, 1);
            tc::mbar_init(&bar_w, 1);
            tc::fence_barrier_init();
        }
        __syncwarp();
        tc::tmem_alloc<128>(&tmem_slot);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int rl = lane & 7, fl = lane >> 3;
    const int q = warp & 3, part = warp >> 2;
    const float inv_sigma = 1.0f / a.sigma;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row0 = tile * RF_M;
        for (int h = 0; h < a.n_half; ++h, ++it) {
            if (tid == 0) {                                    // (the previous MMAs are complete: both buffers are free)
                tc::mbar_expect_tx(&bar_w, 2 * RF_TERM);
                bulk_load(Bt, a.tiles + (size_t)h * 4 * RF_TERM + 2 * RF_TERM, 2 * RF_TERM, &bar_w);
            }
            // A half: G[b][r] = -coef[r] sin(pre[b][r] / sigma + b[r]), K-major (rows = batch, K = feature), split bf16;
            // warp-task = 8 rows x 16 features, lane = (row in block, float4 of the 64-byte piece)
            constexpr int NT = (RF_M / 8) * (RF_NH / 16);
#pragma unroll 4
            for (int t = warp; t < NT; t += LT_THREADS / 32) {
                const int rb = t % (RF_M / 8), kg = t / (RF_M / 8);
                const int r = rb * 8 + rl, k0 = kg * 16 + fl * 4, f0 = h * RF_NH + k0;
                float x[4] = {0.f, 0.f, 0.f, 0.f};
                if (row0 + r < B && f0 < R) {
                    const float4 p = ld_stream4(a.pre + (size_t)(row0 + r) * R + f0);
                    const float4 cf = ldg4(a.coef + f0), bb = ldg4(a.rf_b + f0);
                    x[0] = -cf.x * sin_rr(fmaf(p.x, inv_sigma, bb.x));
                    x[1] = -cf.y * sin_rr(fmaf(p.y, inv_sigma, bb.y));
                    x[2] = -cf.z * sin_rr(fmaf(p.z, inv_sigma, bb.z));
                    x[3] = -cf.w * sin_rr(fmaf(p.w, inv_sigma, bb.w));
                }
                uint2 hi, lo;
                split4x<false>(x, hi, lo);
                const int off = (k0 >> 3) * (RF_M * 16) + rb * 128 + rl * 16 + (k0 & 7) * 2;
                *reinterpret_cast<uint2*>(A + off) = hi;
                *reinterpret_cast<uint2*>(A + GR_A + off) = lo;
            }
            tc::fence_proxy_async();
            tc::tc_fence_before();
            __syncthreads();                                   // (also: the previous tile's accumulator has been read out)
            tc::tc_fence_after();
            if (tid == 0) {
                tc::mbar_wait(&bar_w, it & 1);
                const uint32_t a0 = tc::smem_u32(A), b0 = tc::smem_u32(Bt);
                issue_products(tmem, RF_M, a0, a0 + GR_A, RF_M, b0, b0 + RF_TERM, false, RF_KZ, 0, RF_KZ, RF_NH, h == 0, false);
                tc::umma_commit(&bar_mma);
            }
            tc::mbar_wait(&bar_mma, it & 1);
            tc::tc_fence_after();
        }
        bool has;
        const int r_ = epi_row<RF_M>(q, lane, has), row = row0 + r_;
        const bool live = has && row < B;
        for (int c0 = part * 16; c0 < RF_KZ; c0 += 16 * LT_PARTS) {
            float v[16];
            tmem_ld_cols<16>(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            if (live) {
#pragma unroll
                for (int e = 0; e < 16; e += 4)
                    if (c0 + e < ZD) st4(a.dz + (size_t)row * ZD + c0 + e, make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]));
            }
        }
        tc::tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<128>(tmem);
}

bool set_smem(const void* fn, size_t bytes) {
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) { cudaGetLastError(); return false; }
    return true;
}
}  // namespace

int g_opt_rf_grid = 0;        // CTAs of the RF kernels (0 = RF_GRID)
int g_opt_rf_tc = 1;          // 1: tcgen05 random-feature kernels when B >= 512, 2: always, 0: fp32 SIMT GEMMs
bool rf_uses_tc(int B, int R) { return R % 4 == 0 && R >= 16 && (g_opt_rf_tc == 2 || (g_opt_rf_tc == 1 && B >= 512)); }
int rf_tc_halves(int R) { return ceil_div(R, RF_NH); }
size_t rf_tc_tile_bytes(int R) { return (size_t)rf_tc_halves(R) * 4 * RF_TERM; }
int rf_tc_parts(int B) { return 4 * ceil_div(B, RF_MF); }

void launch_prep_rf_tiles(cudaStream_t s, const float* rf_w, int R, unsigned char* tiles) {
    const int nh = rf_tc_halves(R);
    CPG_LAUNCH(k_prep_rf_tiles, ceil_div(nh * 2 * (RF_NH / 8) * RF_KZ, 256), 256, 0, s, rf_w, R, nh, tiles);
}

int launch_rf_feat_tc(cudaStream_t s, const float* x, const unsigned char* tiles, const float* rf_b, int B, int R, float sigma,
                      float* pre_out, float* part) {
    static bool set = false;
    if (!set) {
        if (!set_smem((const void*)k_rf_feat_tc, FE_SMEM)) return CPG_ECUDA;
        set = true;
    }
    RfFeatArgs a{x, tiles, rf_b, pre_out, part, sigma, B, R, rf_tc_halves(R)};
    CPG_LAUNCH(k_rf_feat_tc, std::min(g_opt_rf_grid > 0 ? g_opt_rf_grid : RF_GRID, ceil_div(B, RF_MF)), LT_THREADS, FE_SMEM, s, a);
    return CPG_OK;
}

int launch_rf_grad_tc(cudaStream_t s, const float* pre, const unsigned char* tiles, const float* rf_b, const float* coef, int B, int R,
                      float sigma, float* dz) {
    static bool set = false;
    if (!set) {
        if (!set_smem((const void*)k_rf_grad_tc, GR_SMEM)) return CPG_ECUDA;
        set = true;
    }
    RfGradArgs a{pre, tiles, rf_b, coef, dz, sigma, B, R, rf_tc_halves(R)};
    CPG_LAUNCH(k_rf_grad_tc, std::min(g_opt_rf_grid > 0 ? g_opt_rf_grid : RF_GRID, ceil_div(B, RF_M)), LT_THREADS, GR_SMEM, s, a);
    return CPG_OK;
}

}  // namespace cpg
#else
namespace cpg {
int g_opt_rf_tc = 0;
bool rf_uses_tc(int, int) { return false; }
int rf_tc_halves(int) { return 1; }
size_t rf_tc_tile_bytes(int) { return 16; }
int rf_tc_parts(int) { return 1; }
void launch_prep_rf_tiles(cudaStream_t, const float*, int, unsigned char*) {}
int launch_rf_feat_tc(cudaStream_t, const float*, const unsigned char*, const float*, int, int, float, float*, float*) { return CPG_ECUDA; }
int launch_rf_grad_tc(cudaStream_t, const float*, const unsigned char*, const float*, const float*, int, int, float, float*) { return CPG_ECUDA; }
}  // namespace cpg
#endif
