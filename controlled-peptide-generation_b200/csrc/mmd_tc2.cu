// Full-kernel MMD Gram sum, persistent warp-specialised tcgen05 version.
//
// Same math as mmd_tc.cu (X = tf32-rounded [z ; z_prior] padded to K = 128; sum over the upper
// triangle of s_i s_j exp(-(|x_i|^2 + |x_j|^2 - 2 x_i.x_j) / sigma^2)), but instead of one tile per
// CTA each CTA walks a list of work items (row tile ti, up to 8 consecutive column tiles tj >= ti):
//   warp 4  : TMA producer  -- A row tile once per item, B column tiles double-buffered in smem
//   warp 5  : MMA issuer    -- 13 tcgen05.mma (M=128, N=128, K=8, tf32) per tile into one of two
//                              TMEM accumulators, tcgen05.commit to the epilogue and back to the producer
//   warps 0-3: epilogue     -- tcgen05.ld the finished accumulator while the next tile is being
//                              multiplied; per element  ex2(min(c*dot + a_i + b_j, 0)) * s_j  (4 instructions)
// so TMA, tensor core and the exp epilogue overlap.  Diagonal tiles take a slower exact path
// (d^2 = 0 on the diagonal, strict upper triangle weighted 2).
#include "ctx.h"
#ifndef CPG_EMU
#include "tc_common.cuh"

namespace cpg {
int check_launch(const char* where);
__global__ void k_mmd_pack(const float* __restrict__ z, const float* __restrict__ zp, int N, float sigma,
                           float* __restrict__ X, float* __restrict__ norms, float* __restrict__ diag_part);
__global__ void k_mmd_final_tc(const float* __restrict__ part, int nparts, const float* __restrict__ diag_part,
                               int ndiag, int N, float* __restrict__ out);

constexpr int G2_T = 128;                 // tile edge
constexpr int G2_KSTEPS = 13;             // ceil(100 / 8)
constexpr int G2_CHUNK = 8;               // column tiles per work item
constexpr int G2_TILE_BYTES = 4 * G2_T * 128;      // 4 k-blocks of [128 rows][128 B]

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// item index -> (ti, first tj, number of tiles)
__device__ __forceinline__ void g2_item(int item, int T, int& ti, int& tj0, int& cnt) {
    ti = 0;
    for (;;) {
        const int per = (T - ti + G2_CHUNK - 1) / G2_CHUNK;
        if (item < per) break;
        item -= per;
        ++ti;
    }
    tj0 = ti + item * G2_CHUNK;
    cnt = min(G2_CHUNK, T - tj0);
}

__global__ void __launch_bounds__(192, 1)
k_mmd_gram_tc2(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ norms, int N, float sigma, int T,
               int n_items, float* __restrict__ part) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* sA = smem;
    unsigned char* sB = smem + G2_TILE_BYTES;               // 2 stages
    __shared__ __align__(8) uint64_t full_b[2], empty_b[2], tmem_full[2], tmem_empty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float b2_s[2][G2_T], sg_s[2][G2_T];
    __shared__ float red[4];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int M2 = 2 * N;

    if (warp == 5) {
        if (lane == 0) {
            for (int s = 0; s < 2; ++s) {
                tc::mbar_init(&full_b[s], 1);
                tc::mbar_init(&empty_b[s], 1);
                tc::mbar_init(&tmem_full[s], 1);
                tc::mbar_init(&tmem_empty[s], 128);
            }
            tc::fence_barrier_init();
        }
        __syncwarp();
        tc::tmem_alloc<256>(&tmem_slot);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = tmem_slot;

    if (warp == 4) {
        // ---------------- TMA producer
        if (lane == 0) {
            tc::tma_prefetch_desc(&tmap);
            int n = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                int ti, tj0, cnt;
                g2_item(item, T, ti, tj0, cnt);
                for (int j = 0; j < cnt; ++j, ++n) {
                    const int s = n & 1, u = n >> 1;
                    if (u >= 1) tc::mbar_wait(&empty_b[s], (u - 1) & 1);
                    const bool new_a = (j == 0);
                    if (new_a && n >= 1) tc::mbar_wait(&empty_b[(n - 1) & 1], ((n - 1) >> 1) & 1);   // A still read by tile n-1
                    tc::mbar_expect_tx(&full_b[s], (new_a ? 2u : 1u) * G2_TILE_BYTES);
                    unsigned char* sb = sB + (size_t)s * G2_TILE_BYTES;
                    for (int kb = 0; kb < 4; ++kb) {
                        if (new_a) tc::tma_load_2d(sA + kb * G2_T * 128, &tmap, &full_b[s], kb * 32, ti * G2_T);
                        tc::tma_load_2d(sb + kb * G2_T * 128, &tmap, &full_b[s], kb * 32, (tj0 + j) * G2_T);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        // ---------------- MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = tc::make_idesc_tf32(G2_T, G2_T, 0, 0);
            int n = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                int ti, tj0, cnt;
                g2_item(item, T, ti, tj0, cnt);
                for (int j = 0; j < cnt; ++j, ++n) {
                    const int s = n & 1, u = n >> 1;
                    tc::mbar_wait(&full_b[s], u & 1);
                    if (u >= 1) tc::mbar_wait(&tmem_empty[s], (u - 1) & 1);
                    tc::tc_fence_after();
                    const uint32_t a0 = tc::smem_u32(sA), b0 = tc::smem_u32(sB + (size_t)s * G2_TILE_BYTES);
#pragma unroll
                    for (int ks = 0; ks < G2_KSTEPS; ++ks) {
                        const int kb = ks >> 2, q = ks & 3;
                        const uint64_t da = tc::make_smem_desc_sw128(a0 + kb * G2_T * 128 + q * 32, 16, 1024);
                        const uint64_t db = tc::make_smem_desc_sw128(b0 + kb * G2_T * 128 + q * 32, 16, 1024);
                        tc::umma_tf32(tmem_d + s * G2_T, da, db, idesc, ks > 0 ? 1u : 0u);
                    }
                    tc::umma_commit(&tmem_full[s]);
                    tc::umma_commit(&empty_b[s]);
                }
            }
        }
        __syncwarp();
    } else {
        // ---------------- epilogue (128 threads, thread = accumulator row)
        const float kk = 1.4426950408889634f / (sigma * sigma);    // log2(e) / sigma^2
        const float c2 = 2.0f * kk;
        float total = 0.f;
        int n = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            int ti, tj0, cnt;
            g2_item(item, T, ti, tj0, cnt);
            const int gi = ti * G2_T + tid;
            const bool row_ok = gi < M2;
            const float a2 = row_ok ? -norms[gi] * kk : 0.f;
            const float si = gi < N ? 1.f : -1.f;
            float acc_off = 0.f, acc_diag = 0.f;
            for (int j = 0; j < cnt; ++j, ++n) {
                const int s = n & 1, u = n >> 1;
                const int tj = tj0 + j;
                {
                    const int gj = tj * G2_T + tid;
                    b2_s[s][tid] = gj < M2 ? -norms[gj] * kk : 0.f;
                    sg_s[s][tid] = gj < M2 ? (gj < N ? 1.f : -1.f) : 0.f;
                }
                epi_bar_sync();
                tc::mbar_wait(&tmem_full[s], u & 1);
                tc::tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < G2_T / 32; ++c) {
                    float v[32];
                    tc::tmem_ld_32x32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(s * G2_T + c * 32), v);
                    if (tj != ti) {
#pragma unroll
                        for (int q = 0; q < 32; ++q) {
                            const float arg = fminf(fmaf(v[q], c2, a2 + b2_s[s][c * 32 + q]), 0.f);
                            acc_off = fmaf(ex2_approx(arg), sg_s[s][c * 32 + q], acc_off);
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 32; ++q) {
                            const int col = c * 32 + q;
                            const float arg = fminf(fmaf(v[q], c2, a2 + b2_s[s][col]), 0.f);
                            const float kv = ex2_approx(arg) * sg_s[s][col];
                            if (col > tid) acc_off += kv;            // strict upper triangle (weight 2 below)
                            else if (col == tid) acc_diag += sg_s[s][col];   // K(x_i, x_i) = 1 exactly
                        }
                    }
                }
                tc::tc_fence_before();
                tc::mbar_arrive(&tmem_empty[s]);
            }
            if (row_ok) total += si * (2.0f * acc_off + acc_diag);
        }
        total = warp_sum(total);
        if (lane == 0) red[warp] = total;
    }
    tc::tc_fence_before();
    __syncthreads();
    if (tid == 0) part[blockIdx.x] = red[0] + red[1] + red[2] + red[3];
    if (warp == 5) tc::tmem_dealloc<256>(tmem_d);
}

int g_opt_mmd_grid = 0;      // CTAs of the persistent Gram kernel (0 = one per SM)
int launch_mmd_full_tc2(cudaStream_t s, const float* z, const float* zp, int N, float sigma, int sm_count, float* ws,
                        float* out) {
    const int M2 = 2 * N;
    const int T = ceil_div(M2, G2_T);
    float* X = ws;
    float* norms = X + (size_t)M2 * 128;
    float* diag_part = norms + M2;
    float* part = diag_part + 256;
    int ndiag = std::max(1, std::min(256, ceil_div(N, 8)));
    int n_items = 0;
    for (int ti = 0; ti < T; ++ti) n_items += ceil_div(T - ti, G2_CHUNK);
    const int grid = std::max(1, std::min(n_items, g_opt_mmd_grid > 0 ? g_opt_mmd_grid : sm_count));
    CUtensorMap tmap;
    int rc = make_tmap_2d_f32_sw128(&tmap, X, 128, (uint64_t)M2, (uint64_t)128 * sizeof(float), G2_T);
    if (rc) return rc;
    CPG_LAUNCH(k_mmd_pack, ndiag, 256, 0, s, z, zp, N, sigma, X, norms, diag_part);
    const size_t smem = 3 * (size_t)G2_TILE_BYTES + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute((const void*)k_mmd_gram_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    CPG_LAUNCH(k_mmd_gram_tc2, grid, 192, smem, s, tmap, norms, N, sigma, T, n_items, part);
    CPG_LAUNCH(k_mmd_final_tc, 1, 256, 0, s, part, grid, diag_part, ndiag, N, out);
    return CPG_OK;
}

}  // namespace cpg
#endif  // CPG_EMU
