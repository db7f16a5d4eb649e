// Full-kernel MMD Gram sum on the 5th-gen tensor cores (tcgen05, kind::tf32, TMEM accumulators,
// TMA-staged 128-byte-swizzled operands).
//
// Reference: losses.py:47-56,96-108 builds three [B,B,100] broadcasts (6.7 GB each at B=4096, 66 %
// of the reference's CPU step).  Here X = [z ; z_prior] (2B rows, K padded 100 -> 128, rounded to
// tf32) and  sum(H) = sum_{i,j} s_i s_j exp(-(|x_i|^2 + |x_j|^2 - 2 x_i.x_j) / sigma^2):
// the cross term x_i.x_j is one 128x128x104 UMMA tile per CTA (13 tcgen05.mma of K=8), the norms stay
// fp32, and the exp + signed sum is the TMEM epilogue.  Only upper-triangular tiles are computed.
// The diagonal term K(z_j, zp_j) and the norms come from k_mmd_pack in exact fp32.
#include "ctx.h"
#ifndef CPG_EMU
#include "tc_common.cuh"

namespace cpg {
int check_launch(const char* where);

constexpr int TCM = 128;              // tile rows (i) and columns (j)
constexpr int TCK = 128;              // padded K (floats); 4 swizzle atoms of 32 floats
constexpr int TC_KSTEPS = 13;         // ceil(100 / 8) MMAs of K = 8

int make_tmap_2d_f32_sw128(CUtensorMap* out, const float* base, uint64_t cols, uint64_t rows, uint64_t row_stride_bytes,
                           uint32_t box_rows, bool atom32) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || p == nullptr) {
            set_error("cuTensorMapEncodeTiled entry point not available");
            return CPG_ECUDA;
        }
        fn = (EncodeFn)p;
    }
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {row_stride_bytes};
    cuuint32_t box[2] = {32, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
        return CPG_ECUDA;
    }
    return CPG_OK;
}

// X[2N][128] = tf32-rounded [z ; zp] (zero padded), fp32 squared norms of the ROUNDED rows, and the
// per-CTA partial of sum_j exp(-|z_j - zp_j|^2 / sigma^2) (exact fp32 differences).
__global__ void __launch_bounds__(256)
k_mmd_pack(const float* __restrict__ z, const float* __restrict__ zp, int N, float sigma, float* __restrict__ X,
           float* __restrict__ norms, float* __restrict__ diag_part) {
    __shared__ float red[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float dacc = 0.f;
    for (int b = blockIdx.x * 8 + warp; b < N; b += gridDim.x * 8) {
        float n1 = 0.f, n2 = 0.f, dd = 0.f;
        for (int j = lane; j < TCK; j += 32) {
            float a = 0.f, c = 0.f;
            if (j < ZD) {
                float ra = z[(size_t)b * ZD + j], rc = zp[(size_t)b * ZD + j];
                dd += (ra - rc) * (ra - rc);
                a = tc::to_tf32_rn(ra);
                c = tc::to_tf32_rn(rc);
            }
            n1 += a * a; n2 += c * c;
            X[(size_t)b * TCK + j] = a;
            X[(size_t)(N + b) * TCK + j] = c;
        }
        n1 = warp_sum(n1); n2 = warp_sum(n2); dd = warp_sum(dd);
        if (lane == 0) { norms[b] = n1; norms[N + b] = n2; dacc += expf(-dd / (sigma * sigma)); }
    }
    if (lane == 0) red[warp] = dacc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += red[w];
        diag_part[blockIdx.x] = s;
    }
}

// One 128 x 128 Gram tile per CTA.  Warps 0-3: epilogue (thread = TMEM lane = tile row);
// warp 4: barrier init, TMEM alloc, TMA producer and MMA issuer (one elected lane).
__global__ void __launch_bounds__(160, 1)
k_mmd_gram_tc(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ norms, int N, float sigma,
              float* __restrict__ part) {
    const int ti = blockIdx.y, tj = blockIdx.x;
    const int pidx = blockIdx.y * gridDim.x + blockIdx.x;
    if (tj < ti) { if (threadIdx.x == 0) part[pidx] = 0.f; return; }

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    float* sA = reinterpret_cast<float*>(smem);                       // 4 k-blocks x [128 rows][32 floats]
    float* sB = reinterpret_cast<float*>(smem + 4 * TCM * 128);       // same for the j rows
    __shared__ __align__(8) uint64_t bar_full, bar_mma;
    __shared__ uint32_t tmem_slot;
    __shared__ float nj_s[TCM];
    __shared__ float red[4];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int M2 = 2 * N;

    if (warp == 4) {
        if (lane == 0) {
            tc::tma_prefetch_desc(&tmap);
            tc::mbar_init(&bar_full, 1);
            tc::mbar_init(&bar_mma, 1);
            tc::fence_barrier_init();
        }
        __syncwarp();
        tc::tmem_alloc<TCM>(&tmem_slot);
    } else if (tid < TCM) {
        int gj = tj * TCM + tid;
        nj_s[tid] = gj < M2 ? norms[gj] : 0.f;
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = tmem_slot;

    if (warp == 4) {
        if (lane == 0) {
            tc::mbar_expect_tx(&bar_full, 2u * 4u * TCM * 128u);
            for (int kb = 0; kb < 4; ++kb) {
                tc::tma_load_2d(sA + kb * TCM * 32, &tmap, &bar_full, kb * 32, ti * TCM);
                tc::tma_load_2d(sB + kb * TCM * 32, &tmap, &bar_full, kb * 32, tj * TCM);
            }
            tc::mbar_wait(&bar_full, 0);
            tc::tc_fence_after();
            constexpr uint32_t idesc = tc::make_idesc_tf32(TCM, TCM, 0, 0);
            for (int ks = 0; ks < TC_KSTEPS; ++ks) {
                const int kb = ks >> 2, s = ks & 3;
                const uint64_t da = tc::make_smem_desc_sw128(tc::smem_u32(sA + kb * TCM * 32) + s * 32, 16, 1024);
                const uint64_t db = tc::make_smem_desc_sw128(tc::smem_u32(sB + kb * TCM * 32) + s * 32, 16, 1024);
                tc::umma_tf32(tmem_d, da, db, idesc, ks > 0 ? 1u : 0u);
            }
            tc::umma_commit(&bar_mma);
        }
        __syncwarp();
    } else {
        tc::mbar_wait(&bar_mma, 0);
        tc::tc_fence_after();
        const int gi = ti * TCM + tid;
        const float ni = gi < M2 ? norms[gi] : 0.f;
        const float si = gi < N ? 1.f : -1.f;
        const float inv_s2 = 1.0f / (sigma * sigma);
        float tsum = 0.f;
#pragma unroll 1
        for (int c = 0; c < TCM / 32; ++c) {
            float v[32];
            tc::tmem_ld_32x32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), v);
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                const int j = c * 32 + q;
                const int gj = tj * TCM + j;
                if (gi < M2 && gj < M2 && !(ti == tj && gj < gi)) {
                    float d2 = fmaxf(ni + nj_s[j] - 2.0f * v[q], 0.f);
                    if (gi == gj) d2 = 0.f;
                    const float kv = __expf(-d2 * inv_s2);
                    const float sj = gj < N ? 1.f : -1.f;
                    tsum += (gi == gj ? 1.f : 2.f) * si * sj * kv;
                }
            }
        }
        tsum = warp_sum(tsum);
        if (lane == 0) red[warp] = tsum;
    }
    tc::tc_fence_before();
    __syncthreads();
    if (tid == 0) part[pidx] = red[0] + red[1] + red[2] + red[3];
    if (warp == 4) tc::tmem_dealloc<TCM>(tmem_d);
}

__global__ void k_mmd_final_tc(const float* __restrict__ part, int nparts, const float* __restrict__ diag_part,
                               int ndiag, int N, float* __restrict__ out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += (double)part[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double hs = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) hs += red[i];
        double k12d = 0.0;
        for (int i = 0; i < ndiag; ++i) k12d += (double)diag_part[i];
        *out = (float)((hs - (double)N * (2.0 * N - 2.0 * k12d)) / ((double)N * (double)(N - 1)));
    }
}

size_t mmd_tc_ws_floats(int N) {
    int nt = ceil_div(2 * N, TCM);
    return (size_t)2 * N * TCK + (size_t)2 * N + 256 + (size_t)nt * nt + 64;
}

int launch_mmd_full_tc(cudaStream_t s, const float* z, const float* zp, int N, float sigma, float* ws, float* out) {
    const int M2 = 2 * N;
    const int nt = ceil_div(M2, TCM);
    float* X = ws;                                  // 512-byte aligned by the workspace allocator
    float* norms = X + (size_t)M2 * TCK;
    float* diag_part = norms + M2;
    float* part = diag_part + 256;
    int ndiag = std::max(1, std::min(256, ceil_div(N, 8)));
    CUtensorMap tmap;
    int rc = make_tmap_2d_f32_sw128(&tmap, X, TCK, (uint64_t)M2, (uint64_t)TCK * sizeof(float), TCM);
    if (rc) return rc;
    CPG_LAUNCH(k_mmd_pack, ndiag, 256, 0, s, z, zp, N, sigma, X, norms, diag_part);
    const size_t smem = 2 * 4 * TCM * 128 + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute((const void*)k_mmd_gram_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    CPG_LAUNCH(k_mmd_gram_tc, dim3(nt, nt), 160, smem, s, tmap, norms, N, sigma, part);
    CPG_LAUNCH(k_mmd_final_tc, 1, 256, 0, s, part, nt * nt, diag_part, ndiag, N, out);
    return CPG_OK;
}

}  // namespace cpg
#endif  // CPG_EMU
