// fp32 SIMT GEMM used for the dense layers around the recurrences (q_mu / q_logvar heads, [z;c]
// input projection, random-feature map and their backward contractions; ~6 % of the step's FLOPs).
// 128x64x16 tiles, 256 threads, 8x4 register tile per thread, register-prefetched (double-buffered)
// global loads, general operand strides, and a deterministic split-K (partials + ordered
// reduction, no float atomics).  These stay in exact fp32 because the forward ones (mu, logvar,
// the [z;c] projection) feed results that are compared at 1e-4.
#include "kernels.h"

namespace cpg {

constexpr int GM = 128, GN = 64, GK = 16;

__global__ void __launch_bounds__(256)
k_sgemm(int M, int N, int K, float alpha, const float* __restrict__ A, int64_t sam, int64_t sak,
        const float* __restrict__ Bm, int64_t sbk, int64_t sbn, float beta, float* __restrict__ C, int64_t ldc,
        const float* __restrict__ bias, int kchunk, float* __restrict__ ws) {
    __shared__ __align__(16) float As[2][GK][GM + 4];
    __shared__ __align__(16) float Bs[2][GK][GN + 4];
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;            // 16 column groups (4 cols) x 16 row groups (8 rows)
    const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
    const int kbeg = blockIdx.z * kchunk;
    const int kend = min(K, kbeg + kchunk);
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    // element -> (m, k) / (k, n) assignment of this thread for the global->shared copies:
    // the fastest-varying index follows the operand's unit stride so that loads coalesce
    const bool a_kfast = (sak == 1);
    const bool b_nfast = (sbn == 1);
    float ra[8], rb[4];
    auto load_tile = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int idx = tid + 256 * i;
            int m, k;
            if (a_kfast) { k = idx % GK; m = idx / GK; } else { m = idx % GM; k = idx / GM; }
            const int gm = m0 + m, gk = k0 + k;
            ra[i] = (gm < M && gk < kend) ? A[gm * sam + gk * sak] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + 256 * i;
            int n, k;
            if (b_nfast) { n = idx % GN; k = idx / GN; } else { k = idx % GK; n = idx / GK; }
            const int gn = n0 + n, gk = k0 + k;
            rb[i] = (gn < N && gk < kend) ? Bm[gk * sbk + gn * sbn] : 0.f;
        }
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int idx = tid + 256 * i;
            int m, k;
            if (a_kfast) { k = idx % GK; m = idx / GK; } else { m = idx % GM; k = idx / GM; }
            As[buf][k][m] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + 256 * i;
            int n, k;
            if (b_nfast) { n = idx % GN; k = idx / GN; } else { k = idx % GK; n = idx / GK; }
            Bs[buf][k][n] = rb[i];
        }
    };

    int buf = 0;
    if (kbeg < kend) {
        load_tile(kbeg);
        store_tile(0);
    }
    __syncthreads();
    for (int k0 = kbeg; k0 < kend; k0 += GK) {
        const bool more = k0 + GK < kend;
        if (more) load_tile(k0 + GK);                   // global loads in flight during the FMAs below
#pragma unroll
        for (int k = 0; k < GK; ++k) {
            const float4 a0 = ld4(&As[buf][k][ty * 8]);
            const float4 a1 = ld4(&As[buf][k][ty * 8 + 4]);
            const float4 bv = ld4(&Bs[buf][k][tx * 4]);
            const float a8[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a8[i], b4[j], acc[i][j]);
        }
        if (more) store_tile(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }
    const bool split = gridDim.z > 1;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int gm = m0 + ty * 8 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            if (split) {
                ws[((size_t)blockIdx.z * M + gm) * N + gn] = acc[i][j];
            } else {
                float v = alpha * acc[i][j];
                if (bias != nullptr) v += bias[gn];
                if (beta != 0.f) v += beta * C[gm * ldc + gn];
                C[gm * ldc + gn] = v;
            }
        }
    }
}

__global__ void k_splitk_reduce(int M, int N, int splits, float alpha, const float* __restrict__ ws, float beta,
                                float* __restrict__ C, int64_t ldc, const float* __restrict__ bias) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * N) return;
    int m = i / N, n = i % N;
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += ws[(size_t)z * M * N + i];
    float v = alpha * s;
    if (bias != nullptr) v += bias[n];
    if (beta != 0.f) v += beta * C[m * ldc + n];
    C[m * ldc + n] = v;
}

void launch_sgemm(cudaStream_t s, int M, int N, int K, float alpha, const float* A, int64_t sam, int64_t sak,
                  const float* B, int64_t sbk, int64_t sbn, float beta, float* C, int64_t ldc,
                  const float* bias, int split_k, float* ws) {
    if (split_k < 1 || ws == nullptr) split_k = 1;
    int kchunk = ceil_div(ceil_div(K, split_k), GK) * GK;
    split_k = ceil_div(K, kchunk);
    dim3 grid(ceil_div(N, GN), ceil_div(M, GM), split_k);
    CPG_LAUNCH(k_sgemm, grid, 256, 0, s, M, N, K, alpha, A, sam, sak, B, sbk, sbn, beta, C, ldc, bias, kchunk, ws);
    if (split_k > 1)
        CPG_LAUNCH(k_splitk_reduce, ceil_div(M * N, 256), 256, 0, s, M, N, split_k, alpha, ws, beta, C, ldc, bias);
}

// column sums, two deterministic stages: partial[chunk][n] then ordered sum over chunks
__global__ void k_colsum_partial(const float* __restrict__ A, int M, int N, int64_t lda, int rows_per_chunk,
                                 float* __restrict__ part) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    int c = blockIdx.y;
    if (n >= N) return;
    int mbeg = c * rows_per_chunk, mend = min(M, mbeg + rows_per_chunk);
    float s = 0.f;
    for (int m = mbeg; m < mend; ++m) s += A[m * lda + n];
    part[(size_t)c * N + n] = s;
}
__global__ void k_colsum_final(const float* __restrict__ part, int nchunk, int N, float* __restrict__ out) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float s = 0.f;
    for (int c = 0; c < nchunk; ++c) s += part[(size_t)c * N + n];
    out[n] = s;
}

void launch_colsum(cudaStream_t s, const float* A, int M, int N, int64_t lda, float* out, float* ws, int nchunk) {
    nchunk = max(1, min(nchunk, M));
    int rpc = ceil_div(M, nchunk);
    nchunk = ceil_div(M, rpc);
    CPG_LAUNCH(k_colsum_partial, dim3(ceil_div(N, 128), nchunk), 128, 0, s, A, M, N, lda, rpc, ws);
    CPG_LAUNCH(k_colsum_final, ceil_div(N, 128), 128, 0, s, ws, nchunk, N, out);
}

}  // namespace cpg
