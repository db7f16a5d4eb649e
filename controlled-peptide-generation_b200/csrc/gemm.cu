// Small fp32 SIMT GEMM used for the dense layers around the recurrences
// (q_mu / q_logvar heads, [z;c] input projection, random-feature map and their
// backward contractions).  These are 1-3 % of the step's FLOPs; the kernel is a
// plain 64x64x16 register-tiled SGEMM with general operand strides and a
// deterministic split-K (partials + ordered reduction, no float atomics).
#include "kernels.h"

namespace cpg {

constexpr int GM = 64, GN = 64, GK = 16;

__global__ void __launch_bounds__(256)
k_sgemm(int M, int N, int K, float alpha, const float* __restrict__ A, int64_t sam, int64_t sak,
        const float* __restrict__ Bm, int64_t sbk, int64_t sbn, float beta, float* __restrict__ C, int64_t ldc,
        const float* __restrict__ bias, int kchunk, float* __restrict__ ws) {
    __shared__ __align__(16) float As[GK][GM + 4];
    __shared__ __align__(16) float Bs[GK][GN + 4];
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
    const int kbeg = blockIdx.z * kchunk;
    const int kend = min(K, kbeg + kchunk);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const bool a_kfast = (sak == 1);
    const bool b_nfast = (sbn == 1);
    for (int k0 = kbeg; k0 < kend; k0 += GK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int idx = tid + 256 * i;
            int m, k;
            if (a_kfast) { k = idx % GK; m = idx / GK; } else { m = idx % GM; k = idx / GM; }
            int gm = m0 + m, gk = k0 + k;
            As[k][m] = (gm < M && gk < kend) ? A[gm * sam + gk * sak] : 0.f;
            int n, kb;
            if (b_nfast) { n = idx % GN; kb = idx / GN; } else { kb = idx % GK; n = idx / GK; }
            int gn = n0 + n, gkb = k0 + kb;
            Bs[kb][n] = (gn < N && gkb < kend) ? Bm[gkb * sbk + gn * sbn] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GK; ++k) {
            const float4 av = ld4(&As[k][ty * 4]);
            const float4 bv = ld4(&Bs[k][tx * 4]);
            const float a4[4] = {av.x, av.y, av.z, av.w};
            const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
        }
        __syncthreads();
    }
    const bool split = gridDim.z > 1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int gn = n0 + tx * 4 + j;
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
