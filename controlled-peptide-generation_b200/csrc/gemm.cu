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

struct GemmSecond { const float* A; const float* B; const float* bias; float* C; int mode; };   // second operand set of a launch

#ifndef CPG_EMU   // cp.async pipeline: GPU build only (the CPU emulation of tools/cuda_emu keeps the plain kernel above)
// ---- pipelined variant for 16-byte-aligned operands (every product of the training step) -------------
// 64x64x16 tiles (B = 4096 gives 128..512 CTAs for the step's shapes instead of 64..256), 4x4 register
// tile, and a 4-stage cp.async pipeline so that three k-tiles of global latency are always in flight
// (the K loops here are only 7..32 tiles long: with one tile of prefetch the kernel was latency-bound).
// Each operand keeps its own fastest axis in shared memory (16-byte cp.async chunks along it); the FMA
// loop consumes four k's at a time so that either orientation is read with float4.
constexpr int PM = 64, PN = 64, PK = 16, PST = 4, PPAD = 4;
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;"
                 ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }


template <bool A_KFAST, bool B_KFAST>
__global__ void __launch_bounds__(256)
k_sgemm_pipe(int M, int N, int K, float alpha, const float* __restrict__ A, int64_t lda,
             const float* __restrict__ Bm, int64_t ldb, float beta, float* __restrict__ C, int64_t ldc,
             const float* __restrict__ bias, int kchunk, float* __restrict__ ws, GemmSecond sec) {
    // second operand set (same shapes and strides): sec.mode 1 -> C = alpha (A B + A2 B2) + ..., one pass more over K;
    // sec.mode 2 -> blockIdx.z == 1 computes the independent product C2 = alpha A2 B2 + bias2 (no split-K)
    if (sec.mode == 2 && blockIdx.z == 1) { A = sec.A; Bm = sec.B; C = sec.C; bias = sec.bias; }
    const int npass = sec.mode == 1 ? 2 : 1;
    // A_KFAST: A(m,k) = A[m*lda + k] -> As[m][k]   else A(m,k) = A[k*lda + m] -> As[k][m]
    // B_KFAST: B(k,n) = B[n*ldb + k] -> Bs[n][k]   else B(k,n) = B[k*ldb + n] -> Bs[k][n]
    constexpr int AROWS = A_KFAST ? PM : PK, ACOLS = (A_KFAST ? PK : PM) + PPAD;
    constexpr int BROWS = B_KFAST ? PN : PK, BCOLS = (B_KFAST ? PK : PN) + PPAD;
    __shared__ __align__(16) float As[PST][AROWS][ACOLS];
    __shared__ __align__(16) float Bs[PST][BROWS][BCOLS];
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const int m0 = blockIdx.y * PM, n0 = blockIdx.x * PN;
    const int kbeg = sec.mode == 2 ? 0 : blockIdx.z * kchunk;
    const int kend = min(K, kbeg + kchunk);
    const int ntile = kend > kbeg ? (kend - kbeg + PK - 1) / PK : 0;

    auto issue = [&](int t) {                                  // k-tile t of this CTA -> stage t % PST
        const int st = t % PST, k0 = kbeg + t * PK;
        {   // one 16-byte chunk of A per thread
            int r, c, gm, gk, valid;
            if (A_KFAST) { r = tid / 4; c = (tid % 4) * 4; gm = m0 + r; gk = k0 + c; valid = gm < M ? (kend - gk) : 0; }
            else { r = tid / 16; c = (tid % 16) * 4; gk = k0 + r; gm = m0 + c; valid = gk < kend ? (M - gm) : 0; }
            valid = max(0, min(4, valid));
            const float* src = valid > 0 ? (A_KFAST ? A + (int64_t)gm * lda + gk : A + (int64_t)gk * lda + gm) : A;
            cp_async16(&As[st][r][c], src, valid * 4);
        }
        {
            int r, c, gn, gk, valid;
            if (B_KFAST) { r = tid / 4; c = (tid % 4) * 4; gn = n0 + r; gk = k0 + c; valid = gn < N ? (kend - gk) : 0; }
            else { r = tid / 16; c = (tid % 16) * 4; gk = k0 + r; gn = n0 + c; valid = gk < kend ? (N - gn) : 0; }
            valid = max(0, min(4, valid));
            const float* src = valid > 0 ? (B_KFAST ? Bm + (int64_t)gn * ldb + gk : Bm + (int64_t)gk * ldb + gn) : Bm;
            cp_async16(&Bs[st][r][c], src, valid * 4);
        }
    };

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int pass = 0; pass < npass; ++pass) {
    if (pass == 1) {
        cp_async_wait<0>();
        __syncthreads();                                       // every stage of the first pass consumed
        A = sec.A; Bm = sec.B;
    }
#pragma unroll
    for (int t = 0; t < PST - 1; ++t) {
        if (t < ntile) issue(t);
        cp_async_commit();
    }
    for (int t = 0; t < ntile; ++t) {
        cp_async_wait<PST - 2>();
        __syncthreads();                                       // tile t landed; everyone is done with tile t-1's stage
        if (t + PST - 1 < ntile) issue(t + PST - 1);
        cp_async_commit();
        const int st = t % PST;
#pragma unroll
        for (int kq = 0; kq < PK / 4; ++kq) {
            float a[4][4], b[4][4];                            // [row i | col j][k within the group of 4]
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                if (A_KFAST) {
                    const float4 v = ld4(&As[st][ty * 4 + x][kq * 4]);
                    a[x][0] = v.x; a[x][1] = v.y; a[x][2] = v.z; a[x][3] = v.w;
                } else {
                    const float4 v = ld4(&As[st][kq * 4 + x][ty * 4]);
                    a[0][x] = v.x; a[1][x] = v.y; a[2][x] = v.z; a[3][x] = v.w;
                }
                if (B_KFAST) {                                 // columns tx + 16 j: conflict-free with the 20-float row stride
                    const float4 v = ld4(&Bs[st][tx + 16 * x][kq * 4]);
                    b[x][0] = v.x; b[x][1] = v.y; b[x][2] = v.z; b[x][3] = v.w;
                } else {                                       // columns 4 tx + j
                    const float4 v = ld4(&Bs[st][kq * 4 + x][tx * 4]);
                    b[0][x] = v.x; b[1][x] = v.y; b[2][x] = v.z; b[3][x] = v.w;
                }
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i][kk], b[j][kk], acc[i][j]);
        }
    }
    }   // pass
    const bool split = gridDim.z > 1 && sec.mode != 2;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + (B_KFAST ? tx + 16 * j : tx * 4 + j);
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

#endif  // CPG_EMU

__global__ void k_splitk_reduce(int M, int N, int splits, float alpha, const float* __restrict__ ws, float beta,
                                float* __restrict__ C, int64_t ldc, const float* __restrict__ bias) {
    const int i = blockIdx.x * RED_X + threadIdx.x;
    const bool ok = i < M * N;
    const float s = block_split_sum(ws, (size_t)M * N, splits, (size_t)(ok ? i : 0), ok);
    if (!ok || threadIdx.y != 0) return;
    int m = i / N, n = i % N;
    float v = alpha * s;
    if (bias != nullptr) v += bias[n];
    if (beta != 0.f) v += beta * C[m * ldc + n];
    C[m * ldc + n] = v;
}

static void launch_sgemm_impl(cudaStream_t s, int M, int N, int K, float alpha, const float* A, int64_t sam, int64_t sak,
                              const float* B, int64_t sbk, int64_t sbn, float beta, float* C, int64_t ldc,
                              const float* bias, int split_k, float* ws, GemmSecond sec);

void launch_sgemm(cudaStream_t s, int M, int N, int K, float alpha, const float* A, int64_t sam, int64_t sak,
                  const float* B, int64_t sbk, int64_t sbn, float beta, float* C, int64_t ldc,
                  const float* bias, int split_k, float* ws) {
    GemmSecond none{nullptr, nullptr, nullptr, nullptr, 0};
    launch_sgemm_impl(s, M, N, K, alpha, A, sam, sak, B, sbk, sbn, beta, C, ldc, bias, split_k, ws, none);
}
// C = alpha (A B + A2 B2) + beta C (+ bias): two products of equal shape and strides in one launch
void launch_sgemm_sum2(cudaStream_t s, int M, int N, int K, float alpha, const float* A, const float* A2, int64_t sam, int64_t sak,
                       const float* B, const float* B2, int64_t sbk, int64_t sbn, float beta, float* C, int64_t ldc) {
    GemmSecond sec{A2, B2, nullptr, nullptr, 1};
    launch_sgemm_impl(s, M, N, K, alpha, A, sam, sak, B, sbk, sbn, beta, C, ldc, nullptr, 1, nullptr, sec);
}
// C = alpha A B + bias and C2 = alpha A B2 + bias2 (shared A, equal shapes and strides) in one launch
void launch_sgemm_pair(cudaStream_t s, int M, int N, int K, float alpha, const float* A, int64_t sam, int64_t sak,
                       const float* B, const float* B2, int64_t sbk, int64_t sbn, float* C, float* C2, int64_t ldc,
                       const float* bias, const float* bias2) {
    GemmSecond sec{A, B2, bias2, C2, 2};
    launch_sgemm_impl(s, M, N, K, alpha, A, sam, sak, B, sbk, sbn, 0.f, C, ldc, bias, 1, nullptr, sec);
}

static void launch_sgemm_impl(cudaStream_t s, int M, int N, int K, float alpha, const float* A, int64_t sam, int64_t sak,
                              const float* B, int64_t sbk, int64_t sbn, float beta, float* C, int64_t ldc,
                              const float* bias, int split_k, float* ws, GemmSecond sec) {
    if (split_k < 1 || ws == nullptr) split_k = 1;
    int kchunk = ceil_div(ceil_div(K, split_k), GK) * GK;
    split_k = ceil_div(K, kchunk);
    // pipelined kernel: one operand axis has unit stride, the other a multiple of 4 floats, bases 16-byte aligned
    const bool a_k = sak == 1, a_m = sam == 1, b_k = sbk == 1, b_n = sbn == 1;
    const int64_t lda = a_k ? sam : sak, ldb = b_k ? sbn : sbk;
    const bool aligned = (a_k || a_m) && (b_k || b_n) && lda % 4 == 0 && ldb % 4 == 0 &&
                         ((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0 &&
                         (sec.mode == 0 || (((uintptr_t)sec.A & 15) == 0 && ((uintptr_t)sec.B & 15) == 0));
#ifdef CPG_EMU
    const bool use_pipe = false;
#else
    const bool use_pipe = aligned;
#endif
    if (!use_pipe && sec.mode != 0) {
        // unaligned operands: two plain launches with the same meaning
        GemmSecond none{nullptr, nullptr, nullptr, nullptr, 0};
        launch_sgemm_impl(s, M, N, K, alpha, A, sam, sak, B, sbk, sbn, beta, C, ldc, bias, 1, nullptr, none);
        if (sec.mode == 1) launch_sgemm_impl(s, M, N, K, alpha, sec.A, sam, sak, sec.B, sbk, sbn, 1.f, C, ldc, nullptr, 1, nullptr, none);
        else launch_sgemm_impl(s, M, N, K, alpha, sec.A, sam, sak, sec.B, sbk, sbn, 0.f, sec.C, ldc, sec.bias, 1, nullptr, none);
        return;
    }
#ifndef CPG_EMU
    if (use_pipe) {
        dim3 grid(ceil_div(N, PN), ceil_div(M, PM), sec.mode == 2 ? 2 : split_k);
        const char* lbl = g_profile_on ? shape_label("k_sgemm", M, N, K, sec.mode ? 2 : 1) : "k_sgemm";
#define CPG_PIPE(AK, BK) CPG_LAUNCH_NAMED(lbl, (k_sgemm_pipe<AK, BK>), grid, 256, 0, s, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, kchunk, ws, sec)
        if (a_k && !b_k) CPG_PIPE(true, false);
        else if (a_k && b_k) CPG_PIPE(true, true);
        else if (!a_k && !b_k) CPG_PIPE(false, false);
        else CPG_PIPE(false, true);
#undef CPG_PIPE
        if (split_k > 1)
            CPG_LAUNCH(k_splitk_reduce, CPG_RED_GRID(M * N), CPG_RED_BLOCK, 0, s, M, N, split_k, alpha, ws, beta, C, ldc, bias);
        return;
    }
#endif
    dim3 grid(ceil_div(N, GN), ceil_div(M, GM), split_k);
    CPG_LAUNCH(k_sgemm, grid, 256, 0, s, M, N, K, alpha, A, sam, sak, B, sbk, sbn, beta, C, ldc, bias, kchunk, ws);
    if (split_k > 1)
        CPG_LAUNCH(k_splitk_reduce, CPG_RED_GRID(M * N), CPG_RED_BLOCK, 0, s, M, N, split_k, alpha, ws, beta, C, ldc, bias);
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
    const int n = blockIdx.x * RED_X + threadIdx.x;
    const bool ok = n < N;
    const float s = block_split_sum(part, (size_t)N, nchunk, (size_t)(ok ? n : 0), ok);
    if (ok && threadIdx.y == 0) out[n] = s;
}

void launch_colsum(cudaStream_t s, const float* A, int M, int N, int64_t lda, float* out, float* ws, int nchunk) {
    nchunk = max(1, min(nchunk, M));
    int rpc = ceil_div(M, nchunk);
    nchunk = ceil_div(M, rpc);
    CPG_LAUNCH(k_colsum_partial, dim3(ceil_div(N, 128), nchunk), 128, 0, s, A, M, N, lda, rpc, ws);
    CPG_LAUNCH(k_colsum_final, CPG_RED_GRID(N), CPG_RED_BLOCK, 0, s, ws, nchunk, N, out);
}

}  // namespace cpg
