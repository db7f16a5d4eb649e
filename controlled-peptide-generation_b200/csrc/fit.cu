// Fitting of the latent density Q(z) and of the z-space attribute classifiers on the device (the reference fits both
// with scikit-learn on the host: density_modeling.py:64-73 GaussianMixture(covariance_type='diag').fit, 26 s for
// 50,000 x 100 points at K = 100; sample_pipeline.py:169-192 LogisticRegression(lbfgs, 200)).
//
//   cpg_gmm_em_step   one EM iteration of a diagonal-covariance mixture, fp64, the arithmetic of sklearn
//                     mixture/_gaussian_mixture.py (_estimate_log_gaussian_prob 'diag' + logsumexp responsibilities;
//                     _estimate_gaussian_parameters: nk + 10 eps, means, avg_X2 - 2 avg_X_means + avg_means2 + reg_covar;
//                     weights nk / N renormalised).  E-step: one warp per point (lanes over components); M-step: one CTA per
//                     (component, chunk of points) with ordered partials -- bit-reproducible, no atomics.
//   cpg_logreg_newton_stats  loss, gradient and Hessian of the L2-penalised logistic loss sklearn minimises
//                     (C = 1: sum_i log(1 + exp(-y_i s_i)) + 1/2 |w|^2, intercept unpenalised); the 101 x 101 Newton solve is
//                     host work (cpg_b200/fit.py).  The optimum is unique, so Newton and sklearn's L-BFGS meet there.
#include "ctx.h"

namespace cpg {
int check_launch(const char* where);

constexpr int EM_WARPS = 8;
// resp[n][k] (fp64) and the per-point log-likelihood; x fp32 [N][D], parameters fp64 [K][D]
__global__ void __launch_bounds__(EM_WARPS * 32)
k_gmm_estep(const float* __restrict__ x, int64_t N, const double* __restrict__ mean, const double* __restrict__ prec,
            const double* __restrict__ logw_norm, int K, double* __restrict__ resp, double* __restrict__ loglik) {
    __shared__ double xs[EM_WARPS][ZD];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t i = (int64_t)blockIdx.x * EM_WARPS + warp; i < N; i += (int64_t)gridDim.x * EM_WARPS) {
        for (int d = lane; d < ZD; d += 32) xs[warp][d] = (double)x[i * ZD + d];
        __syncwarp();
        double m = -INFINITY;
        for (int k = lane; k < K; k += 32) {                 // pass 1: weighted log-probabilities, running maximum
            double q = 0.0;
            const double* mu = mean + (size_t)k * ZD;
            const double* pr = prec + (size_t)k * ZD;
            for (int d = 0; d < ZD; ++d) { const double df = xs[warp][d] - mu[d]; q = fma(df * df, pr[d], q); }
            const double lp = logw_norm[k] - 0.5 * q;
            resp[i * K + k] = lp;
            m = fmax(m, lp);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        double ssum = 0.0;
        for (int k = lane; k < K; k += 32) ssum += exp(resp[i * K + k] - m);
        ssum = warp_sum(ssum);
        const double lse = m + log(ssum);
        for (int k = lane; k < K; k += 32) resp[i * K + k] = exp(resp[i * K + k] - lse);
        if (lane == 0) loglik[i] = lse;
        __syncwarp();
    }
}

// partial sufficient statistics of component k over the points of chunk c: [nk | sum r x (D) | sum r x^2 (D)]
constexpr int EM_MT = 128;
__global__ void __launch_bounds__(EM_MT)
k_gmm_mstep_partial(const float* __restrict__ x, const double* __restrict__ resp, int64_t N, int K, int64_t rows_per_chunk,
                    double* __restrict__ part) {
    const int k = blockIdx.x, c = blockIdx.y, d = threadIdx.x;
    const int64_t n0 = (int64_t)c * rows_per_chunk, n1 = min(N, n0 + rows_per_chunk);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int64_t n = n0; n < n1; ++n) {
        const double r = resp[n * K + k];
        if (d < ZD) { const double v = (double)x[n * ZD + d]; s1 = fma(r, v, s1); s2 = fma(r * v, v, s2); }
        else if (d == ZD) s0 += r;
    }
    double* out = part + ((size_t)c * K + k) * (2 * ZD + 1);
    if (d < ZD) { out[1 + d] = s1; out[1 + ZD + d] = s2; }
    else if (d == ZD) out[0] = s0;
}
__global__ void __launch_bounds__(EM_MT)
k_gmm_mstep_final(const double* __restrict__ part, int nchunk, int K, int64_t N, double reg_covar, double* __restrict__ weights,
                  double* __restrict__ mean, double* __restrict__ cov) {
    const int k = blockIdx.x, d = threadIdx.x;
    __shared__ double nk_s;
    if (d == ZD) {
        double s = 0.0;
        for (int c = 0; c < nchunk; ++c) s += part[((size_t)c * K + k) * (2 * ZD + 1)];
        nk_s = s + 10.0 * 2.220446049250313e-16;               // sklearn: nk = resp.sum(0) + 10 * eps
        weights[k] = nk_s / (double)N;                          // renormalised by the host (sum over k)
    }
    __syncthreads();
    if (d < ZD) {
        double s1 = 0.0, s2 = 0.0;
        for (int c = 0; c < nchunk; ++c) {
            const double* p = part + ((size_t)c * K + k) * (2 * ZD + 1);
            s1 += p[1 + d]; s2 += p[1 + ZD + d];
        }
        const double mu = s1 / nk_s;
        const double avg_x2 = s2 / nk_s, avg_means2 = mu * mu, avg_x_means = mu * s1 / nk_s;
        mean[(size_t)k * ZD + d] = mu;
        cov[(size_t)k * ZD + d] = avg_x2 - 2.0 * avg_x_means + avg_means2 + reg_covar;
    }
}

// ---- logistic regression: per-CTA partials of [loss | grad (D+1) | Hessian upper triangle ((D+1)(D+2)/2)]
constexpr int LR_D1 = ZD + 1;
constexpr int LR_H = LR_D1 * (LR_D1 + 1) / 2;
constexpr int LR_T = 256;
__global__ void __launch_bounds__(LR_T)
k_logreg_stats(const float* __restrict__ x, const float* __restrict__ y, int64_t N, const double* __restrict__ w, int64_t rows_per_cta,
               double* __restrict__ part) {
    __shared__ double xs[LR_D1];
    __shared__ double ws[LR_D1];
    __shared__ double sc[2];
    __shared__ uint8_t hi_s[LR_H], hj_s[LR_H];                 // Hessian entry e -> (i, j), i <= j, row-major upper triangle
    const int t = threadIdx.x;
    for (int i = t; i < LR_D1; i += LR_T) {
        ws[i] = w[i];
        const int base = i * LR_D1 - i * (i - 1) / 2;
        for (int j = i; j < LR_D1; ++j) { hi_s[base + j - i] = (uint8_t)i; hj_s[base + j - i] = (uint8_t)j; }
    }
    // every thread owns a fixed set of Hessian entries (row-major upper triangle) and gradient entries
    double hacc[(LR_H + LR_T - 1) / LR_T];
    double gacc = 0.0, lacc = 0.0;
#pragma unroll
    for (int q = 0; q < (LR_H + LR_T - 1) / LR_T; ++q) hacc[q] = 0.0;
    const int64_t n0 = (int64_t)blockIdx.x * rows_per_cta, n1 = min(N, n0 + rows_per_cta);
    for (int64_t n = n0; n < n1; ++n) {
        __syncthreads();
        for (int i = t; i < LR_D1; i += LR_T) xs[i] = i < ZD ? (double)x[n * ZD + i] : 1.0;
        __syncthreads();
        if (t == 0) {
            double s = 0.0;
            for (int i = 0; i < LR_D1; ++i) s = fma(xs[i], ws[i], s);
            const double yy = (double)y[n];                                  // labels 0 / 1
            const double p = 1.0 / (1.0 + exp(-s));
            sc[0] = p - yy;                                                    // d loss / d s
            sc[1] = p * (1.0 - p);
            lacc += (s > 0 ? s : 0.0) + log1p(exp(-fabs(s))) - yy * s;         // log(1 + e^s) - y s, stable
        }
        __syncthreads();
        if (t < LR_D1) gacc = fma(sc[0], xs[t], gacc);
#pragma unroll
        for (int q = 0; q < (LR_H + LR_T - 1) / LR_T; ++q) {
            const int e = t + q * LR_T;
            if (e < LR_H) {
                const int i = hi_s[e], j = hj_s[e];
                hacc[q] = fma(sc[1] * xs[i], xs[j], hacc[q]);
            }
        }
    }
    double* out = part + (size_t)blockIdx.x * (1 + LR_D1 + LR_H);
    if (t == 0) out[0] = lacc;
    if (t < LR_D1) out[1 + t] = gacc;
#pragma unroll
    for (int q = 0; q < (LR_H + LR_T - 1) / LR_T; ++q) { const int e = t + q * LR_T; if (e < LR_H) out[1 + LR_D1 + e] = hacc[q]; }
}
__global__ void k_sum_partials_f64(const double* __restrict__ part, int nparts, int len, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += part[(size_t)p * len + i];
    out[i] = s;
}

}  // namespace cpg

using namespace cpg;

extern "C" {

int cpg_gmm_em_step(cpg_ctx* ctx, cpg_stream stream, const float* x, int64_t N, int K, const double* mean_in, const double* prec_in,
                    const double* logw_norm_in, double reg_covar, double* resp_ws, double* weights_out, double* mean_out,
                    double* cov_out, double* loglik_out) {
    if (!ctx || !x || !mean_in || !prec_in || !logw_norm_in || !resp_ws || !weights_out || !mean_out || !cov_out || !loglik_out || N < 1 || K < 1) {
        set_error("cpg_gmm_em_step: bad argument"); return CPG_EINVAL;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int nchunk = (int)std::max<int64_t>(1, std::min<int64_t>((N + 255) / 256, 64));
    const int64_t rpc = (N + nchunk - 1) / nchunk;
    int rc = ensure_aux(ctx, (size_t)nchunk * K * (2 * ZD + 1) * sizeof(double), s);
    if (rc) return rc;
    double* part = (double*)ctx->aux;
    int grid = (int)std::min<int64_t>((N + EM_WARPS - 1) / EM_WARPS, (int64_t)ctx->sm_count * 8);
    CPG_LAUNCH(k_gmm_estep, grid, EM_WARPS * 32, 0, s, x, N, mean_in, prec_in, logw_norm_in, K, resp_ws, loglik_out);
    CPG_LAUNCH(k_gmm_mstep_partial, dim3(K, nchunk), EM_MT, 0, s, x, resp_ws, N, K, rpc, part);
    CPG_LAUNCH(k_gmm_mstep_final, K, EM_MT, 0, s, part, nchunk, K, N, reg_covar, weights_out, mean_out, cov_out);
    return check_launch("cpg_gmm_em_step");
}

int cpg_logreg_newton_stats(cpg_ctx* ctx, cpg_stream stream, const float* x, const float* y01, int64_t N, const double* w101,
                            double* out) {
    if (!ctx || !x || !y01 || !w101 || !out || N < 1) { set_error("cpg_logreg_newton_stats: bad argument"); return CPG_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    const int len = 1 + LR_D1 + LR_H;
    const int nparts = (int)std::max<int64_t>(1, std::min<int64_t>((N + 63) / 64, (int64_t)ctx->sm_count * 2));
    const int64_t rpc = (N + nparts - 1) / nparts;
    int rc = ensure_aux(ctx, (size_t)nparts * len * sizeof(double), s);
    if (rc) return rc;
    double* part = (double*)ctx->aux;
    CPG_LAUNCH(k_logreg_stats, nparts, LR_T, 0, s, x, y01, N, w101, rpc, part);
    CPG_LAUNCH(k_sum_partials_f64, ceil_div(len, 256), 256, 0, s, part, nparts, len, out);
    return check_launch("cpg_logreg_newton_stats");
}

int cpg_logreg_stats_len(void) { return 1 + LR_D1 + LR_H; }

}  // extern "C"
