// CLaSS latent-space sampling: draw z ~ Q (diag Gaussian mixture), score with the z-space
// attribute classifiers, rejection-accept; plus batched mixture / prior log-densities.
//
// Replaces, per draw, the CPU numpy / scikit-learn calls of density_modeling.py:50-60
// (`rejection_sample`): GaussianMixture.sample (sklearn mixture/_base.py:434-511; 83 % of the
// reference's time), LogisticRegression.predict_proba (density_modeling.py:43-48), the running
// product of target-class probabilities and `np.random.uniform(n) < prod`.  Log densities replace
// the per-point Python loop of evaluate_nll (density_modeling.py:99-108): mogQ.logpdf (:75-77,
// sklearn _gaussian_mixture.py:536-542 + logsumexp) and prior_logpdf (:11-14).
//
// Score arithmetic follows the dtype of the fitted classifier (see oracle/class_sampling.py):
// float32 when the classifier was fitted on float32 z's (the reference pipeline), float64 otherwise.
#include "ctx.h"

namespace cpg {
int check_launch(const char* where);

constexpr int MAX_CLF = 4;
struct ClfSpec {
    const double* coef[MAX_CLF];   // [D] each (device, stored as fp64; rounded to fp32 when f32 != 0)
    double intercept[MAX_CLF];
    int target_col[MAX_CLF];
    int f32[MAX_CLF];
    int n_clf;
};

__device__ __forceinline__ float expit_f(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ double expit_d(double x) { return 1.0 / (1.0 + exp(-x)); }

// ---- parity mode: z and u are inputs (the reference's own draws) ------------------------------
// one thread per draw; coefficients in shared memory
__global__ void __launch_bounds__(128)
k_score_accept(const float* __restrict__ z, const double* __restrict__ u, int64_t n, ClfSpec cs,
               double* __restrict__ probs, double* __restrict__ accum_out, uint8_t* __restrict__ accept) {
    __shared__ double coef_s[MAX_CLF][ZD];
    for (int i = threadIdx.x; i < cs.n_clf * ZD; i += blockDim.x) coef_s[i / ZD][i % ZD] = cs.coef[i / ZD][i % ZD];
    __syncthreads();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* zi = z + i * ZD;
    double acc_d = 1.0;
    float acc_f = 1.0f;
    bool all_f32 = true;
    for (int a = 0; a < cs.n_clf; ++a) {
        double p;
        if (cs.f32[a]) {
            float s = 0.f;
            for (int d = 0; d < ZD; ++d) s = fmaf(zi[d], (float)coef_s[a][d], s);
            s += (float)cs.intercept[a];
            float p1 = expit_f(s);
            float pf = cs.target_col[a] == 1 ? p1 : 1.0f - p1;
            acc_f *= pf;
            acc_d *= (double)pf;
            p = (double)pf;
        } else {
            all_f32 = false;
            double s = 0.0;
            for (int d = 0; d < ZD; ++d) s = fma((double)zi[d], coef_s[a][d], s);
            s += cs.intercept[a];
            double p1 = expit_d(s);
            p = cs.target_col[a] == 1 ? p1 : 1.0 - p1;
            acc_d *= p;
        }
        if (probs != nullptr) probs[(size_t)a * n + i] = p;
    }
    double accum = all_f32 ? (double)acc_f : acc_d;     // numpy keeps float32 products in float32
    if (accum_out != nullptr) accum_out[i] = accum;
    accept[i] = u[i] < accum ? 1 : 0;
}

// ---- perf mode: everything drawn in-kernel (Philox), ONE LANE PER DRAW ---------------------------------------
// A thread owns a draw from the component choice to the accept test: 1 + 25 Philox4x32-10 calls (stream 0: component
// and acceptance uniforms; stream 1 + q: the four normals of dimensions 4q .. 4q+3), Box-Muller, z = mean_k + sd_k n,
// the classifier dots accumulated in registers.  No shuffles, no warp reductions, no idle lanes (the first version
// spread one draw over a warp: 32 Philox calls for 100 normals and a reduction tree per classifier -- 3x the
// instructions).  mean / sd of all components sit in shared memory as interleaved (mean, sd) pairs with an odd row
// stride (lanes read different components: rows land in different banks); classifier coefficients are broadcast reads.
// z rows leave through a per-warp staging tile in shared memory, 20 dimensions at a time: the warp then writes 80-byte
// row pieces with consecutive lanes on consecutive 16 bytes (lane-per-row 16-byte stores 400 bytes apart cost twice the
// L2 transactions: measured 2.8 -> see DESIGN.md); scores and the mask are lane-contiguous.
struct GmmSpec {
    const float* mean;      // [K][D] fp32
    const float* sd;        // [K][D] sqrt(cov), fp32
    const float* cdf;       // [K] cumulative weights (last = 1)
    int K;
};
constexpr int CS_THREADS = 256;
constexpr int CS_STRIDE = ZD + 1;                 // float2 per row, odd
constexpr int CS_KMAX_SMEM = 128;                 // components whose tables fit in shared memory
constexpr int CS_ZCH = 20;                        // dimensions per staged piece of a z row (80 bytes)
template <bool TAB_SMEM>
__global__ void __launch_bounds__(CS_THREADS)
k_class_sample(GmmSpec g, ClfSpec cs, uint64_t seed, int64_t offset, const int64_t* __restrict__ gid_list, int64_t n,
               float* __restrict__ z_out, double* __restrict__ probs, double* __restrict__ accum_out,
               uint8_t* __restrict__ accept, int* __restrict__ comp_out, unsigned long long* __restrict__ n_accepted) {
    CPG_DYN_SMEM(float, dsm);
    float2* tab_s = reinterpret_cast<float2*>(dsm);                       // [K][CS_STRIDE] (mean, sd)   (TAB_SMEM)
    float* coef_s = dsm + (TAB_SMEM ? ((2 * (size_t)g.K * CS_STRIDE + 3) & ~(size_t)3) : 0);  // [MAX_CLF][ZD]
    float* cdf_s = coef_s + MAX_CLF * ZD;                                 // [K]
    float* stage = cdf_s + ((g.K + 3) & ~3) + (threadIdx.x >> 5) * (32 * CS_ZCH);   // this warp's [32 draws][CS_ZCH] tile
    const int lane = threadIdx.x & 31;
    __shared__ unsigned long long acc_s;
    if (TAB_SMEM)
        for (int i = threadIdx.x; i < g.K * ZD; i += blockDim.x)
            tab_s[(i / ZD) * CS_STRIDE + (i % ZD)] = make_float2(g.mean[i], g.sd[i]);
    for (int i = threadIdx.x; i < MAX_CLF * ZD; i += blockDim.x) coef_s[i] = i < cs.n_clf * ZD ? (float)cs.coef[i / ZD][i % ZD] : 0.f;
    for (int i = threadIdx.x; i < g.K; i += blockDim.x) cdf_s[i] = g.cdf[i];
    if (threadIdx.x == 0) acc_s = 0ull;
    __syncthreads();
    unsigned long long local_acc = 0;
    // whole warps iterate together (the staged z write-out is warp-collective); lanes past n idle through the math
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < n; i0 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = i0 + lane;
        const bool live = i < n;
        // draw number: consecutive from `offset`, or (re-generation of selected draws) taken from a list
        const uint64_t gid = !live ? 0ull : (gid_list != nullptr ? (uint64_t)gid_list[i] : (uint64_t)(offset + i));
        uint32_t r[4];
        Philox::gen(seed, gid, 0u, r);
        const float uc = (float)(r[0] >> 8) * (1.0f / 16777216.0f);
        const double ua = u64_to_unit(r[2], r[3]);
        int lo = 0, hi = g.K - 1;                                          // first k with cdf[k] > uc
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (cdf_s[mid] > uc) hi = mid; else lo = mid + 1; }
        const int k = lo;
        float dots[MAX_CLF] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 5
        for (int q = 0; q < ZD / 4; ++q) {
            Philox::gen(seed, gid, 1u + (uint32_t)q, r);
            float nrm[4];
            {
                float u1 = u32_to_unit_open(r[0]), u2 = u32_to_unit_open(r[1]);
                float rad = sqrtf(-2.0f * __logf(u1)); float sn, cn; __sincosf(6.283185307179586f * u2, &sn, &cn);
                nrm[0] = rad * cn; nrm[1] = rad * sn;
                u1 = u32_to_unit_open(r[2]); u2 = u32_to_unit_open(r[3]);
                rad = sqrtf(-2.0f * __logf(u1)); __sincosf(6.283185307179586f * u2, &sn, &cn);
                nrm[2] = rad * cn; nrm[3] = rad * sn;
            }
            float zv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int d = 4 * q + e;
                float2 ms;
                if (TAB_SMEM) ms = tab_s[k * CS_STRIDE + d];
                else ms = make_float2(__ldg(g.mean + (size_t)k * ZD + d), __ldg(g.sd + (size_t)k * ZD + d));
                zv[e] = fmaf(nrm[e], ms.y, ms.x);
#pragma unroll
                for (int a = 0; a < MAX_CLF; ++a) dots[a] = fmaf(zv[e], coef_s[a * ZD + d], dots[a]);
            }
            if (z_out != nullptr) {
                constexpr int QPC = CS_ZCH / 4;                               // quads per staged piece
                *reinterpret_cast<float4*>(stage + lane * CS_ZCH + (q % QPC) * 4) = make_float4(zv[0], zv[1], zv[2], zv[3]);
                if (q % QPC == QPC - 1) {
                    __syncwarp();
                    const int d0 = (q / QPC) * CS_ZCH;
#pragma unroll
                    for (int j = 0; j < QPC; ++j) {                           // 32 * QPC float4 of the tile, lane-linear
                        const int idx = j * 32 + lane, row = idx / QPC, pc = idx % QPC;
                        if (i0 + row < n)
                            *reinterpret_cast<float4*>(z_out + (size_t)(i0 + row) * ZD + d0 + pc * 4) =
                                *reinterpret_cast<const float4*>(stage + row * CS_ZCH + pc * 4);
                    }
                    __syncwarp();
                }
            }
        }
        if (!live) continue;
        float accf = 1.0f;
#pragma unroll
        for (int a = 0; a < MAX_CLF; ++a) {
            if (a < cs.n_clf) {
                const float sc = dots[a] + (float)cs.intercept[a];
                const float p1 = expit_f(sc);
                const float p = cs.target_col[a] == 1 ? p1 : 1.0f - p1;
                accf *= p;
                if (probs != nullptr) probs[(size_t)a * n + i] = (double)p;
            }
        }
        const int acc = ua < (double)accf ? 1 : 0;
        if (accept != nullptr) accept[i] = (uint8_t)acc;
        if (accum_out != nullptr) accum_out[i] = (double)accf;
        if (comp_out != nullptr) comp_out[i] = k;
        local_acc += acc;
    }
    if (n_accepted != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) local_acc += __shfl_xor_sync(0xffffffffu, local_acc, o);
        if ((threadIdx.x & 31) == 0 && local_acc) atomicAdd(&acc_s, local_acc);
        __syncthreads();
        if (threadIdx.x == 0 && acc_s) atomicAdd(n_accepted, acc_s);
    }
}

static int launch_class_sample(cpg_ctx* ctx, cudaStream_t s, const char* label, const GmmSpec& g, const ClfSpec& cs, uint64_t seed,
                               int64_t offset, const int64_t* gid_list, int64_t n, float* z_out, double* probs, double* accum,
                               uint8_t* accept, int* comp_out, unsigned long long* n_accepted) {
    const bool tab = g.K <= CS_KMAX_SMEM;
    const size_t smem = ((tab ? ((2 * (size_t)g.K * CS_STRIDE + 3) & ~(size_t)3) : 0) + MAX_CLF * ZD + ((g.K + 3) & ~3) + (CS_THREADS / 32) * 32 * CS_ZCH) * sizeof(float);
    const int64_t want = (n + CS_THREADS - 1) / CS_THREADS;
    const int per_sm = tab ? 2 : 4;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)ctx->sm_count * per_sm));
    if (tab) {
        auto kfn = k_class_sample<true>;
        CPG_SET_MAX_SMEM(kfn, smem);
        CPG_LAUNCH_NAMED(label, kfn, grid, CS_THREADS, smem, s, g, cs, seed, offset, gid_list, n, z_out, probs, accum, accept, comp_out, n_accepted);
    } else {
        auto kfn = k_class_sample<false>;
        CPG_LAUNCH_NAMED(label, kfn, grid, CS_THREADS, smem, s, g, cs, seed, offset, gid_list, n, z_out, probs, accum, accept, comp_out, n_accepted);
    }
    return CPG_OK;
}

// ---- log densities (fp64) -----------------------------------------------------------------------
// log sum_k w_k N(x; mu_k, diag(cov_k)); lanes over components, LP_PTS points per warp: every (mean, precision) pair a
// lane loads is used for LP_PTS points (with one point per warp the kernel ran at 54 B/clk/SM of L1 loads, 8 % of the
// fp64 rate).  Per point the arithmetic and its order are unchanged: q over d in order, online log-sum-exp over the
// lane's components, then the shuffle tree.
constexpr int LP_WARPS = 8;
constexpr int LPG_WARPS = 4;                 // k_gmm_logpdf: 4 warps x 8 points x 100 doubles of shared memory per block
constexpr int LP_PTS = 8;
__global__ void __launch_bounds__(LPG_WARPS * 32)
k_gmm_logpdf(const float* __restrict__ x, int64_t n, const double* __restrict__ mean_t, const double* __restrict__ prec_t,
             const double* __restrict__ logw_norm, int K, double* __restrict__ out) {
    // mean_t / prec_t are [D][K] (component-fastest) so that lanes read consecutive addresses
    __shared__ __align__(16) double xs[LPG_WARPS][ZD][LP_PTS];          // point-fastest: 16-byte broadcast loads
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t i0 = ((int64_t)blockIdx.x * LPG_WARPS + warp) * LP_PTS; i0 < n; i0 += (int64_t)gridDim.x * LPG_WARPS * LP_PTS) {
        for (int e = lane; e < ZD * LP_PTS; e += 32) {
            const int p = e / ZD, d = e % ZD;                          // consecutive lanes: consecutive floats of a point
            xs[warp][d][p] = i0 + p < n ? (double)x[(i0 + p) * ZD + d] : 0.0;
        }
        __syncwarp();
        double m[LP_PTS], ssum[LP_PTS];                                 // running logsumexp over this lane's components
#pragma unroll
        for (int p = 0; p < LP_PTS; ++p) { m[p] = -INFINITY; ssum[p] = 0.0; }
        for (int k = lane; k < K; k += 32) {
            double q[LP_PTS];
#pragma unroll
            for (int p = 0; p < LP_PTS; ++p) q[p] = 0.0;
            for (int d = 0; d < ZD; ++d) {
                const double mu = mean_t[(size_t)d * K + k], pr = prec_t[(size_t)d * K + k];
#pragma unroll
                for (int p2 = 0; p2 < LP_PTS; p2 += 2) {
                    const double2 xv = *reinterpret_cast<const double2*>(&xs[warp][d][p2]);
                    const double d0 = xv.x - mu, d1 = xv.y - mu;
                    q[p2] = fma(d0 * d0, pr, q[p2]);
                    q[p2 + 1] = fma(d1 * d1, pr, q[p2 + 1]);
                }
            }
            const double lw = logw_norm[k];
#pragma unroll
            for (int p = 0; p < LP_PTS; ++p) {
                const double lp = lw - 0.5 * q[p];
                if (lp > m[p]) { ssum[p] = ssum[p] * exp(m[p] - lp) + 1.0; m[p] = lp; } else { ssum[p] += exp(lp - m[p]); }
            }
        }
#pragma unroll
        for (int p = 0; p < LP_PTS; ++p) {
            double mp = m[p], sp = ssum[p];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double om = __shfl_xor_sync(0xffffffffu, mp, o);
                const double os = __shfl_xor_sync(0xffffffffu, sp, o);
                const double nm = fmax(mp, om);
                if (nm == -INFINITY) { sp = 0.0; } else { sp = sp * exp(mp - nm) + os * exp(om - nm); }
                mp = nm;
            }
            if (lane == 0 && i0 + p < n) out[i0 + p] = mp + log(sp);
        }
        __syncwarp();
    }
}

// -D/2 log(2 pi) - |z|^2 / 2
__global__ void __launch_bounds__(LP_WARPS * 32)
k_prior_logpdf(const float* __restrict__ x, int64_t n, double* __restrict__ out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t i = (int64_t)blockIdx.x * LP_WARPS + warp; i < n; i += (int64_t)gridDim.x * LP_WARPS) {
        double s = 0.0;
        for (int d = lane; d < ZD; d += 32) { double v = (double)x[i * ZD + d]; s = fma(v, v, s); }
        s = warp_sum(s);
        if (lane == 0) out[i] = -0.5 * (double)ZD * 1.8378770664093453 - 0.5 * s;
    }
}

static int fill_clf(ClfSpec& cs, int n_clf, const double* const* coef, const double* intercept, const int* target_col,
                    const int* f32) {
    if (n_clf < 0 || n_clf > MAX_CLF) { set_error("at most 4 attribute classifiers are supported"); return CPG_EINVAL; }
    memset(&cs, 0, sizeof(cs));
    cs.n_clf = n_clf;
    for (int a = 0; a < n_clf; ++a) {
        if (!coef[a]) { set_error("null classifier coefficients"); return CPG_EINVAL; }
        cs.coef[a] = coef[a]; cs.intercept[a] = intercept[a]; cs.target_col[a] = target_col[a]; cs.f32[a] = f32[a];
    }
    return CPG_OK;
}

}  // namespace cpg

using namespace cpg;

extern "C" {

int cpg_class_score_accept(cpg_ctx* ctx, cpg_stream stream, const float* z, const double* u, int64_t n, int n_clf,
                           const double* const* coef, const double* intercept, const int* target_col, const int* f32,
                           double* probs, double* accum, uint8_t* accept) {
    if (!ctx || !z || !u || !accept || n < 1) { set_error("cpg_class_score_accept: bad argument"); return CPG_EINVAL; }
    ClfSpec cs;
    int rc = fill_clf(cs, n_clf, coef, intercept, target_col, f32);
    if (rc) return rc;
    CPG_LAUNCH(k_score_accept, (unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream, z, u, n, cs, probs, accum, accept);
    return check_launch("cpg_class_score_accept");
}

int cpg_class_sample(cpg_ctx* ctx, cpg_stream stream, const float* gmm_mean, const float* gmm_sd, const float* gmm_cdf,
                     int K, int n_clf, const double* const* coef, const double* intercept, const int* target_col,
                     const int* f32, uint64_t seed, int64_t offset, int64_t n, float* z_out, double* probs,
                     double* accum, uint8_t* accept, int* comp_out, unsigned long long* n_accepted) {
    if (!ctx || !gmm_mean || !gmm_sd || !gmm_cdf || !accept || n < 1) { set_error("cpg_class_sample: bad argument"); return CPG_EINVAL; }
    if (K < 1 || K > 1024) { set_error("cpg_class_sample: 1 <= n_components <= 1024"); return CPG_EINVAL; }
    ClfSpec cs;
    int rc = fill_clf(cs, n_clf, coef, intercept, target_col, f32);
    if (rc) return rc;
    GmmSpec g{gmm_mean, gmm_sd, gmm_cdf, K};
    launch_class_sample(ctx, (cudaStream_t)stream, "k_class_sample", g, cs, seed, offset, nullptr, n, z_out, probs, accum, accept, comp_out, n_accepted);
    return check_launch("cpg_class_sample");
}

int cpg_class_regen(cpg_ctx* ctx, cpg_stream stream, const float* gmm_mean, const float* gmm_sd, const float* gmm_cdf,
                    int K, int n_clf, const double* const* coef, const double* intercept, const int* target_col,
                    const int* f32, uint64_t seed, const int64_t* draw_index, int64_t m, float* z_out, double* probs,
                    double* accum) {
    if (!ctx || !gmm_mean || !gmm_sd || !gmm_cdf || !draw_index || !z_out || m < 0) { set_error("cpg_class_regen: bad argument"); return CPG_EINVAL; }
    if (m == 0) return CPG_OK;
    if (K < 1 || K > 1024) { set_error("cpg_class_regen: 1 <= n_components <= 1024"); return CPG_EINVAL; }
    ClfSpec cs;
    int rc = fill_clf(cs, n_clf, coef, intercept, target_col, f32);
    if (rc) return rc;
    GmmSpec g{gmm_mean, gmm_sd, gmm_cdf, K};
    launch_class_sample(ctx, (cudaStream_t)stream, "k_class_regen", g, cs, seed, 0, draw_index, m, z_out, probs, accum, nullptr, nullptr, nullptr);
    return check_launch("cpg_class_regen");
}

int cpg_gmm_logpdf(cpg_ctx* ctx, cpg_stream stream, const float* x, int64_t n, const double* mean_t,
                   const double* prec_t, const double* logw_norm, int K, double* out) {
    if (!ctx || !x || !mean_t || !prec_t || !logw_norm || !out || n < 1 || K < 1) { set_error("cpg_gmm_logpdf: bad argument"); return CPG_EINVAL; }
    int64_t want = (n + LPG_WARPS * LP_PTS - 1) / (LPG_WARPS * LP_PTS);
    int grid = (int)std::min<int64_t>(want, (int64_t)ctx->sm_count * 8);
    CPG_LAUNCH(k_gmm_logpdf, grid, LPG_WARPS * 32, 0, (cudaStream_t)stream, x, n, mean_t, prec_t, logw_norm, K, out);
    return check_launch("cpg_gmm_logpdf");
}

int cpg_prior_logpdf(cpg_ctx* ctx, cpg_stream stream, const float* x, int64_t n, double* out) {
    if (!ctx || !x || !out || n < 1) { set_error("cpg_prior_logpdf: bad argument"); return CPG_EINVAL; }
    int64_t want = (n + LP_WARPS - 1) / LP_WARPS;
    int grid = (int)std::min<int64_t>(want, (int64_t)ctx->sm_count * 8);
    CPG_LAUNCH(k_prior_logpdf, grid, LP_WARPS * 32, 0, (cudaStream_t)stream, x, n, out);
    return check_launch("cpg_prior_logpdf");
}

}  // extern "C"
