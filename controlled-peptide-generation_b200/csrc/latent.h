#pragma once
#include "cpg_common.cuh"
#include "kernels.h"

namespace cpg {

// slots of the per-step scalar block (matches CPG_SC_* in include/cpg_b200.h)
enum ScalarSlot {
    SC_LOSS = 0, SC_RECON, SC_KL, SC_MMD, SC_MMDRF, SC_LOGVAR_L1, SC_LOGVAR_KL, SC_Z_MU_L1, SC_Z_LOGVAR,
    SC_BETA, SC_GRAD_NORM, SC_NTOK, SC_NLL_SUM, SC_COUNT = 16
};

struct LatentBwdArgs {
    const float* mu; const float* logvar; const float* eps;   // eps null: z = mu
    const float* dzc;       // [B][104] gradient at [z;c] from the decoder, or null
    const float* dz_rf;     // [B][100] RF-MMD gradient (already weighted), or null
    const float* dz_ext;    // [B][100] external upstream gradient at z, or null
    const float* dmu_ext;   // external upstream gradients at mu / logvar, or null
    const float* dlv_ext;
    float w_kl;             // beta if z_regu == kl else 0
    float w_klsm;           // lambda_logvar_KL
    float w_l1;             // lambda_logvar_L1
    int B, B_global;
    float* dmu; float* dlv; // [B][100] out
    const StepDyn* dyn;     // set by the launcher (see kernels.h)
};

struct ComposeArgs {
    const float* ntok; const float* nll_sum; const float* lat_sums; const float* mmd; const float* mmdrf;
    float beta, lambda_l1, lambda_kl;
    int z_regu, B_global;
    float* out;
    const StepDyn* dyn;     // set by the launcher
};

void launch_reparam(cudaStream_t s, const float* mu, const float* logvar, const float* eps, const float* c, int B,
                    float* z, float* zc);
void launch_make_zc(cudaStream_t s, const float* z, const float* c, int B, float* zc);
void launch_latent_stats(cudaStream_t s, const float* mu, const float* logvar, int B, float* part, int nparts,
                         float* sums5);
void launch_latent_bwd(cudaStream_t s, const LatentBwdArgs& a);
void launch_rf_colsum(cudaStream_t s, const float* pre, const float* rf_b, int B, int R, float sigma, float* part,
                      int nchunk, float* out);
void launch_rf_loss(cudaStream_t s, const float* sum1, const float* sum2, int R, int B_global, float sigma, float w,
                    float* coef, float* loss_out);
void launch_rf_grad_prep(cudaStream_t s, float* pre, const float* rf_b, const float* coef, int B, int R, float sigma);
size_t mmd_full_ws_floats(int N);
void launch_mmd_full_simt(cudaStream_t s, const float* z, const float* zp, int N, float sigma, float* ws, float* out);
// tcgen05 / TMA version (mmd_tc.cu) and the dispatcher used by the API (option "mmd_tensor_core")
size_t mmd_tc_ws_floats(int N);
int launch_mmd_full_tc(cudaStream_t s, const float* z, const float* zp, int N, float sigma, float* ws, float* out);
size_t mmd_ws_floats(int N);
int launch_mmd_full(cudaStream_t s, const float* z, const float* zp, int N, float sigma, float* ws, float* out);
int launch_mmd_full_tc2(cudaStream_t s, const float* z, const float* zp, int N, float sigma, int sm_count, float* ws,
                        float* out);
// dz = w * d mmd_full_kernel(z, zp) / dz   (fp32 SIMT)
void launch_mmd_full_grad(cudaStream_t s, const float* z, const float* zp, int N, float sigma, float w, float* dz);
extern int g_opt_mmd_tc;     // 0 = fp32 SIMT, 1 = persistent tcgen05 (default), 3 = one-tile-per-CTA tcgen05
extern int g_sm_count;
void launch_compose_scalars(cudaStream_t s, const ComposeArgs& a);
// data-parallel tail: extra floats all-reduced together with the flat gradient ([0] = NLL sum; rest reserved)
constexpr int DP_TAIL = 8;
void launch_dp_pack_tail(cudaStream_t s, const float* nll_sum, float* tail);
void launch_dp_apply_tail(cudaStream_t s, const float* tail, float* scalars);
void launch_int_to_float(cudaStream_t s, const int* src, float* dst, int n);

}  // namespace cpg
