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
extern int g_opt_mmd_grid;   // mmd_tc2.cu
extern int g_opt_mmd_tc;     // 0 = fp32 SIMT, 1 = persistent tcgen05 (default), 3 = one-tile-per-CTA tcgen05
extern int g_sm_count;
// tcgen05 versions of the dense layers around the latent code, fused with the element-wise work (latent_tc.cu).
// Their weight operands are pre-split (leading terms | remainders) into the shared-memory tile image once per step by
// k_prep_weights (prep.cu) and bulk-copied by the kernels; byte offsets inside Derived::lat_tiles:
constexpr int LT_F1_ROWS = 208, LT_F1_K = 160;       // [q_mu | q_logvar] rows interleaved (n = 2 j | 2 j + 1), K-major, fp16 split
constexpr int LT_F2_ROWS = 320, LT_F2_K = 112;       // W_ih[:,150:] (row = padded gate index, K = [z;c] index), K-major, fp16 split
constexpr int LT_B1_ROWS = 112, LT_B1_K = 320;       // its transpose (row = [z;c] index, K = gate index), K-major, bf16 split
constexpr int LT_B2_N = 160, LT_B2_K = 208;          // heads^T (n = hfin column, k = 2 j | 2 j + 1), MN-major, bf16 split
constexpr size_t LT_F1_TERM = (size_t)LT_F1_ROWS * LT_F1_K * 2, LT_F2_TERM = (size_t)LT_F2_ROWS * LT_F2_K * 2;
constexpr size_t LT_B1_TERM = (size_t)LT_B1_ROWS * LT_B1_K * 2, LT_B2_TERM = (size_t)LT_B2_N * LT_B2_K * 2;
constexpr size_t LT_F1_OFF = 0, LT_F2_OFF = LT_F1_OFF + 2 * LT_F1_TERM, LT_B1_OFF = LT_F2_OFF + 2 * LT_F2_TERM;
constexpr size_t LT_B2_OFF = LT_B1_OFF + 2 * LT_B1_TERM, LT_TILES_BYTES = LT_B2_OFF + 2 * LT_B2_TERM;
bool latent_uses_tc(int B);
extern int g_opt_latent_tc, g_opt_latent_rows;
extern int g_opt_adam_fused;          // api_wae.cu
extern int g_opt_chain_priority;      // api_wae.cu: 1 = the fused step's dependent chain runs on a highest-priority internal stream
int launch_latent_fwd_tc(cudaStream_t s, const float* hfin, const float* bmu, const float* blv, const float* eps, const float* c,
                         const unsigned char* tiles, int B, float* mu, float* logvar, float* z, float* zc, float* rowbias);
// hg_part (optional): per-CTA partials [latent_bwd_tc_ctas(B)][LT_HG_ROWS][LT_HG_COLS] of the head weight / bias gradients,
// contracted in the same kernel; launch_head_grad_reduce sums them in CTA order into the gradient buffers
constexpr int LT_HG_ROWS = 208, LT_HG_COLS = 176;
int latent_bwd_tc_ctas(int B);
int launch_latent_bwd_tc(cudaStream_t s, const float* drow, const float* dh0, const unsigned char* tiles, const LatentBwdArgs& lat,
                         float* dhfin, const float* hfin, float* hg_part);
void launch_head_grad_reduce(cudaStream_t s, const float* hg_part, int B, float* g_wmu, float* g_wlv, float* g_bmu, float* g_blv);
// tcgen05 random-feature kernels (rf_tc.cu): pre-split rf_w tiles, feature map + column partials, gradient
extern int g_opt_rf_tc, g_opt_rf_grid, g_opt_wd_grid;
bool rf_uses_tc(int B, int R);
size_t rf_tc_tile_bytes(int R);
int rf_tc_parts(int B);                  // rows of the column-sum partials written by launch_rf_feat_tc
void launch_prep_rf_tiles(cudaStream_t s, const float* rf_w, int R, unsigned char* tiles);
int launch_rf_feat_tc(cudaStream_t s, const float* x, const unsigned char* tiles, const float* rf_b, int B, int R, float sigma,
                      float* pre_out, float* part);
int launch_rf_grad_tc(cudaStream_t s, const float* pre, const unsigned char* tiles, const float* rf_b, const float* coef, int B, int R,
                      float sigma, float* dz);
void launch_rf_colsum_final(cudaStream_t s, const float* part, int nchunk, int R, float* out);
// weight gradients of the dense layers around the latent code as batch contractions on tcgen05 (wgrad_dense_tc.cu)
extern int g_opt_wgrad_dense_tc;     // 1 (default): on the reduction lane with these kernels; 0: head gradients inside k_latent_bwd_tc, dW_ih[:,150:] fp32 SIMT
size_t wgrad_dense_part_floats(int B);
int launch_wgrad_zc_tc(cudaStream_t s, const float* drow, const float* zc, int B, float* part, float* dwizc);
int launch_wgrad_heads_tc(cudaStream_t s, const float* dmu, const float* dlv, const float* hfin, int B, float* part, float* g_wmu,
                          float* g_wlv, float* g_bmu, float* g_blv);
void launch_compose_scalars(cudaStream_t s, const ComposeArgs& a);
// data-parallel tail: extra floats all-reduced together with the flat gradient ([0] = NLL sum; rest reserved)
constexpr int DP_TAIL = 8;
void launch_dp_pack_tail(cudaStream_t s, const float* nll_sum, float* tail);
void launch_dp_apply_tail(cudaStream_t s, const float* tail, float* scalars);
void launch_int_to_float(cudaStream_t s, const int* src, float* dst, int n);

}  // namespace cpg
