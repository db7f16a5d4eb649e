#pragma once
#include "cpg_common.cuh"

namespace cpg {

struct DecOutArgs {
    const float* hs;          // [B][L][104] decoder hidden states by step
    const uint8_t* out_keep;  // [B][L][102] 1 = keep, or null (eval mode)
    float keep_scale;         // 1 / (1 - p_out_dropout)
    const float* fc_w;        // [VMAX][104] zero padded
    const float* fc_b;        // [VMAX]
    const uint8_t* tgt;       // [B][L]
    const float* ntok;        // device scalar: global number of non-<pad> targets
    int nprod;                // 0 / 3: full-precision split products, 1: leading bf16 product only (g_opt_matmul_terms)
    const int* ntok_i;        // ... or, when not null, the same count as the integer the token preparation produced
    const float* dlogits_in;  // [B][L][V] upstream gradient (module API) or null
    float* logits_out;        // [B][L][V] or null
    float* dh_out;            // [B][L][104] or null (forward only)
    float* part_w;            // [parts][VMAX][104]
    float* part_b;            // [parts][VMAX]
    float* part_nll;          // [parts] or null
    int fused_ce;             // compute CE + its gradient in place
    int B, L, V;
};

int dec_out_parts(int B, int L, int sm_count);
void launch_dec_out(cudaStream_t s, const DecOutArgs& a, int sm_count);
// tcgen05 variant (dec_out_tc.cu): one persistent CTA per SM; *parts_out = number of per-CTA partials written
int launch_dec_out_tc(cudaStream_t s, const DecOutArgs& a, int sm_count, int* parts_out);
extern int g_opt_dec_out_tc;      // 0 = never, 1 = auto (rows >= 8192), 2 = always
void launch_dec_out_reduce(cudaStream_t s, const DecOutArgs& a, int sm_count, float* dW, float* db, float* nll_sum);

}  // namespace cpg
