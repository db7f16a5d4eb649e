// Internal launch wrappers shared between the .cu files (not part of the C ABI).
#pragma once
#include "cpg_common.cuh"

namespace cpg {

// Flat layout of the VAE parameters (the unique tensors of RNN_VAE.vae_params(),
// models/model.py:88-94), each segment padded to a multiple of 4 floats.
enum ParamId {
    P_EMB = 0,
    P_ENC_WIH_F, P_ENC_WHH_F, P_ENC_BIH_F, P_ENC_BHH_F,
    P_ENC_WIH_R, P_ENC_WHH_R, P_ENC_BIH_R, P_ENC_BHH_R,
    P_QMU_W, P_QMU_B, P_QLV_W, P_QLV_B,
    P_DEC_WIH, P_DEC_WHH, P_DEC_BIH, P_DEC_BHH,
    P_FC_W, P_FC_B,
    P_COUNT
};
struct ParamLayout {
    int64_t off[P_COUNT];
    int64_t size[P_COUNT];     // unpadded element counts
    int64_t total;             // padded total
};
ParamLayout make_layout(int n_vocab);

// Derived, per-step weight forms (recomputed from the parameters every step).
struct Derived {
    float* t_enc[2];      // [V][3*80]   token -> input-side gate pre-activations (+ folded biases)
    float* t_dec;         // [V][3*104]
    float* whh_t_enc[2];  // [80][240]    W_hh^T
    float* whh_t_dec;     // [104][312]   W_hh^T, zero padded
    float* whh_dec;       // [312][104]   W_hh, zero padded (natural layout)
    float* wizc_t;        // [104][312]   (W_ih[:,150:])^T, zero padded
    float* wizc;          // [312][104]   W_ih[:,150:], zero padded
    float* bhn_enc[2];    // [80]
    float* bhn_dec;       // [104]
    float* fc_w;          // [VMAX][104]  zero padded
    float* fc_b;          // [VMAX]
    unsigned char* lat_tiles;   // pre-split operand tiles of latent_tc.cu (latent.h: LT_*), or null
};
size_t derived_floats(int n_vocab);

struct GruSeq {            // one recurrence (a direction of the encoder, or the decoder)
    const uint8_t* tok;    // [B][L] tokens feeding this GRU
    const float* table;    // [V][3*HP]
    const float* rowbias;  // [B][3*HP] or null
    const float* whh_t;    // [HP][3*HP]
    const float* whh;      // [3*HP][HP] natural (backward)
    const float* bhn;      // [HP]
    const float* h0;       // [B][HP] or null (zeros)
    float* hs;             // [B][L][HP] hidden after each step, indexed by STEP (not time) ; may be null
    float* gates;          // [B][L][4][HP] r, z, n, hn by step ; may be null
    float* hfin;           // final hidden -> hfin[b*hfin_stride + j] ; may be null
    int hfin_stride;
    int reverse;           // step s reads time L-1-s
    // backward
    const float* dh_out;   // [B][L][HP] by step, or null
    const float* dh_fin;   // [B][dh_fin_stride] gradient wrt final hidden, or null
    int dh_fin_stride;
    float* dg;             // [B][L][4][HP] dr_pre, dz_pre, dn_pre, dhn by step
    float* dh0;            // [B][HP] or null
    float* drow;           // [B][3*HP] sum over steps of (dr_pre,dz_pre,dn_pre), or null
};

// Per-step scalars that change between replays of a captured iteration (CUDA graph): they live in device memory and one
// tiny kernel at the head of the graph refreshes them; g_dyn is non-null while such an iteration is being enqueued, and
// the kernels that take these values by argument then read them from there instead.
struct StepDyn {
    float beta;
    float step_size[3], bc2_sqrt[3];      // Adam: ordinary tensors, first / second update of the duplicated embedding
    uint32_t noise_step;
};
extern const StepDyn* g_dyn;
// 3 (default): every tensor-core contraction of the recurrences / decoder-output layer as the three split-bf16 products
// (fp32-grade, the parity configuration); 1: the leading bf16 product only ("bf16 matmul tiles" of BASELINE.json configs[2]:
// a reduced-precision mode, NOT covered by the parity bars)
extern int g_opt_matmul_terms;

// per-iteration noise (noise.cu); part bit 0 = eps / c / word dropout, bit 1 = z_prior x2 / out-dropout mask
struct StepNoiseArgs {
    uint64_t seed; uint32_t step;
    int B, L;
    float p_word, p_out;
    float* eps; float* c; uint8_t* word_drop; uint8_t* out_keep; float* zp_full; float* zp_rf;
};
// part bits: what to generate
constexpr int NOISE_WORD = 1;      // word-dropout mask (the token preparation needs it first)
constexpr int NOISE_LATE = 2;      // z_prior x2, out-dropout mask (the bulk of the work)
constexpr int NOISE_LATENT = 4;    // eps, c
constexpr int NOISE_ALL = 7;
void launch_step_noise(cudaStream_t s, const StepNoiseArgs& a, int part);

void launch_prep_tokens(cudaStream_t s, const int64_t* tokens, const uint8_t* word_drop, int B, int L, int V,
                        uint8_t* tok, uint8_t* tokd, uint8_t* tgt, int* ntok, int* err, const StepNoiseArgs* gen = nullptr);
// part: 1 = the forms the encoder recurrence reads, 2 = all the others, 3 = everything in one launch
void launch_prep_weights(cudaStream_t s, const float* params, const ParamLayout& lay, int V, const Derived& d, int part = 3);

void launch_gru_fwd_enc(cudaStream_t s, const GruSeq* two_dirs, int B, int L);
void launch_gru_fwd_dec(cudaStream_t s, const GruSeq& seq, int B, int L);
void launch_gru_bwd_enc(cudaStream_t s, const GruSeq* two_dirs, int B, int L);
void launch_gru_bwd_dec(cudaStream_t s, const GruSeq& seq, int B, int L);
// tcgen05 recurrences (gru_tc.cu); GruSeq::whh must hold the natural [3H][H] recurrent weights
int launch_gru_fwd_enc_tc(cudaStream_t s, const GruSeq* two_dirs, int B, int L, int V);
int launch_gru_fwd_dec_tc(cudaStream_t s, const GruSeq& seq, int B, int L, int V);
int launch_gru_bwd_enc_tc(cudaStream_t s, const GruSeq* two_dirs, int B, int L, int round_dg);   // round_dg: dg stored as tf32
int launch_gru_bwd_dec_tc(cudaStream_t s, const GruSeq& seq, int B, int L, int round_dg);
// BPTT with the dW_hh / token-table gradient contractions fused in (gru_bwd_fused.cu): no dg planes in HBM.
// part_w: [ctas][3*HP][HP], part_t: [ctas][V][4*HP] per direction, ctas = bptt_fused_ctas_*(B).
int bptt_fused_ctas_enc(int B);
int bptt_fused_ctas_dec(int B);
int launch_gru_bwd_enc_fused(cudaStream_t s, const GruSeq* two_dirs, const uint8_t* tok, int B, int L, int V,
                             float* const part_w[2], float* const part_t[2]);
int launch_gru_bwd_dec_fused(cudaStream_t s, const GruSeq& seq, const uint8_t* tok, int B, int L, int V, float* part_w,
                             float* part_t);
extern int g_opt_gru_tc;
extern int g_opt_bptt_fused;
extern int g_opt_graph;
extern int g_opt_side_stream;

// C[M,N] = alpha * op(A)[M,K] * op(B)[K,N] + beta * C ; generic strides (elements):
// A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn], C(m,n) = C[m*ldc + n]; optional bias[n].
// split_k > 1 uses `ws` (split_k*M*N floats) and a deterministic reduction.
void launch_sgemm(cudaStream_t s, int M, int N, int K, float alpha, const float* A, int64_t sam, int64_t sak,
                  const float* B, int64_t sbk, int64_t sbn, float beta, float* C, int64_t ldc,
                  const float* bias, int split_k, float* ws);
// C = alpha (A B + A2 B2) + beta C: two products of equal shape / strides accumulated in one launch
void launch_sgemm_sum2(cudaStream_t s, int M, int N, int K, float alpha, const float* A, const float* A2, int64_t sam, int64_t sak,
                       const float* B, const float* B2, int64_t sbk, int64_t sbn, float beta, float* C, int64_t ldc);
// C = alpha A B + bias and C2 = alpha A B2 + bias2 (shared A) in one launch
void launch_sgemm_pair(cudaStream_t s, int M, int N, int K, float alpha, const float* A, int64_t sam, int64_t sak,
                       const float* B, const float* B2, int64_t sbk, int64_t sbn, float* C, float* C2, int64_t ldc,
                       const float* bias, const float* bias2);
// out[n] = sum_m A[m*lda + n], deterministic two-stage
void launch_colsum(cudaStream_t s, const float* A, int M, int N, int64_t lda, float* out, float* ws, int nchunk);

}  // namespace cpg
