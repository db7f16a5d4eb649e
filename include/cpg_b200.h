/* cpg_b200 -- C ABI of the B200-native (sm_100a) hot path of IBM/controlled-peptide-generation.
 *
 * The reference (pure Python on PyTorch / scikit-learn) has no FFI boundary of its own: its hot path
 * is reached through the Python API of models/model.py, losses.py, train_vae.py, density_modeling.py
 * and sample_pipeline.py.  This library sits UNDER that API: every entry point below replaces the
 * vendor-library calls the cited reference lines make, and the Python package next to it
 * (controlled-peptide-generation_b200/) keeps the reference's module / class / function names and
 * binds these symbols with ctypes.
 *
 * Conventions
 *   - All pointers are DEVICE pointers unless a parameter is documented as host.  The caller
 *     (PyTorch) owns every buffer; the library owns only the per-context workspace.
 *   - `stream` is a cudaStream_t passed as void*.  Calls enqueue work and return; no hidden syncs
 *     (workspace growth, which happens when a larger batch is seen first, synchronises once).
 *   - Return value: 0 on success, negative CPG_E* code on failure; cpg_last_error() gives the text.
 *   - Geometry is the reference configuration (cfg.py:258-281): emb 150, encoder bi-GRU h 80,
 *     z 100, c 2, decoder GRU h 102; vocabulary n_vocab <= 32, sequence length <= 32.
 *   - fp32 storage and arithmetic except where stated (fp64 for the GMM / score path).
 */
#ifndef CPG_B200_H
#define CPG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPG_ABI_VERSION 1

#define CPG_OK 0
#define CPG_EINVAL (-1)   /* bad argument (shape, null pointer, unsupported geometry) */
#define CPG_ECUDA (-2)    /* CUDA runtime error */
#define CPG_ENOMEM (-3)   /* workspace allocation failed */
#define CPG_ETOKEN (-4)   /* token id outside [0, n_vocab) seen by a previous call */

typedef struct cpg_ctx cpg_ctx;
typedef void* cpg_stream;

/* ---- lifecycle ------------------------------------------------------------------------------ */
int cpg_abi_version(void);
const char* cpg_last_error(void);
int cpg_create(cpg_ctx** out, int device_ordinal);
int cpg_destroy(cpg_ctx* ctx);
int cpg_sm_count(const cpg_ctx* ctx);
/* bytes currently held by the context's workspace */
int64_t cpg_workspace_bytes(const cpg_ctx* ctx);
/* number of kernel launches enqueued by this context since creation */
int64_t cpg_launch_count(const cpg_ctx* ctx);
/* Generation of the activation stash held by the context: changes with every cpg_wae_forward(keep_for_backward=1)
 * / cpg_wae_step_phase1, -1 when no stash is valid (an inference call or a re-layout dropped it).  A caller that
 * defers cpg_wae_backward (autograd) records it after the forward and compares before the backward. */
int64_t cpg_stash_generation(const cpg_ctx* ctx);
/* device->host check of the sticky token-range flag (synchronises `stream`) */
int cpg_check_errors(cpg_ctx* ctx, cpg_stream stream);

/* ---- flat parameter layout ------------------------------------------------------------------
 * The 19 unique tensors of RNN_VAE.vae_params() (models/model.py:88-94) live in ONE flat fp32
 * buffer, each segment padded to a multiple of 4 floats, in this order:
 *   word_emb.weight, encoder.rnn.{weight_ih,weight_hh,bias_ih,bias_hh}_l0, the same four with
 *   _reverse, encoder.q_mu.{weight,bias}, encoder.q_logvar.{weight,bias},
 *   decoder.rnn.{weight_ih,weight_hh,bias_ih,bias_hh}_l0, decoder.fc.1.{weight,bias}.
 * Gradients and Adam moments use the same layout. */
#define CPG_N_PARAM_TENSORS 19
int64_t cpg_vae_param_count(int n_vocab);
int cpg_vae_param_layout(int n_vocab, int64_t offsets[CPG_N_PARAM_TENSORS], int64_t sizes[CPG_N_PARAM_TENSORS]);

/* ---- WAE forward / backward (module-level API) ----------------------------------------------
 * Replaces RNN_VAE.forward (models/model.py:146-195): forward_encoder -> sample_z ->
 * forward_decoder, i.e. nn.Embedding + bi-nn.GRU + 2 nn.Linear (models/encoder.py:38-52),
 * the reparameterisation (model.py:107-112), WordDropout + nn.Embedding + nn.GRU + nn.Dropout +
 * nn.Linear (models/decoder.py:56-84).  All randomness is an INPUT. */
typedef struct {
    const int64_t* tokens;     /* [B, L] int64, batch.text of the reference                       */
    const float* eps;          /* [B, 100] N(0,1) noise of sample_z; NULL -> z = mu ("max")       */
    const float* c;            /* [B, 2] code fed to the decoder (one-hot prior or softmax)       */
    const uint8_t* word_drop;  /* [B, L] 1 = replace input token by <unk>; NULL = no word dropout */
    const uint8_t* out_keep;   /* [B, L, 102] 1 = keep (nn.Dropout mask); NULL = eval mode        */
    float p_out_dropout;       /* 0.3 in cfg.py:279; kept activations scale by 1/(1-p)            */
} cpg_wae_inputs;

/* mu, logvar, z: [B,100]; logits: [B,L,V] or NULL.  keep_for_backward != 0 stashes activations in
 * the context for a following cpg_wae_backward on the same (B, L). */
int cpg_wae_forward(cpg_ctx* ctx, cpg_stream stream, const float* params, int n_vocab, int B, int L,
                    const cpg_wae_inputs* in, float* mu, float* logvar, float* z, float* logits,
                    int keep_for_backward);
/* encoder only (sample_pipeline.py:49-70 / build_index.py:93-118 extraction): mu, logvar */
int cpg_wae_encode(cpg_ctx* ctx, cpg_stream stream, const float* params, int n_vocab, int B, int L,
                   const int64_t* tokens, float* mu, float* logvar);
/* teacher-forced decoder only (RNN_VAE.forward_decoder, models/model.py:128-133): logits [B,L,V] from given z, c */
int cpg_wae_decode_teacher(cpg_ctx* ctx, cpg_stream stream, const float* params, int n_vocab, int B, int L,
                           const cpg_wae_inputs* in, const float* z, float* logits);
/* Backward of the last cpg_wae_forward(keep_for_backward=1): upstream gradients (any may be NULL)
 * -> flat parameter gradient (overwritten).  This is what loss.backward() (train_vae.py:40) does. */
int cpg_wae_backward(cpg_ctx* ctx, cpg_stream stream, const float* params, int n_vocab, int B, int L,
                     const cpg_wae_inputs* in, const float* d_mu, const float* d_logvar, const float* d_z,
                     const float* d_logits, float* grads);

/* ---- fused training iteration (train_vae.py:24-42) ------------------------------------------ */
#define CPG_ZREGU_KL 0
#define CPG_ZREGU_MMD 1
#define CPG_ZREGU_MMDRF 2
typedef struct {
    float lr, beta1, beta2, adam_eps;   /* optim.Adam(lr=cfg.vae.lr) defaults: 1e-3, .9, .999, 1e-8  */
    float clip_norm;                    /* cfg.shared.clip_grad = 5.0                                 */
    float beta;                         /* utils.anneal(cfg.vae.beta, it)                             */
    float lambda_logvar_l1;             /* cfg.vae.lambda_logvar_L1                                   */
    float lambda_logvar_kl;             /* cfg.vae.lambda_logvar_KL                                   */
    int z_regu;                         /* CPG_ZREGU_* (cfg.vae.z_regu_loss)                          */
    float mmd_sigma;                    /* cfg.losses.wae_mmd.sigma                                   */
    int rf_dim;                         /* cfg.losses.wae_mmd.rf_dim                                  */
    int compute_full_mmd;               /* the reference always evaluates it (train_vae.py:29)        */
    int adam_step;                      /* 1-based iteration count (Adam's `step` of ordinary params) */
    int global_batch;                   /* B summed over data-parallel ranks (= B on one GPU)         */
} cpg_train_hparams;

typedef struct {
    const float* z_prior_full;  /* [B,100] randn_like(z) of the full-kernel MMD call (losses.py:37) */
    const float* z_prior_rf;    /* [B,100] randn_like(z) of the RF MMD call                         */
    const float* rf_w;          /* [100, R] cached random features (losses.py:75)                   */
    const float* rf_b;          /* [R]      (losses.py:76)                                          */
} cpg_loss_noise;

/* slots of the scalar block written by the step (device float[CPG_SC_COUNT]) */
#define CPG_SC_LOSS 0
#define CPG_SC_RECON 1
#define CPG_SC_KL 2
#define CPG_SC_MMD 3
#define CPG_SC_MMDRF 4
#define CPG_SC_LOGVAR_L1 5
#define CPG_SC_LOGVAR_KL 6
#define CPG_SC_Z_MU_L1 7
#define CPG_SC_Z_LOGVAR 8
#define CPG_SC_BETA 9
#define CPG_SC_GRAD_NORM 10
#define CPG_SC_NTOK 11
#define CPG_SC_NLL_SUM 12
#define CPG_SC_COUNT 16

/* One whole iteration on one GPU: forward, the five losses, backward, clip_grad_norm_, Adam.
 * params / grads / adam_m / adam_v: flat buffers (cpg_vae_param_count floats), updated in place.
 * scalars: device float[CPG_SC_COUNT].  mu/logvar/z/logits: optional outputs (may be NULL). */
int cpg_wae_train_step(cpg_ctx* ctx, cpg_stream stream, float* params, float* grads, float* adam_m, float* adam_v,
                       int n_vocab, int B, int L, const cpg_wae_inputs* in, const cpg_loss_noise* noise,
                       const cpg_train_hparams* hp, float* scalars, float* mu, float* logvar, float* z,
                       float* logits);

/* The iteration with its Philox noise as ONE call (what train_vae's default loop issues per step): the per-step noise
 * buffers are regenerated from (seed, noise_step) and the step of cpg_wae_train_step runs on them.  From the third call
 * with the same buffers / shapes / settings on, the whole step is a replay of a captured CUDA graph (option "cuda_graph");
 * only hp->beta, hp->adam_step and noise_step may change between replays without a re-capture. */
typedef struct {
    float* eps; float* c; uint8_t* word_drop; uint8_t* out_keep;      /* per-step noise, regenerated every call   */
    float* z_prior_full; float* z_prior_rf;                            /* (z_prior_full may be NULL)               */
    const float* rf_w; const float* rf_b;                              /* cached random features (losses.py:75-76) */
} cpg_step_noise_buffers;
int cpg_wae_train_step_philox(cpg_ctx* ctx, cpg_stream stream, float* params, float* grads, float* adam_m, float* adam_v,
                              int n_vocab, int B, int L, const int64_t* tokens, const cpg_step_noise_buffers* noise,
                              const cpg_train_hparams* hp, uint64_t seed, uint32_t noise_step, float p_word, float p_out,
                              float* scalars);

/* The same iteration split at its data-parallel exchange points:
 *   phase1: forward through the decoder GRU + local loss statistics.  Writes `coupled`
 *           (device float[cpg_coupled_count(rf_dim)]) = [n_tok, nll placeholder, 5 latent sums,
 *           RF feature sums of z (R), of z_prior (R)] -- the caller all-reduces (sums) it.
 *   phase2: CE fwd/bwd with the GLOBAL token count, BPTT, weight gradients -> grads (local sum
 *           contribution; the caller all-reduces grads), scalars from the reduced statistics.
 *   cpg_clip_adam_step: global-norm clip + Adam with the duplicated-embedding semantics. */
int64_t cpg_coupled_count(int rf_dim);
int cpg_wae_step_phase1(cpg_ctx* ctx, cpg_stream stream, const float* params, int n_vocab, int B, int L,
                        const cpg_wae_inputs* in, const cpg_loss_noise* noise, const cpg_train_hparams* hp,
                        float* coupled, float* mu, float* logvar, float* z);
int cpg_wae_step_phase2(cpg_ctx* ctx, cpg_stream stream, const float* params, float* grads, int n_vocab, int B, int L,
                        const cpg_wae_inputs* in, const cpg_loss_noise* noise, const cpg_train_hparams* hp,
                        const float* coupled, float* scalars, float* logits);
int cpg_clip_adam_step(cpg_ctx* ctx, cpg_stream stream, float* params, float* grads, float* adam_m, float* adam_v,
                       int n_vocab, const cpg_train_hparams* hp, float* grad_norm_out);

/* Per-step scalars in device memory, for a caller that captures the iteration into a CUDA graph ITSELF (e.g. the
 * data-parallel step with its collectives, cpg_b200/parallel.py).  While cpg_step_dyn_use(ctx, 1) is in effect every kernel
 * the library enqueues reads beta, the Adam bias corrections and the noise counter from the context's device block
 * instead of from the host arguments of the call; cpg_step_dyn_write refreshes that block (one tiny kernel on `stream`,
 * outside the captured graph) before each replay.  cpg_wae_train_step_philox does this internally. */
int cpg_step_dyn_write(cpg_ctx* ctx, cpg_stream stream, const cpg_train_hparams* hp, uint32_t noise_step);
int cpg_step_dyn_use(cpg_ctx* ctx, int on);

/* Data-parallel plumbing (no reference counterpart: the reference is single-process).
 * cpg_side_stream / cpg_aux_stream: the context's two internal streams (cudaStream_t; NULL when option side_stream = 0).
 *   phase1 produces coupled[1..] on the side stream (under the decoder recurrence of the caller's stream) and coupled[0],
 *   the token count, on the aux stream right after the token preparation; it does NOT join them into the caller's stream.
 *   A data-parallel caller enqueues the all-reduce of coupled[0:1] on the aux stream and of coupled[1:] on the side
 *   stream (so that both exchanges start as soon as their operands exist and stay off the caller's stream); phase2
 *   orders its consumers behind those streams itself (the decoder-output layer behind the aux stream, the RF-MMD
 *   chain runs on the side stream, everything is joined before phase2 returns its scalars).  With the streams off
 *   (NULL) everything happens on the caller's stream.
 * cpg_dp_pack_tail: after phase2, writes float[cpg_dp_tail_count()] = {local NLL sum, 0...}; the caller appends
 *   it to the flat gradient so that ONE all-reduce carries both.
 * cpg_dp_apply_tail: after the all-reduce, re-bases scalars[RECON], [LOSS], [NLL_SUM] on the global NLL sum. */
void* cpg_side_stream(cpg_ctx* ctx);
void* cpg_aux_stream(cpg_ctx* ctx);
int cpg_dp_tail_count(void);
int cpg_dp_pack_tail(cpg_ctx* ctx, cpg_stream stream, float* tail);
int cpg_dp_apply_tail(cpg_ctx* ctx, cpg_stream stream, const float* tail_reduced, float* scalars);

/* ---- individual losses (losses.py) ---------------------------------------------------------- */
/* recon_dec (losses.py:18-31): mean NLL over non-<pad> next-token targets; optional d_logits.
 * loss_out: device float[2] = {mean nll, n_tok}. */
int cpg_softmax_xent(cpg_ctx* ctx, cpg_stream stream, const float* logits, const int64_t* tokens, int B, int L,
                     int n_vocab, float* loss_out, float* d_logits);
/* kl_gaussianprior, kl_gaussian_sharedmu (losses.py:8-15), logvar L1 (train_vae.py:33),
 * mean|mu|, mean logvar -> device float[5] */
int cpg_latent_stats(cpg_ctx* ctx, cpg_stream stream, const float* mu, const float* logvar, int B, float* out5);
/* mmd_full_kernel (losses.py:47-56,96-108), gaussian kernel, as executed by the reference */
int cpg_mmd_full(cpg_ctx* ctx, cpg_stream stream, const float* z, const float* z_prior, int B, float sigma,
                 float* loss_out);
/* d mmd_full_kernel / d z: what loss.backward() propagates when cfg.vae.z_regu_loss = 'mmd' (the reference default
 * 'mmdrf' only logs the full-kernel value); dz: [B,100] */
int cpg_mmd_full_grad(cpg_ctx* ctx, cpg_stream stream, const float* z, const float* z_prior, int B, float sigma, float* dz);
/* mmd_rf (losses.py:59-93); dz may be NULL; dz = d loss / d z */
int cpg_mmd_rf(cpg_ctx* ctx, cpg_stream stream, const float* z, const float* z_prior, const float* rf_w,
               const float* rf_b, int B, int rf_dim, float sigma, float* loss_out, float* dz);

/* ---- perf-mode noise (Philox4x32-10 counter RNG) ----------------------------------------------
 * Replaces the per-iteration host/device RNG draws of SURVEY.md appendix A: torch.randn (eps,
 * model.py:111), np.random.multinomial (c, model.py:125), np.random.binomial (word dropout,
 * decoder.py:124-127), nn.Dropout mask (decoder.py:44), torch.randn_like (z_prior x2, losses.py:37).
 * Every tensor is a pure function of (seed, step, index).  Any output pointer may be NULL. */
int cpg_fill_step_noise(cpg_ctx* ctx, cpg_stream stream, uint64_t seed, uint32_t step, int B, int L, float p_word,
                        float p_out, float* eps, float* c, uint8_t* word_drop, uint8_t* out_keep,
                        float* z_prior_full, float* z_prior_rf);
/* Same tensors, same values, drawn LATER: the call only records the request.  The entry point that reads the buffers next
 * draws the word-dropout mask inside its token preparation and everything else on the context's side stream once that
 * preparation is through (joined where it is consumed) -- drawn at call time, the large noise kernel would slow the
 * preparation kernels the encoder recurrence waits for.  ONLY for callers whose next use of these buffers is
 * cpg_wae_step_phase1/2, cpg_wae_train_step, cpg_wae_forward or cpg_wae_backward on the same context and stream with
 * these very buffers as inputs (in any other case the library draws everything immediately at that later call; a caller
 * that wants to READ the tensors itself must use cpg_fill_step_noise). */
int cpg_fill_step_noise_overlapped(cpg_ctx* ctx, cpg_stream stream, uint64_t seed, uint32_t step, int B, int L,
                                   float p_word, float p_out, float* eps, float* c, uint8_t* word_drop,
                                   uint8_t* out_keep, float* z_prior_full, float* z_prior_rf);
/* N(0,1) / U[0,1)*scale fills, e.g. rf_w = randn(100,R), rf_b = 2*pi*rand(R) (losses.py:75-76) */
int cpg_fill_normal(cpg_ctx* ctx, cpg_stream stream, uint64_t seed, uint32_t stream_id, int64_t n, float* out);
int cpg_fill_uniform(cpg_ctx* ctx, cpg_stream stream, uint64_t seed, uint32_t stream_id, float scale, int64_t n,
                     float* out);

/* ---- decoding from (z, c) -------------------------------------------------------------------
 * Replaces RNN_VAE.sample_G (models/model.py:225-385) + GRUDecoder.forward_sample
 * (models/decoder.py:86-109) + Beam (models/Beam.py:56-132) as driven by generate_sentences
 * (model.py:197-223) and sample_pipeline.decode_from_z (sample_pipeline.py:129-139).  Eval mode. */
/* Beam search (beam_size must be 5, n_best <= 5).  out_tokens: int32 [n][n_best][L+1], -1 padded,
 * hypothesis starts with <start>; out_len: [n][n_best]; out_score: [n][n_best] summed log-probs.
 * Ties in top-k: larger score first, then lower flat (beam, word) index. */
int cpg_beam_decode(cpg_ctx* ctx, cpg_stream stream, const float* params, int n_vocab, int n, int L, const float* z,
                    const float* c, int beam_size, int n_best, int* out_tokens, int* out_len, float* out_score);
/* mode 1 = greedy (argmax), 2 = categorical (softmax(logits/temp), Philox uniforms keyed by seed).
 * out_tokens: int32 [n][L+1] with <start> first and <pad> after <eos>; *out_steps = number of steps
 * until every sample had finished (the reference truncates its output there, model.py:361-363). */
int cpg_sample_decode(cpg_ctx* ctx, cpg_stream stream, const float* params, int n_vocab, int n, int L, const float* z,
                      const float* c, int mode, float temp, uint64_t seed, int* out_tokens, int* out_steps);

/* Soft sampling modes of sample_G (models/model.py:330-359), forward only: mode 3 = none_softmax (the hard token is never
 * updated, as in the reference), 4 = greedy_softmax, 5 = categorical_softmax.  out_soft: float [n][L+1][n_vocab] =
 * softmax(logits / temp) of every step (one-hot <start> first, zeros from the <eos> step on); each step's softmax is fed
 * back as a soft embedding (decoder.py:86-92, mutils.soft_embed).  out_tokens / out_steps as cpg_sample_decode. */
int cpg_soft_decode(cpg_ctx* ctx, cpg_stream stream, const float* params, int n_vocab, int n, int L, const float* z,
                    const float* c, int mode, float temp, uint64_t seed, int* out_tokens, float* out_soft, int* out_steps);

/* ---- normalising flow on the latent code (models/flow.py:30-160) ------------------------------
 * All layers in one launch.  kind[l]: 0 planar, 1 radial (HOST array).  vec_a / vec_b: HOST arrays of DEVICE pointers to
 * [100] floats -- planar (weight, scale), radial (initial point, unused).  scalar_a / scalar_b: HOST arrays -- planar
 * (bias, weight.scale), radial (alpha, beta).  train != 0 also writes loss_out = mean_b sum_l log(|det J_l| + 1e-7) with
 * the radial determinant exactly as the reference evaluates it (batch-wide Frobenius norm of the radii, flow.py:87). */
int cpg_flow_forward(cpg_ctx* ctx, cpg_stream stream, const float* z_in, int B, int n_layers, const int* kind,
                     const float* const* vec_a, const float* const* vec_b, const float* scalar_a, const float* scalar_b,
                     int train, float* z_out, float* loss_out);

/* ---- CNN attribute classifier forward (models/classifier.py:39-60, eval mode) ---------------- */
/* conv weights [100][1][w][150] for w = 3,4,5; fc [2][300]; table_ws: 12*n_vocab*100 floats scratch. */
int cpg_cnn_classifier_fwd(cpg_ctx* ctx, cpg_stream stream, const float* emb, const float* conv_w3, const float* conv_b3,
                           const float* conv_w4, const float* conv_b4, const float* conv_w5, const float* conv_b5,
                           const float* fc_w, const float* fc_b, int n_vocab, int B, int L, const int64_t* tokens,
                           float* table_ws, float* logits);

/* ---- CLaSS latent sampling (density_modeling.py) ---------------------------------------------
 * Classifier spec, n_clf <= 4: coef[a] = DEVICE pointer to 100 doubles (the array of pointers itself
 * is HOST memory), intercept[a], target_col[a] in {0,1} (column of predict_proba kept,
 * sample_pipeline.py:290), f32[a] != 0 -> float32 score arithmetic (classifier fitted on float32). */
/* parity mode of RejSampleBase.rejection_sample (density_modeling.py:50-60) after the draw:
 * z [n][100] fp32 and u [n] fp64 are the reference's own draws.  probs: [n_clf][n] (may be NULL),
 * accum [n] (may be NULL), accept uint8 [n] = (u < prod_a p_a). */
int cpg_class_score_accept(cpg_ctx* ctx, cpg_stream stream, const float* z, const double* u, int64_t n, int n_clf,
                           const double* const* coef, const double* intercept, const int* target_col, const int* f32,
                           double* probs, double* accum, uint8_t* accept);
/* perf mode: draw n samples with global indices [offset, offset+n) from the diag-GMM
 * (mean, sd = sqrt(cov): fp32 [K][100]; cdf: cumulative weights [K]) with Philox(seed), score, accept.
 * z_out / probs / accum / comp_out / n_accepted may be NULL.  Draws are i.i.d. (the reference returns
 * them grouped by component, sklearn mixture/_base.py:461-511; same distribution). */
int cpg_class_sample(cpg_ctx* ctx, cpg_stream stream, const float* gmm_mean, const float* gmm_sd, const float* gmm_cdf,
                     int K, int n_clf, const double* const* coef, const double* intercept, const int* target_col,
                     const int* f32, uint64_t seed, int64_t offset, int64_t n, float* z_out, double* probs,
                     double* accum, uint8_t* accept, int* comp_out, unsigned long long* n_accepted);
/* Re-generation of selected draws: z (and optionally the scores) of the draws whose global indices are listed in
 * draw_index[m] -- bit-identical to what cpg_class_sample(seed, ...) produced / would produce for those indices (every
 * draw is a pure function of (seed, index)).  With it a sampling round needs no z in HBM for the rejected draws:
 * cpg_class_sample(z_out = NULL) -> cpg_compact_accepted -> cpg_class_regen. */
int cpg_class_regen(cpg_ctx* ctx, cpg_stream stream, const float* gmm_mean, const float* gmm_sd, const float* gmm_cdf,
                    int K, int n_clf, const double* const* coef, const double* intercept, const int* target_col,
                    const int* f32, uint64_t seed, const int64_t* draw_index, int64_t m, float* z_out, double* probs,
                    double* accum);
/* ---- after the accept test (sample_pipeline.py:195-218,312-316) --------------------------------
 * cpg_compact_accepted: ascending indices (first_index + position) of the non-zero entries of accept[n] -> idx_out
 *   (at most `cap` are written), their number -> *count (device).  Stable: boolean-mask indexing of the reference.
 * cpg_gather_rows: dst[r][:] = src[idx[r] - index_base][:] for r < m (rows of D floats).
 * cpg_dedup_rows: rows int32 [n][width] (decoded token rows, -1 padded): first_index[i] = lowest row index with the
 *   same content, is_first[i] = (first_index[i] == i) -- pandas drop_duplicates() keeps exactly those rows.  Exact.
 * cpg_peptide_descriptors: per token row, over the residues only (aa_of_token[t] in 0..19, -1 = not a residue; HOST
 *   array of n_tokens entries): H = mean hydrophobicity, uH = |sum_i h_i exp(i * pos_i * angle)| / len (hydrophobic
 *   moment over the whole sequence: modlamp calculate_moment with window >= len), charge = charge_ends + sum of the
 *   side-chain partial charges, rounded to 3 decimals.  hydrophobicity20 / side_chain_charge20: HOST tables indexed by
 *   residue code.  length (residues per row) may be NULL. */
int cpg_compact_accepted(cpg_ctx* ctx, cpg_stream stream, const uint8_t* accept, int64_t n, int64_t first_index, int64_t cap,
                         int64_t* idx_out, unsigned long long* count);
int cpg_gather_rows(cpg_ctx* ctx, cpg_stream stream, const float* src, const int64_t* idx, int64_t index_base, int64_t m, int D,
                    float* dst);
int cpg_dedup_rows(cpg_ctx* ctx, cpg_stream stream, const int* rows, int64_t n, int width, int* first_index, uint8_t* is_first);
int cpg_peptide_descriptors(cpg_ctx* ctx, cpg_stream stream, const int* tokens, int64_t n, int width, const int8_t* aa_of_token,
                            int n_tokens, const float* hydrophobicity20, const double* side_chain_charge20, double charge_ends,
                            float angle_deg, float* H, float* uH, float* charge, int* length);
/* ---- data feed (data_processing/dataset.py:60-77,242-244,285-286) -----------------------------
 * The dataset lives on the device as padded token rows uint8 [n_examples][L] (Field(init_token, eos_token, fix_length)
 * layout) with the cumulative sampling weights cdf fp64 [n_examples].  One call draws a batch of the weighted random
 * iterator: B examples i.i.d. ~ weights with replacement (Philox(seed, step * B + row)), out_tokens int64 [B][L] in
 * `batch.text` form; out_index (chosen example per row) may be NULL. */
int cpg_feed_batch(cpg_ctx* ctx, cpg_stream stream, const uint8_t* tokens, const double* cdf, int64_t n_examples, int L,
                   uint64_t seed, uint64_t step, int B, int64_t* out_tokens, int64_t* out_index);
/* mogQ.logpdf batched (density_modeling.py:75-77): mean_t, prec_t fp64 [100][K] (component fastest),
 * logw_norm[k] = log w_k - 50 log(2 pi) + 0.5 sum_d log prec_kd; out fp64 [n]. */
int cpg_gmm_logpdf(cpg_ctx* ctx, cpg_stream stream, const float* x, int64_t n, const double* mean_t,
                   const double* prec_t, const double* logw_norm, int K, double* out);
/* prior_logpdf batched (density_modeling.py:11-14) */
int cpg_prior_logpdf(cpg_ctx* ctx, cpg_stream stream, const float* x, int64_t n, double* out);

/* ---- fitting Q(z) and the z-space classifiers on the device (density_modeling.py:64-73, sample_pipeline.py:169-192) ----
 * cpg_gmm_em_step: one EM iteration of a diagonal Gaussian mixture in fp64 (sklearn's arithmetic).  x fp32 [N][100];
 *   mean_in / prec_in fp64 [K][100] (prec = 1 / cov), logw_norm_in[k] = log w_k - 50 log(2 pi) + 0.5 sum_d log prec_kd;
 *   resp_ws: fp64 [N][K] scratch (responsibilities); outputs: weights_out[k] = n_k / N (un-normalised), mean_out, cov_out
 *   (incl. reg_covar), loglik_out fp64 [N] = log-likelihood of every point under the INPUT parameters.
 * cpg_logreg_newton_stats: for w fp64 [101] (100 coefficients + intercept) and labels y in {0,1} (fp32 [N]):
 *   out fp64 [cpg_logreg_stats_len()] = [sum_i log-loss | gradient (101) | Hessian upper triangle, row-major (5151)],
 *   WITHOUT the L2 penalty (the caller adds it). */
int cpg_gmm_em_step(cpg_ctx* ctx, cpg_stream stream, const float* x, int64_t N, int K, const double* mean_in, const double* prec_in,
                    const double* logw_norm_in, double reg_covar, double* resp_ws, double* weights_out, double* mean_out,
                    double* cov_out, double* loglik_out);
int cpg_logreg_newton_stats(cpg_ctx* ctx, cpg_stream stream, const float* x, const float* y01, int64_t N, const double* w101,
                            double* out);
int cpg_logreg_stats_len(void);

/* ---- options ------------------------------------------------------------------------------------
 * Which flavour of a kernel runs (the defaults pick the tcgen05 paths where the batch is large enough to
 * fill the chip; the fp32 SIMT kernels stay for small batches and are what the parity tests compare with):
 *   "mmd_tensor_core"     1 (default) persistent tcgen05 Gram kernel (tf32), 3 one tile per CTA, 0 fp32 SIMT
 *   "wgrad_tensor_core"   1 (default) tf32 tcgen05 weight / token-table gradients when B*L >= 8192, 2 always, 0 never
 *   "gru_tensor_core"     1 (default) split-bf16 tcgen05 recurrences (forward + BPTT) when B >= 512, 2 always, 0 never
 *   "dec_out_tensor_core" 1 (default) tcgen05 decoder-output layer when B*L >= 8192, 2 always, 0 never
 *   "bptt_fused"          1 (default) on the tcgen05 path the BPTT kernels also contract dW_hh and the token-table gradient
 *                         (the gate-gradient planes never reach HBM), 0 = separate tf32 weight-gradient kernels
 *   "latent_tensor_core"  1 (default) the dense layers around the latent code (heads, [z;c] projection and their backward) as two
 *                         fused split-bf16 tcgen05 kernels when B >= 512, 2 always, 0 fp32 SIMT GEMMs + element-wise kernels
 *   "latent_tile_rows"    64 (default) | 128: batch rows per CTA of the forward latent kernel
 *   "rf_tensor_core"      1 (default) random-feature map and its gradient (RF-MMD) as split-precision tcgen05 kernels when
 *                         B >= 512 (rf_dim a multiple of 4), 2 always, 0 fp32 SIMT GEMMs + element-wise kernels
 *   "wgrad_dense_tensor_core" 1 (default) head / [z;c]-projection weight gradients as split-bf16 tcgen05 batch contractions on
 *                         the reduction stream (where the tcgen05 latent layers are on), 0 head gradients inside the latent
 *                         backward kernel + dW_ih[:,150:] as an fp32 SIMT product
 *   "matmul_terms"        3 (default) every tensor-core contraction of the recurrences / decoder-output layer is the sum of the
 *                         three split-bf16 products (fp32-grade: the configuration all parity bars are stated for); 1 = the
 *                         leading bf16 product only ("bf16 matmul tiles", BASELINE.json configs[2]) -- a reduced-precision mode
 *   "adam_fused"          1 (default) sum of squares, norm, clip and Adam in ONE launch (grid-wide barrier), 0 two launches
 *   "chain_priority"      1 (default) the dependent chain of the fused step runs on a highest-priority internal stream
 *                         (forked from / joined to the caller's), 0 = on the caller's stream
 *   "cuda_graph"          1 (default) cpg_wae_train_step_philox replays a captured CUDA graph of the iteration, 0 = eager launches
 *   "side_stream"         1 (default) loss / weight-gradient kernels overlap the recurrences on an internal stream
 *                         (event fork/join inside each call; results identical), 0 everything on the caller's stream */
int cpg_set_option(const char* name, int value);

/* ---- per-kernel timing (CUDA events on the launching stream; off by default) ------------------ */
int cpg_profile_enable(int on);
/* Synchronises the device; fills names[i*name_stride..], total_ms[i], counts[i] per kernel label in first-launch
 * order and clears the records.  Returns the number of labels written (<= cap). */
int cpg_profile_read(char* names, int name_stride, float* total_ms, int* counts, int cap);
/* The launches seen by the last cpg_profile_read, unaggregated and in launch order: start offset from the first launch
 * and duration in ms (events on the launching streams, so overlap between the two streams is visible). */
int cpg_profile_timeline(char* names, int name_stride, float* start_ms, float* dur_ms, int cap);

#ifdef __cplusplus
}
#endif
#endif /* CPG_B200_H */
