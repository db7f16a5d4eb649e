// C ABI: context, workspace and the WAE training / inference entry points.
#include <string.h>
#include <math.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include <stdio.h>
#include "ctx.h"

namespace cpg {

long long g_launch_count = 0;
static thread_local std::string g_err;
void set_error(const std::string& m) { g_err = m; }

#ifdef CPG_EMU
static int dev_alloc(void** p, size_t n) { *p = calloc(1, n); return *p ? 0 : -1; }
static void dev_free(void* p) { free(p); }
static int dev_d2h(void* dst, const void* src, size_t n, cudaStream_t) { memcpy(dst, src, n); return 0; }
static int dev_sync(cudaStream_t) { return 0; }
static int dev_memset(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
static int dev_copy(void* d, const void* s_, size_t n, cudaStream_t) { memcpy(d, s_, n); return 0; }
static int dev_last_error(const char** msg) { *msg = ""; return 0; }
#else
static int dev_alloc(void** p, size_t n) { return cudaMalloc(p, n) == cudaSuccess ? 0 : -1; }
static void dev_free(void* p) { cudaFree(p); }
static int dev_d2h(void* dst, const void* src, size_t n, cudaStream_t s) {
    if (cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, s) != cudaSuccess) return -1;
    return cudaStreamSynchronize(s) == cudaSuccess ? 0 : -1;
}
static int dev_sync(cudaStream_t s) { return cudaStreamSynchronize(s) == cudaSuccess ? 0 : -1; }
static int dev_memset(void* p, int v, size_t n, cudaStream_t s) {
    return cudaMemsetAsync(p, v, n, s) == cudaSuccess ? 0 : -1;
}
static int dev_copy(void* d, const void* s_, size_t n, cudaStream_t s) {
    return cudaMemcpyAsync(d, s_, n, cudaMemcpyDeviceToDevice, s) == cudaSuccess ? 0 : -1;
}
static int dev_last_error(const char** msg) {
    cudaError_t e = cudaGetLastError();
    *msg = cudaGetErrorString(e);
    return e == cudaSuccess ? 0 : -1;
}
#endif

int check_launch(const char* where) {
    const char* msg;
    if (dev_last_error(&msg) != 0) {
        set_error(std::string(where) + ": CUDA error: " + msg);
        return CPG_ECUDA;
    }
    return CPG_OK;
}

struct Bump {
    char* base; size_t off;
    template <typename T> T* take(size_t n) {
        off = align_up(off, 256);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += n * sizeof(T);
        return p;
    }
};

static size_t layout_workspace(cpg_ctx* ctx, Workspace& w, int B, int L, int V, int R, char* base) {
    Bump a{base, 0};
    const size_t BL = (size_t)B * L;
    const size_t BL32 = (size_t)(ceil_div(B, 32) * 32) * L;     // gate stash of the tcgen05 recurrences: 32-row tiles
    const int sm = ctx->sm_count;
    w.B = B; w.L = L; w.V = V; w.R = R;
    // derived weights
    for (int d = 0; d < 2; ++d) w.d.t_enc[d] = a.take<float>((size_t)V * 3 * ENC_H);
    w.d.t_dec = a.take<float>((size_t)V * 3 * DEC_HP);
    for (int d = 0; d < 2; ++d) w.d.whh_t_enc[d] = a.take<float>((size_t)ENC_H * 3 * ENC_H);
    w.d.whh_t_dec = a.take<float>((size_t)DEC_HP * 3 * DEC_HP);
    w.d.whh_dec = a.take<float>((size_t)DEC_HP * 3 * DEC_HP);
    w.d.wizc_t = a.take<float>((size_t)DEC_HP * 3 * DEC_HP);
    w.d.wizc = a.take<float>((size_t)DEC_HP * 3 * DEC_HP);
    for (int d = 0; d < 2; ++d) w.d.bhn_enc[d] = a.take<float>(ENC_H);
    w.d.bhn_dec = a.take<float>(DEC_HP);
    w.d.fc_w = a.take<float>((size_t)VMAX * DEC_HP);
    w.d.fc_b = a.take<float>(VMAX);
#ifndef CPG_EMU
    w.d.lat_tiles = a.take<unsigned char>(LT_TILES_BYTES);
#else
    w.d.lat_tiles = nullptr;
#endif
    // tokens
    w.tok = a.take<uint8_t>(BL); w.tokd = a.take<uint8_t>(BL); w.tgt = a.take<uint8_t>(BL);
    // encoder
    for (int d = 0; d < 2; ++d) {
        w.enc_hs[d] = a.take<float>(BL * ENC_H);
        w.enc_gates[d] = a.take<float>(BL32 * 4 * ENC_H);
        w.enc_dg[d] = a.take<float>(BL * 4 * ENC_H);
    }
    w.hfin = a.take<float>((size_t)B * 2 * ENC_H);
    w.mu = a.take<float>((size_t)B * ZD); w.logvar = a.take<float>((size_t)B * ZD); w.z = a.take<float>((size_t)B * ZD);
    w.zc = a.take<float>((size_t)B * DEC_HP);
    w.rowbias = a.take<float>((size_t)B * 3 * DEC_HP);
    // decoder
    w.dec_hs = a.take<float>(BL * DEC_HP);
    w.dec_gates = a.take<float>(BL32 * 4 * DEC_HP);
    w.dec_dg = a.take<float>(BL * 4 * DEC_HP);
    w.dec_dh_out = a.take<float>(BL * DEC_HP);
    w.drow = a.take<float>((size_t)B * 3 * DEC_HP);
    w.dh0 = a.take<float>((size_t)B * DEC_HP);
    w.dmu = a.take<float>((size_t)B * ZD); w.dlv = a.take<float>((size_t)B * ZD);
    w.dhfin = a.take<float>((size_t)B * 2 * ENC_H);
    // random features
    w.rf_nchunk = std::max(1, std::min(B, 2 * sm));
    w.rf_pre1 = a.take<float>((size_t)B * R); w.rf_pre2 = a.take<float>((size_t)B * R);
    const size_t rf_rows = (size_t)std::max(w.rf_nchunk, rf_tc_parts(B));
    w.rf_part = a.take<float>(rf_rows * R);
    w.rf_part2 = a.take<float>(rf_rows * R);
    w.rf_tiles = a.take<unsigned char>(rf_tc_tile_bytes(R));
    w.rf_sum1 = a.take<float>(R); w.rf_sum2 = a.take<float>(R); w.rf_coef = a.take<float>(R);
    w.dz_rf = a.take<float>((size_t)B * ZD);
    w.lat_nparts = std::max(1, std::min(ceil_div(B, 8), 2 * sm));
    w.lat_part = a.take<float>((size_t)w.lat_nparts * 5); w.lat_sums = a.take<float>(8);
    w.mmd_ws = a.take<float>(mmd_ws_floats(B)); w.mmd_out = a.take<float>(4); w.mmdrf_out = a.take<float>(4);
    // decoder output partials
    int parts = dec_out_parts(B, L, sm);
    w.do_part_w = a.take<float>((size_t)parts * VMAX * DEC_HP);
    w.do_part_b = a.take<float>((size_t)parts * VMAX);
    w.do_part_nll = a.take<float>(parts);
    w.nll_sum = a.take<float>(4);
    // weight-gradient partials
    int ws_ = std::max(wgrad_splits(B, L, sm), wgrad_tc_splits(sm));
    ws_ = std::max(ws_, std::max(bptt_fused_ctas_enc(B), bptt_fused_ctas_dec(B)));   // fused BPTT: one partial per CTA
    w.wg_part = a.take<float>((size_t)ws_ * 3 * DEC_HP * DEC_HP);
    int ds_ = std::max(dtable_splits(B, L, sm), std::max(bptt_fused_ctas_enc(B), bptt_fused_ctas_dec(B)));
    w.dt_part = a.take<float>((size_t)ds_ * V * 4 * DEC_HP);
    w.wg_part_dec = a.take<float>((size_t)ws_ * 3 * DEC_HP * DEC_HP);     // own partials: runs concurrently with the encoder's
    w.dt_part_dec = a.take<float>((size_t)ds_ * V * 4 * DEC_HP);
    w.wg_part_enc1 = a.take<float>((size_t)ws_ * 3 * ENC_H * ENC_H);
    w.dt_part_enc1 = a.take<float>((size_t)ds_ * V * 4 * ENC_H);
    for (int d = 0; d < 2; ++d) w.dT_enc[d] = a.take<float>((size_t)V * 4 * ENC_H);
    w.dT_dec = a.take<float>((size_t)V * 4 * DEC_HP);
    w.dwizc = a.take<float>((size_t)3 * DEC_HP * DEC_HP);
    w.gemm_splits = std::max(1, std::min(ceil_div(B, 64), sm / 2));
    w.gemm_ws = a.take<float>((size_t)w.gemm_splits * 3 * DEC_HP * (2 * ENC_H));
    w.colsum_ws = a.take<float>((size_t)64 * 512);
    w.hg_part = a.take<float>((size_t)latent_bwd_tc_ctas(B) * LT_HG_ROWS * LT_HG_COLS);
    w.wd_part = a.take<float>(wgrad_dense_part_floats(B));
    w.emb_dec = a.take<float>((size_t)VMAX * EMB);
    w.norm_part = a.take<float>(2 * sm + 8);
    w.clip_coef = a.take<float>(4);
    w.scalars = a.take<float>(SC_COUNT);
    w.ntok_f = a.take<float>(4);
    w.coupled = a.take<float>(8 + 2 * (size_t)R);
    return align_up(a.off, 256);
}

int ensure_workspace(cpg_ctx* ctx, int B, int L, int V, int R, cudaStream_t stream) {
    Workspace& w = ctx->ws;
    if (w.B == B && w.L == L && w.V == V && w.R >= R && ctx->base != nullptr) return CPG_OK;
    if (w.R > R) R = w.R;                           // never shrink the random-feature scratch
    Workspace probe;
    size_t need = layout_workspace(ctx, probe, B, L, V, R, nullptr);
    if (need > ctx->capacity) {
        dev_sync(stream);
        if (ctx->base) dev_free(ctx->base);
        ctx->base = nullptr;
        ctx->capacity = 0;
        void* p = nullptr;
        if (dev_alloc(&p, need) != 0) {
            set_error("workspace allocation of " + std::to_string(need) + " bytes failed");
            return CPG_ENOMEM;
        }
        ctx->base = p;
        ctx->capacity = need;
    }
    layout_workspace(ctx, w, B, L, V, R, (char*)ctx->base);
    ctx->have_stash = false;
    return CPG_OK;
}

int ensure_aux(cpg_ctx* ctx, size_t bytes, cudaStream_t stream) {
    if (bytes <= ctx->aux_capacity && ctx->aux != nullptr) return CPG_OK;
    dev_sync(stream);
    if (ctx->aux) dev_free(ctx->aux);
    ctx->aux = nullptr;
    ctx->aux_capacity = 0;
    void* p = nullptr;
    bytes = align_up(bytes, 1 << 20);
    if (dev_alloc(&p, bytes) != 0) {
        set_error("scratch allocation of " + std::to_string(bytes) + " bytes failed");
        return CPG_ENOMEM;
    }
    ctx->aux = p;
    ctx->aux_capacity = bytes;
    return CPG_OK;
}

static int check_dims(int V, int B, int L) {
    if (V < 4 || V > VMAX) { set_error("n_vocab must be in [4, 32]"); return CPG_EINVAL; }
    if (B < 1) { set_error("batch must be >= 1"); return CPG_EINVAL; }
    if (L < 2 || L > LMAX) { set_error("seq_len must be in [2, 32]"); return CPG_EINVAL; }
    return CPG_OK;
}

// -------------------------------------------------------------------------------- internal streams
// g_opt_side_stream: 1 = the loss / reduction / weight-gradient kernels overlap the recurrences on two internal streams
// (default), 0 = everything on the caller's stream.  Lanes: m = the caller's stream (the dependent chain of the
// iteration), s = loss statistics, RF-MMD, full-kernel MMD, t = derived weight forms, ordered reductions of partials,
// weight-gradient products.  A dependency is an event recorded on the producer's lane (`mark`) and waited for on the
// consumer's (`wait_mark`); events come from a rotating pool (a wait binds to the record made before it was issued, so
// an event may be re-recorded once all waits on the earlier record have been ISSUED: far fewer than NEV per call).
int g_opt_side_stream = 1;
typedef void* Mark;
struct Lanes { cudaStream_t m, s, t; bool on; };
#ifndef CPG_EMU
constexpr int NEV = 32;
static bool side_ready(cpg_ctx* ctx) {
    if (!g_opt_side_stream) return false;
    if (ctx->side_stream == nullptr) {
        cudaStream_t q[3];
        cudaEvent_t e[NEV + 1];
        int least = 0, greatest = 0;
        cudaDeviceGetStreamPriorityRange(&least, &greatest);
        bool ok = cudaStreamCreateWithPriority(&q[0], cudaStreamNonBlocking, least) == cudaSuccess &&
                  cudaStreamCreateWithPriority(&q[1], cudaStreamNonBlocking, least) == cudaSuccess &&
                  cudaStreamCreateWithPriority(&q[2], cudaStreamNonBlocking, greatest) == cudaSuccess;
        for (int i = 0; i <= NEV && ok; ++i) ok = cudaEventCreateWithFlags(&e[i], cudaEventDisableTiming) == cudaSuccess;
        if (!ok) { cudaGetLastError(); return false; }
        ctx->side_stream = q[0];
        ctx->aux_stream = q[1];
        ctx->chain_stream = q[2];
        for (int i = 0; i < NEV; ++i) ctx->ev_pool[i] = e[i];
        ctx->ev_noise = e[NEV];
    }
    return true;
}
static Lanes lanes(cpg_ctx* ctx, cudaStream_t m) {
    if (!side_ready(ctx)) return Lanes{m, m, m, false};
    return Lanes{m, (cudaStream_t)ctx->side_stream, (cudaStream_t)ctx->aux_stream, true};
}
// everything enqueued on `s` so far
static Mark mark(cpg_ctx* ctx, const Lanes& l, cudaStream_t s) {
    if (!l.on) return nullptr;
    cudaEvent_t e = (cudaEvent_t)ctx->ev_pool[ctx->ev_next];
    ctx->ev_next = (ctx->ev_next + 1) % NEV;
    cudaEventRecord(e, s);
    return e;
}
static void wait_mark(cudaStream_t s, Mark m) { if (m != nullptr) cudaStreamWaitEvent(s, (cudaEvent_t)m, 0); }
// an event of the pool for a callee that records and waits by itself
static void* pool_event(cpg_ctx* ctx, const Lanes& l) {
    if (!l.on) return nullptr;
    void* e = ctx->ev_pool[ctx->ev_next];
    ctx->ev_next = (ctx->ev_next + 1) % NEV;
    return e;
}
// what is enqueued on `to` from now on runs after everything enqueued on `from` so far
static void order(cpg_ctx* ctx, const Lanes& l, cudaStream_t from, cudaStream_t to) {
    if (l.on && from != to) wait_mark(to, mark(ctx, l, from));
}
// noise produced on the side stream by cpg_fill_step_noise_overlapped: its first reader on another stream joins it
static void noise_join(cpg_ctx* ctx, cudaStream_t s) {
    if (ctx->noise_pending) { cudaStreamWaitEvent(s, (cudaEvent_t)ctx->ev_noise, 0); ctx->noise_pending = false; }
}
#else
static bool side_ready(cpg_ctx*) { return false; }
static Lanes lanes(cpg_ctx*, cudaStream_t m) { return Lanes{m, m, m, false}; }
static Mark mark(cpg_ctx*, const Lanes&, cudaStream_t) { return nullptr; }
static void wait_mark(cudaStream_t, Mark) {}
static void* pool_event(cpg_ctx*, const Lanes&) { return nullptr; }
static void order(cpg_ctx*, const Lanes&, cudaStream_t, cudaStream_t) {}
static void noise_join(cpg_ctx*, cudaStream_t) {}
#endif

// a noise request recorded by cpg_fill_step_noise_overlapped that no forward consumed: draw everything now, on `s`
static void flush_deferred_noise(cpg_ctx* ctx, cudaStream_t s) {
    if (!ctx->gen_deferred) return;
    ctx->gen_deferred = false;
    launch_step_noise(s, ctx->gen_args, NOISE_ALL);
}

// ------------------------------------------------------------------------------------ forward
// Tensor-core recurrences pay off once a 128-row tile per CTA fills a good part of the chip;
// tiny batches stay on the 32-row fp32 SIMT kernels.  g_opt_gru_tc: 0 = never, 1 = auto, 2 = always.
int g_opt_gru_tc = 1;
#ifndef CPG_EMU
static bool use_gru_tc(int B) { return g_opt_gru_tc == 2 || (g_opt_gru_tc == 1 && B >= 512); }
#else
static bool use_gru_tc(int) { return false; }
#endif
// g_opt_bptt_fused: 1 (default) = on the tcgen05 path the BPTT kernels also contract dW_hh / the token-table
// gradient (no dg planes in HBM, no separate weight-gradient kernel); 0 = k_gru_bwd_tc + k_wgrad_tc
int g_opt_bptt_fused = 1;
// Returns the mark "mu, logvar, z are final" (made on the caller's lane right after the latent layers) for the loss
// statistics of the training step; null when the lanes are off.
static Mark forward_impl(cpg_ctx* ctx, const Lanes& ln, const float* params, const ParamLayout& lay, int V, int B, int L,
                         const cpg_wae_inputs* in, float* mu, float* logvar, float* z, bool stash, bool encoder_only,
                         const StepNoiseArgs* gen = nullptr, float* ntok_out = nullptr) {
    Workspace& w = ctx->ws;
    cudaStream_t s = ln.m;
    StepNoiseArgs deferred;
    if (ctx->gen_deferred) {
        // the request of cpg_fill_step_noise_overlapped is for this forward when it names the buffers this forward reads
        if (gen == nullptr && !encoder_only && ctx->gen_args.B == B && ctx->gen_args.L == L && ctx->gen_args.word_drop != nullptr &&
            ctx->gen_args.word_drop == in->word_drop) {
            deferred = ctx->gen_args;
            gen = &deferred;
            ctx->gen_deferred = false;
        } else {
            flush_deferred_noise(ctx, s);
        }
    }
    // the derived weight forms only depend on the parameters: their lane runs beside the token preparation
    order(ctx, ln, s, ln.t);
    Mark weights_ready = nullptr, rest_ready = nullptr;
    if (ln.on) {
        launch_prep_weights(ln.t, params, lay, V, w.d, 1);      // what the encoder recurrence reads ...
        weights_ready = mark(ctx, ln, ln.t);
        launch_prep_weights(ln.t, params, lay, V, w.d, 2);      // ... the rest lands under it
        rest_ready = mark(ctx, ln, ln.t);
    } else {
        launch_prep_weights(s, params, lay, V, w.d, 3);
    }
    // gen: this step's noise is drawn here -- the word-dropout mask inside the token preparation, the rest on lane s
    // once the preparation is through (its first reader, the latent layers, joins it)
    launch_prep_tokens(s, in->tokens, in->word_drop, B, L, V, w.tok, w.tokd, w.tgt, ctx->ints, ctx->ints + 1, gen);
    if (gen != nullptr) {
        if (ln.on) {
            order(ctx, ln, s, ln.s);                // also: earlier readers of the noise buffers (previous iteration) are on `s`
            launch_step_noise(ln.s, *gen, NOISE_LATENT | NOISE_LATE);
#ifndef CPG_EMU
            cudaEventRecord((cudaEvent_t)ctx->ev_noise, ln.s);
#endif
            ctx->noise_pending = true;
        } else {
            launch_step_noise(s, *gen, NOISE_LATENT | NOISE_LATE);
        }
    }
    if (ntok_out != nullptr) {                      // data-parallel callers exchange the token count early, on lane t
        order(ctx, ln, s, ln.t);
        launch_int_to_float(ln.t, ctx->ints, ntok_out, 1);
    }
    wait_mark(s, weights_ready);
    GruSeq enc[2];
    for (int d = 0; d < 2; ++d) {
        GruSeq& q = enc[d];
        memset(&q, 0, sizeof(q));
        q.tok = w.tok; q.table = w.d.t_enc[d]; q.whh_t = w.d.whh_t_enc[d]; q.bhn = w.d.bhn_enc[d];
        q.hs = stash ? w.enc_hs[d] : nullptr;
        q.gates = stash ? w.enc_gates[d] : nullptr;
        q.hfin = w.hfin + d * ENC_H; q.hfin_stride = 2 * ENC_H;
        q.reverse = d;
    }
    if (use_gru_tc(B)) {
        enc[0].whh = params + lay.off[P_ENC_WHH_F];
        enc[1].whh = params + lay.off[P_ENC_WHH_R];
        launch_gru_fwd_enc_tc(s, enc, B, L, V);
    } else {
        launch_gru_fwd_enc(s, enc, B, L);
    }
    wait_mark(s, rest_ready);
    noise_join(ctx, s);                             // eps, c (and the out-dropout mask) of cpg_fill_step_noise_overlapped
    if (latent_uses_tc(B)) {
        // heads -> reparameterisation -> [z;c] -> its input projection, one tcgen05 kernel (latent_tc.cu)
        launch_latent_fwd_tc(s, w.hfin, params + lay.off[P_QMU_B], params + lay.off[P_QLV_B], encoder_only ? nullptr : in->eps,
                             encoder_only ? nullptr : in->c, w.d.lat_tiles, B, mu, logvar, encoder_only ? nullptr : z, w.zc,
                             encoder_only ? nullptr : w.rowbias);
        if (encoder_only) return nullptr;
    } else {
        // q_mu / q_logvar heads (models/encoder.py:50-51)
        launch_sgemm_pair(s, B, ZD, 2 * ENC_H, 1.f, w.hfin, 2 * ENC_H, 1, params + lay.off[P_QMU_W], params + lay.off[P_QLV_W],
                          1, 2 * ENC_H, mu, logvar, ZD, params + lay.off[P_QMU_B], params + lay.off[P_QLV_B]);
        if (encoder_only) return nullptr;
        launch_reparam(s, mu, logvar, in->eps, in->c, B, z, w.zc);
    }
    const Mark latent_ready = mark(ctx, ln, s);     // mu, logvar, z are final: the loss statistics may start
    if (!latent_uses_tc(B)) {
        // per-row input projection of [z;c] (the non-embedding columns of decoder W_ih)
        launch_sgemm(s, B, 3 * DEC_HP, DEC_HP, 1.f, w.zc, DEC_HP, 1, w.d.wizc_t, 3 * DEC_HP, 1, 0.f, w.rowbias,
                     3 * DEC_HP, nullptr, 1, nullptr);
    }
    GruSeq q;
    memset(&q, 0, sizeof(q));
    q.tok = w.tokd; q.table = w.d.t_dec; q.rowbias = w.rowbias; q.whh_t = w.d.whh_t_dec; q.bhn = w.d.bhn_dec;
    q.h0 = w.zc; q.hs = w.dec_hs; q.gates = stash ? w.dec_gates : nullptr;
    if (use_gru_tc(B)) {
        q.whh = w.d.whh_dec;
        launch_gru_fwd_dec_tc(s, q, B, L, V);
    } else {
        launch_gru_fwd_dec(s, q, B, L);
    }
    return latent_ready;
}

static DecOutArgs dec_out_args(cpg_ctx* ctx, const cpg_wae_inputs* in, int V, int B, int L) {
    Workspace& w = ctx->ws;
    DecOutArgs a;
    memset(&a, 0, sizeof(a));
    a.hs = w.dec_hs;
    a.out_keep = in->out_keep;
    a.keep_scale = 1.0f / (1.0f - in->p_out_dropout);
    a.fc_w = w.d.fc_w; a.fc_b = w.d.fc_b; a.tgt = w.tgt;
    a.part_w = w.do_part_w; a.part_b = w.do_part_b; a.part_nll = w.do_part_nll;
    a.B = B; a.L = L; a.V = V;
    a.nprod = g_opt_matmul_terms == 1 ? 1 : 3;
    return a;
}

// ----------------------------------------------------------------------------------- backward
// Everything after the decoder-output layer has produced dec_dh_out.  dz_rf / external latent gradients are optional;
// `dz_ready` = the mark after which lat_in.dz_rf may be read (null: it is already ordered before the caller's lane).
// On return every gradient is complete on the caller's lane (the other lanes are joined).
static void backward_impl(cpg_ctx* ctx, const Lanes& ln, const float* params, const ParamLayout& lay, float* grads,
                          int V, int B, int L, const cpg_wae_inputs* in, const LatentBwdArgs& lat_in, Mark dz_ready) {
    Workspace& w = ctx->ws;
    const int sm = ctx->sm_count;
    cudaStream_t s = ln.m;
    InputGradArgs ia;
    memset(&ia, 0, sizeof(ia));
    ia.emb = params + lay.off[P_EMB];
    ia.enc_wih[0] = params + lay.off[P_ENC_WIH_F]; ia.enc_wih[1] = params + lay.off[P_ENC_WIH_R];
    ia.dec_wih = params + lay.off[P_DEC_WIH];
    ia.dT_enc[0] = w.dT_enc[0]; ia.dT_enc[1] = w.dT_enc[1]; ia.dT_dec = w.dT_dec; ia.dwizc = w.dwizc;
    ia.g_emb = grads + lay.off[P_EMB];
    ia.emb_dec = w.emb_dec;
    ia.g_enc_wih[0] = grads + lay.off[P_ENC_WIH_F]; ia.g_enc_bih[0] = grads + lay.off[P_ENC_BIH_F];
    ia.g_enc_bhh[0] = grads + lay.off[P_ENC_BHH_F];
    ia.g_enc_wih[1] = grads + lay.off[P_ENC_WIH_R]; ia.g_enc_bih[1] = grads + lay.off[P_ENC_BIH_R];
    ia.g_enc_bhh[1] = grads + lay.off[P_ENC_BHH_R];
    ia.g_dec_wih = grads + lay.off[P_DEC_WIH]; ia.g_dec_bih = grads + lay.off[P_DEC_BIH];
    ia.g_dec_bhh = grads + lay.off[P_DEC_BHH];
    ia.V = V;
    // decoder BPTT
    GruSeq q;
    memset(&q, 0, sizeof(q));
    q.whh = w.d.whh_dec; q.h0 = w.zc; q.hs = w.dec_hs; q.gates = w.dec_gates;
    q.dh_out = w.dec_dh_out; q.dg = w.dec_dg; q.dh0 = w.dh0; q.drow = w.drow;
    // the tcgen05 BPTT kernels hand dg to the tf32 weight-gradient kernel already rounded
    const int dg_rounded = (use_gru_tc(B) && wgrad_uses_tc(B * L)) ? 1 : 0;
    const bool fused = use_gru_tc(B) && g_opt_bptt_fused != 0;
    if (fused) launch_gru_bwd_dec_fused(s, q, w.tokd, B, L, V, w.wg_part_dec, w.dt_part_dec);
    else if (use_gru_tc(B)) launch_gru_bwd_dec_tc(s, q, B, L, dg_rounded);
    else launch_gru_bwd_dec(s, q, B, L);
    // decoder W_hh / token-table gradients only need the decoder BPTT's output: lane t, under the dense layers and the
    // encoder BPTT of the caller's lane (joined before the input-side gradients)
    {
        order(ctx, ln, s, ln.t);
        // dW_ih[:,150:] = drow^T @ [z;c]  (a weight gradient: nothing on the BPTT chain waits for it)
        const bool wd_tc = latent_uses_tc(B) && g_opt_wgrad_dense_tc != 0;
        if (wd_tc) launch_wgrad_zc_tc(ln.t, w.drow, w.zc, B, w.wd_part, w.dwizc);
        else launch_sgemm(ln.t, 3 * DEC_HP, DEC_HP, B, 1.f, w.drow, 1, 3 * DEC_HP, w.zc, DEC_HP, 1, 0.f, w.dwizc, DEC_HP,
                          nullptr, w.gemm_splits, w.gemm_ws);
    }
    const float* wmu = params + lay.off[P_QMU_W];
    const float* wlv = params + lay.off[P_QLV_W];
    LatentBwdArgs la = lat_in;
    la.mu = w.mu; la.logvar = w.logvar; la.eps = in->eps; la.dzc = w.dh0; la.B = B;
    la.dmu = w.dmu; la.dlv = w.dlv;
    if (latent_uses_tc(B)) {
        // gradient at [z;c] -> latent backward -> gradient at the encoder's final hidden state, one tcgen05 kernel
        wait_mark(s, dz_ready);                     // dz_rf produced on lane s
        const bool wd_tc = g_opt_wgrad_dense_tc != 0;
        launch_latent_bwd_tc(s, w.drow, w.dh0, w.d.lat_tiles, la, w.dhfin, w.hfin, wd_tc ? nullptr : w.hg_part);
        // head weight / bias gradients on lane t: a batch contraction of their own, or (option off) the ordered sum of the
        // partials the kernel above contracted over its rows
        order(ctx, ln, s, ln.t);
        if (wd_tc) launch_wgrad_heads_tc(ln.t, w.dmu, w.dlv, w.hfin, B, w.wd_part, grads + lay.off[P_QMU_W], grads + lay.off[P_QLV_W],
                                         grads + lay.off[P_QMU_B], grads + lay.off[P_QLV_B]);
        else launch_head_grad_reduce(ln.t, w.hg_part, B, grads + lay.off[P_QMU_W], grads + lay.off[P_QLV_W], grads + lay.off[P_QMU_B],
                                     grads + lay.off[P_QLV_B]);
    } else {
        // gradient at [z;c]:  dh0 + drow @ W_ih[:,150:]   (in place on dh0)
        launch_sgemm(s, B, DEC_HP, 3 * DEC_HP, 1.f, w.drow, 3 * DEC_HP, 1, w.d.wizc, DEC_HP, 1, 1.f, w.dh0, DEC_HP,
                     nullptr, 1, nullptr);
        wait_mark(s, dz_ready);
        launch_latent_bwd(s, la);
        // heads
        launch_sgemm_sum2(s, B, 2 * ENC_H, ZD, 1.f, w.dmu, w.dlv, ZD, 1, wmu, wlv, 2 * ENC_H, 1, 0.f, w.dhfin, 2 * ENC_H);
        // head weight / bias gradients: lane t (after the decoder weight gradients: same split-K scratch), under the encoder BPTT
        order(ctx, ln, s, ln.t);
        launch_sgemm(ln.t, ZD, 2 * ENC_H, B, 1.f, w.dmu, 1, ZD, w.hfin, 2 * ENC_H, 1, 0.f, grads + lay.off[P_QMU_W],
                     2 * ENC_H, nullptr, w.gemm_splits, w.gemm_ws);
        launch_sgemm(ln.t, ZD, 2 * ENC_H, B, 1.f, w.dlv, 1, ZD, w.hfin, 2 * ENC_H, 1, 0.f, grads + lay.off[P_QLV_W],
                     2 * ENC_H, nullptr, w.gemm_splits, w.gemm_ws);
        launch_colsum(ln.t, w.dmu, B, ZD, ZD, grads + lay.off[P_QMU_B], w.colsum_ws, 64);
        launch_colsum(ln.t, w.dlv, B, ZD, ZD, grads + lay.off[P_QLV_B], w.colsum_ws, 64);
    }
    // decoder W_hh / token-table sums and the decoder's input-side gradients: lane t, behind the head gradients (nothing waits
    // for them before the end of the step, while the head gradients' operands are only complete now; the sums ahead of the
    // dense weight gradients, under the latent backward, measured no different in the captured graph)
    {
        if (fused) {
            launch_wgrad_partial_reduce(ln.t, DEC_HP, DEC_H, V, w.wg_part_dec, w.dt_part_dec, bptt_fused_ctas_dec(B),
                                        grads + lay.off[P_DEC_WHH], w.dT_dec);
        } else {
            const bool t2 = launch_wgrad_hh(ln.t, DEC_HP, DEC_H, w.dec_dg, w.dec_hs, w.zc, w.tokd, 0, V, B, L, sm, w.wg_part_dec,
                                            w.dt_part_dec, grads + lay.off[P_DEC_WHH], w.dT_dec, nullptr, nullptr, dg_rounded);
            if (!t2) launch_dtable(ln.t, DEC_HP, w.dec_dg, w.tokd, B, L, 0, V, sm, w.dt_part_dec, w.dT_dec);
        }
        launch_input_grads(ln.t, ia, nullptr, 2 | 8);   // decoder W_ih / bias gradients, its share of the embedding gradient
    }
    // encoder BPTT
    GruSeq enc[2];
    for (int d = 0; d < 2; ++d) {
        GruSeq& e = enc[d];
        memset(&e, 0, sizeof(e));
        e.whh = params + lay.off[d == 0 ? P_ENC_WHH_F : P_ENC_WHH_R];
        e.hs = w.enc_hs[d]; e.gates = w.enc_gates[d];
        e.dh_fin = w.dhfin + d * ENC_H; e.dh_fin_stride = 2 * ENC_H;
        e.dg = w.enc_dg[d];
    }
    if (fused) {
        float* const pw[2] = {w.wg_part, w.wg_part_enc1};
        float* const pt[2] = {w.dt_part, w.dt_part_enc1};
        launch_gru_bwd_enc_fused(s, enc, w.tok, B, L, V, pw, pt);
        // ordered reductions of the per-CTA partials: one direction on lane t, the other on the caller's lane
        const int nc = bptt_fused_ctas_enc(B);
        order(ctx, ln, s, ln.t);
        launch_wgrad_partial_reduce(ln.t, ENC_H, ENC_H, V, pw[0], pt[0], nc, grads + lay.off[P_ENC_WHH_F], w.dT_enc[0]);
        launch_wgrad_partial_reduce(s, ENC_H, ENC_H, V, pw[1], pt[1], nc, grads + lay.off[P_ENC_WHH_R], w.dT_enc[1]);
    } else {
        if (use_gru_tc(B)) launch_gru_bwd_enc_tc(s, enc, B, L, dg_rounded);
        else launch_gru_bwd_enc(s, enc, B, L);
        // recurrent weight gradients (the tensor-core path produces the token-table gradient in the same pass over dg);
        // each direction has its own partials, and the ordered reductions of the partials run on lane s while the caller's
        // lane already contracts the next direction
        cudaStream_t rs = ln.on ? ln.s : nullptr;
        const bool t0 = launch_wgrad_hh(s, ENC_H, ENC_H, w.enc_dg[0], w.enc_hs[0], nullptr, w.tok, 0, V, B, L, sm, w.wg_part,
                                        w.dt_part, grads + lay.off[P_ENC_WHH_F], w.dT_enc[0], rs, pool_event(ctx, ln), dg_rounded);
        if (!t0) launch_dtable(s, ENC_H, w.enc_dg[0], w.tok, B, L, 0, V, sm, w.dt_part, w.dT_enc[0]);
        const bool t1 = launch_wgrad_hh(s, ENC_H, ENC_H, w.enc_dg[1], w.enc_hs[1], nullptr, w.tok, 1, V, B, L, sm, w.wg_part_enc1,
                                        w.dt_part_enc1, grads + lay.off[P_ENC_WHH_R], w.dT_enc[1], rs, pool_event(ctx, ln), dg_rounded);
        if (!t1) launch_dtable(s, ENC_H, w.enc_dg[1], w.tok, B, L, 1, V, sm, w.dt_part_enc1, w.dT_enc[1]);
    }
    // encoder input-side gradients on the caller's lane, the embedding gradient (all three token-table gradients) beside it
    order(ctx, ln, ln.t, s);                        // lane t: the other direction's reduction, decoder-side gradients
    if (ln.on) {
        order(ctx, ln, s, ln.t);
        launch_input_grads(s, ia, ln.t, 1 | 4);
        order(ctx, ln, ln.t, s);
    } else {
        launch_input_grads(s, ia, nullptr, 1 | 4);
    }
    order(ctx, ln, ln.s, s);                        // lane s: loss scalars (and, in the non-fused BPTT path, a reduction)
}

static AdamHyper adam_hyper(const cpg_train_hparams* hp) {
    AdamHyper h;
    h.beta1 = hp->beta1; h.beta2 = hp->beta2; h.eps = hp->adam_eps;
    auto mk = [&](int step) {
        AdamStep st;
        double bc1 = 1.0 - pow((double)hp->beta1, (double)step);
        double bc2 = 1.0 - pow((double)hp->beta2, (double)step);
        st.step_size = (float)((double)hp->lr / bc1);
        st.bc2_sqrt = (float)sqrt(bc2);
        return st;
    };
    h.single = mk(hp->adam_step);
    h.dup_first = mk(2 * hp->adam_step - 1);
    h.dup_second = mk(2 * hp->adam_step);
    return h;
}

// ------------------------------------------------------------------------------- captured iteration
const StepDyn* g_dyn = nullptr;
int g_opt_chain_priority = 1;         // 1: the fused step's dependent chain runs on the highest-priority internal stream
int g_opt_matmul_terms = 3;           // 3: split-bf16 products (parity configuration), 1: single bf16 product (reduced precision)
int g_opt_adam_fused = 1;             // 1: sum of squares + norm + clip + Adam in one launch (grid barrier), 0: two launches
int g_opt_graph = 1;                  // 1: cpg_wae_train_step_philox replays a captured CUDA graph of the iteration
__global__ void k_set_dyn(StepDyn v, StepDyn* __restrict__ dst) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *dst = v;
}

static StepDyn make_dyn(const cpg_train_hparams* hp, uint32_t noise_step) {
    const AdamHyper h = adam_hyper(hp);
    StepDyn d;
    d.beta = hp->beta;
    d.step_size[0] = h.single.step_size; d.bc2_sqrt[0] = h.single.bc2_sqrt;
    d.step_size[1] = h.dup_first.step_size; d.bc2_sqrt[1] = h.dup_first.bc2_sqrt;
    d.step_size[2] = h.dup_second.step_size; d.bc2_sqrt[2] = h.dup_second.bc2_sqrt;
    d.noise_step = noise_step;
    return d;
}

#ifndef CPG_EMU
struct StepGraph {
    std::string key;
    int seen = 0;                     // eager runs of this key so far (the first one sizes every workspace)
    cudaGraphExec_t exec = nullptr;
    cudaGraphNode_t dyn_node = nullptr;
    long long launches = 0;           // kernels inside the graph (for cpg_launch_count)
    long long last_use = 0;
};
constexpr int MAX_GRAPHS = 8;
static StepGraph g_graphs[MAX_GRAPHS];
static long long g_graph_clock = 0;

static StepGraph* find_graph(const std::string& key) {
    StepGraph* lru = &g_graphs[0];
    for (auto& g : g_graphs) {
        if (g.key == key) { g.last_use = ++g_graph_clock; return &g; }
        if (g.last_use < lru->last_use) lru = &g;
    }
    if (lru->exec) { cudaGraphExecDestroy(lru->exec); }
    *lru = StepGraph();
    lru->key = key;
    lru->last_use = ++g_graph_clock;
    return lru;
}
#endif

}  // namespace cpg

using namespace cpg;

extern "C" {

int cpg_abi_version(void) { return CPG_ABI_VERSION; }
const char* cpg_last_error(void) { return g_err.c_str(); }

int cpg_create(cpg_ctx** out, int device) {
    if (out == nullptr) { set_error("cpg_create: out is null"); return CPG_EINVAL; }
    cpg_ctx* c = new cpg_ctx();
    c->device = device;
#ifndef CPG_EMU
    if (cudaSetDevice(device) != cudaSuccess) {
        set_error("cpg_create: cudaSetDevice failed (is a CUDA device present?)");
        delete c;
        return CPG_ECUDA;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        set_error("cpg_create: cudaGetDeviceProperties failed");
        delete c;
        return CPG_ECUDA;
    }
    if (prop.major < 10) {
        set_error(std::string("cpg_create: device '") + prop.name + "' is not sm_100-class; this library is built for sm_100a only");
        delete c;
        return CPG_ECUDA;
    }
    c->sm_count = prop.multiProcessorCount;
    g_sm_count = c->sm_count;
#endif
    void* p = nullptr;
    if (dev_alloc(&p, 256) != 0) { set_error("cpg_create: allocation failed"); delete c; return CPG_ENOMEM; }
    c->ints = (int*)p;                         // [0..15] ints, [32..] StepDyn of the captured iteration
    dev_memset(c->ints, 0, 256, nullptr);
    dev_sync(nullptr);
    *out = c;
    return CPG_OK;
}

int cpg_destroy(cpg_ctx* c) {
    if (c == nullptr) return CPG_OK;
    if (c->base) dev_free(c->base);
    if (c->aux) dev_free(c->aux);
    if (c->ints) dev_free(c->ints);
#ifndef CPG_EMU
    if (c->side_stream) {
        cudaStreamSynchronize((cudaStream_t)c->side_stream);
        cudaStreamSynchronize((cudaStream_t)c->aux_stream);
        cudaStreamSynchronize((cudaStream_t)c->chain_stream);
        cudaStreamDestroy((cudaStream_t)c->side_stream);
        cudaStreamDestroy((cudaStream_t)c->aux_stream);
        cudaStreamDestroy((cudaStream_t)c->chain_stream);
        for (int k = 0; k < NEV; ++k) cudaEventDestroy((cudaEvent_t)c->ev_pool[k]);
        cudaEventDestroy((cudaEvent_t)c->ev_noise);
    }
#endif
    delete c;
    return CPG_OK;
}

int cpg_sm_count(const cpg_ctx* c) { return c ? c->sm_count : 0; }
int64_t cpg_workspace_bytes(const cpg_ctx* c) { return c ? (int64_t)c->capacity : 0; }
int64_t cpg_launch_count(const cpg_ctx*) { return (int64_t)g_launch_count; }
int64_t cpg_stash_generation(const cpg_ctx* c) { return (c && c->have_stash) ? c->stash_gen : -1; }

int cpg_check_errors(cpg_ctx* c, cpg_stream stream) {
    int host[2] = {0, 0};
    if (dev_d2h(host, c->ints, sizeof(host), (cudaStream_t)stream) != 0) { set_error("cpg_check_errors: copy failed"); return CPG_ECUDA; }
    if (host[1] != 0) { set_error("token id outside [0, n_vocab) in an input batch"); return CPG_ETOKEN; }
    return check_launch("cpg_check_errors");
}

int64_t cpg_vae_param_count(int V) { return make_layout(V).total; }
int cpg_vae_param_layout(int V, int64_t offsets[CPG_N_PARAM_TENSORS], int64_t sizes[CPG_N_PARAM_TENSORS]) {
    if (V < 4 || V > VMAX) { set_error("n_vocab must be in [4, 32]"); return CPG_EINVAL; }
    ParamLayout l = make_layout(V);
    for (int i = 0; i < P_COUNT; ++i) { offsets[i] = l.off[i]; sizes[i] = l.size[i]; }
    return CPG_OK;
}

int cpg_wae_encode(cpg_ctx* ctx, cpg_stream stream, const float* params, int V, int B, int L, const int64_t* tokens,
                   float* mu, float* logvar) {
    int rc = check_dims(V, B, L);
    if (rc) return rc;
    if (!ctx || !params || !tokens || !mu || !logvar) { set_error("cpg_wae_encode: null argument"); return CPG_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    if ((rc = ensure_workspace(ctx, B, L, V, ctx->ws.R > 0 ? ctx->ws.R : 500, s))) return rc;
    cpg_wae_inputs in;
    memset(&in, 0, sizeof(in));
    in.tokens = tokens;
    const Lanes ln = lanes(ctx, s);
    forward_impl(ctx, ln, params, make_layout(V), V, B, L, &in, mu, logvar, nullptr, false, true);
    ctx->have_stash = false;
    return check_launch("cpg_wae_encode");
}

int cpg_wae_forward(cpg_ctx* ctx, cpg_stream stream, const float* params, int V, int B, int L, const cpg_wae_inputs* in,
                    float* mu, float* logvar, float* z, float* logits, int keep) {
    int rc = check_dims(V, B, L);
    if (rc) return rc;
    if (!ctx || !params || !in || !in->tokens || !in->c) { set_error("cpg_wae_forward: null argument"); return CPG_EINVAL; }
    if (in->out_keep && !(in->p_out_dropout >= 0.f && in->p_out_dropout < 1.f)) { set_error("p_out_dropout must be in [0,1)"); return CPG_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    if ((rc = ensure_workspace(ctx, B, L, V, ctx->ws.R > 0 ? ctx->ws.R : 500, s))) return rc;
    Workspace& w = ctx->ws;
    ParamLayout lay = make_layout(V);
    const Lanes ln = lanes(ctx, s);
    noise_join(ctx, s);                             // noise of cpg_fill_step_noise_overlapped, if any
    forward_impl(ctx, ln, params, lay, V, B, L, in, w.mu, w.logvar, w.z, keep != 0, false);
    if (mu) dev_copy(mu, w.mu, (size_t)B * ZD * 4, s);
    if (logvar) dev_copy(logvar, w.logvar, (size_t)B * ZD * 4, s);
    if (z) dev_copy(z, w.z, (size_t)B * ZD * 4, s);
    if (logits) {
        DecOutArgs a = dec_out_args(ctx, in, V, B, L);
        a.logits_out = logits;
        a.part_nll = nullptr;
        launch_dec_out(s, a, ctx->sm_count);
    }
    ctx->have_stash = keep != 0;
    if (keep) ctx->stash_gen++;
    return check_launch("cpg_wae_forward");
}

int cpg_wae_backward(cpg_ctx* ctx, cpg_stream stream, const float* params, int V, int B, int L, const cpg_wae_inputs* in,
                     const float* d_mu, const float* d_logvar, const float* d_z, const float* d_logits, float* grads) {
    int rc = check_dims(V, B, L);
    if (rc) return rc;
    if (!ctx || !params || !in || !grads) { set_error("cpg_wae_backward: null argument"); return CPG_EINVAL; }
    Workspace& w = ctx->ws;
    if (!ctx->have_stash || w.B != B || w.L != L || w.V != V) {
        set_error("cpg_wae_backward: no matching cpg_wae_forward(keep_for_backward=1) stash");
        return CPG_EINVAL;
    }
    cudaStream_t s = (cudaStream_t)stream;
    flush_deferred_noise(ctx, s);
    noise_join(ctx, s);
    const Lanes ln = lanes(ctx, s);
    ParamLayout lay = make_layout(V);
    dev_memset(grads, 0, (size_t)lay.total * 4, s);
    DecOutArgs a = dec_out_args(ctx, in, V, B, L);
    a.dlogits_in = d_logits;           // null -> zero gradient through the logits
    a.dh_out = w.dec_dh_out;
    a.part_nll = nullptr;
    launch_dec_out(s, a, ctx->sm_count);
    launch_dec_out_reduce(s, a, ctx->sm_count, grads + lay.off[P_FC_W], grads + lay.off[P_FC_B], nullptr);
    LatentBwdArgs la;
    memset(&la, 0, sizeof(la));
    la.dz_ext = d_z; la.dmu_ext = d_mu; la.dlv_ext = d_logvar;
    la.B_global = B;
    backward_impl(ctx, ln, params, lay, grads, V, B, L, in, la, nullptr);
    return check_launch("cpg_wae_backward");
}

int cpg_fill_step_noise_overlapped(cpg_ctx* ctx, cpg_stream stream, uint64_t seed, uint32_t step, int B, int L, float p_word,
                                   float p_out, float* eps, float* c, uint8_t* word_drop, uint8_t* out_keep,
                                   float* z_prior_full, float* z_prior_rf) {
    if (!ctx || B < 1 || L < 1) { set_error("cpg_fill_step_noise_overlapped: bad argument"); return CPG_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    StepNoiseArgs a;
    a.seed = seed; a.step = step; a.B = B; a.L = L; a.p_word = p_word; a.p_out = p_out;
    a.eps = eps; a.c = c; a.word_drop = word_drop; a.out_keep = out_keep; a.zp_full = z_prior_full; a.zp_rf = z_prior_rf;
    const Lanes ln = lanes(ctx, s);
    if (!ln.on) {
        launch_step_noise(s, a, NOISE_ALL);
        return check_launch("cpg_fill_step_noise_overlapped");
    }
    // Recorded, not launched: the forward that reads these buffers next (cpg_wae_step_phase1, cpg_wae_train_step,
    // cpg_wae_forward) draws the word-dropout mask inside its token preparation and the rest on the loss lane once that
    // preparation is through -- drawn right here, the big kernel would slow the preparation kernels the encoder waits for.
    // Any other entry point that may read them draws everything first (flush_deferred_noise).
    noise_join(ctx, s);                             // a previous batch of noise nobody consumed yet
    flush_deferred_noise(ctx, s);
    ctx->gen_args = a;
    ctx->gen_deferred = true;
    return check_launch("cpg_fill_step_noise_overlapped");
}

int64_t cpg_coupled_count(int rf_dim) { return 8 + 2 * (int64_t)rf_dim; }

// Phase 1.  `join` = the public cpg_wae_step_phase1: every coupled statistic is produced here -- coupled[0] (token count)
// on lane t right after the token preparation, the rest on lane s -- and NOT joined into the caller's stream: the caller
// exchanges them on those lanes (cpg_aux_stream / cpg_side_stream) and phase 2 orders its consumers behind them.  The fused
// single-GPU step (join = false) keeps the log-only statistics for later and reads the token count from the counter.
static int phase1_impl(cpg_ctx* ctx, cpg_stream stream, const float* params, int V, int B, int L,
                       const cpg_wae_inputs* in, const cpg_loss_noise* nz, const cpg_train_hparams* hp,
                       float* coupled, float* mu, float* logvar, float* z, bool join, const StepNoiseArgs* gen = nullptr) {
    int rc = check_dims(V, B, L);
    if (rc) return rc;
    if (!ctx || !params || !in || !nz || !hp || !coupled || !in->tokens || !in->c) { set_error("cpg_wae_step_phase1: null argument"); return CPG_EINVAL; }
    if (!nz->z_prior_rf || !nz->rf_w || !nz->rf_b) { set_error("cpg_wae_step_phase1: RF-MMD noise is required"); return CPG_EINVAL; }
    const int R = hp->rf_dim;
    if (R < 1 || R > 4096) { set_error("rf_dim out of range"); return CPG_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    if ((rc = ensure_workspace(ctx, B, L, V, R, s))) return rc;
    Workspace& w = ctx->ws;
    ParamLayout lay = make_layout(V);
    const Lanes ln = lanes(ctx, s);
    cudaStream_t q = ln.s;
    // explicit noise inputs are ordered before the caller's lane; noise of cpg_fill_step_noise_overlapped / `gen` is
    // produced on lane s itself
    // (also: `coupled` may have been initialised on the caller's stream just before this call)
    order(ctx, ln, s, q);
    const Mark latent_ready = forward_impl(ctx, ln, params, lay, V, B, L, in, w.mu, w.logvar, w.z, true, false, gen,
                                           join ? coupled + 0 : nullptr);
    // the prior's random features do not depend on the batch: lane s, under the encoder recurrence
    const bool rf_tc = rf_uses_tc(B, R);
    if (rf_tc) {
        launch_prep_rf_tiles(q, nz->rf_w, R, w.rf_tiles);
        if ((rc = launch_rf_feat_tc(q, nz->z_prior_rf, w.rf_tiles, nz->rf_b, B, R, hp->mmd_sigma, nullptr, w.rf_part2))) return rc;
        launch_rf_colsum_final(q, w.rf_part2, rf_tc_parts(B), R, coupled + 8 + R);
    } else {
        launch_sgemm(q, B, R, ZD, 1.f, nz->z_prior_rf, ZD, 1, nz->rf_w, R, 1, 0.f, w.rf_pre2, R, nullptr, 1, nullptr);
        launch_rf_colsum(q, w.rf_pre2, nz->rf_b, B, R, hp->mmd_sigma, w.rf_part2, w.rf_nchunk, coupled + 8 + R);
    }
    // local statistics that couple the batch: RF feature sums, token count, latent sums -- they depend on
    // (mu, logvar, z) only and run on lane s under the decoder recurrence
    wait_mark(q, latent_ready);
    if (rf_tc) {
        if ((rc = launch_rf_feat_tc(q, w.z, w.rf_tiles, nz->rf_b, B, R, hp->mmd_sigma, w.rf_pre1, w.rf_part))) return rc;
        launch_rf_colsum_final(q, w.rf_part, rf_tc_parts(B), R, coupled + 8);
    } else {
        launch_sgemm(q, B, R, ZD, 1.f, w.z, ZD, 1, nz->rf_w, R, 1, 0.f, w.rf_pre1, R, nullptr, 1, nullptr);
        launch_rf_colsum(q, w.rf_pre1, nz->rf_b, B, R, hp->mmd_sigma, w.rf_part, w.rf_nchunk, coupled + 8);
    }
    // (the fused step runs the latent statistics after the RF gradient chain, see phase2_impl)
    if (join) launch_latent_stats(q, w.mu, w.logvar, B, w.lat_part, w.lat_nparts, coupled + 2);
    if (mu) dev_copy(mu, w.mu, (size_t)B * ZD * 4, s);
    if (logvar) dev_copy(logvar, w.logvar, (size_t)B * ZD * 4, s);
    if (z) dev_copy(z, w.z, (size_t)B * ZD * 4, s);
    ctx->have_stash = true;
    ctx->stash_gen++;
    return check_launch("cpg_wae_step_phase1");
}

int cpg_wae_step_phase1(cpg_ctx* ctx, cpg_stream stream, const float* params, int V, int B, int L,
                        const cpg_wae_inputs* in, const cpg_loss_noise* nz, const cpg_train_hparams* hp,
                        float* coupled, float* mu, float* logvar, float* z) {
    // the dependent chain on the library's highest-priority stream, forked from / joined to the caller's (as the fused step does)
    if (!ctx) { set_error("cpg_wae_step_phase1: null argument"); return CPG_EINVAL; }
    cudaStream_t caller = (cudaStream_t)stream, chain = caller;
    const Lanes l0 = lanes(ctx, caller);
    if (l0.on && g_opt_chain_priority && ctx->chain_stream != nullptr) {
        chain = (cudaStream_t)ctx->chain_stream;
        noise_join(ctx, caller);
        order(ctx, l0, caller, chain);
    }
    const int rc = phase1_impl(ctx, chain, params, V, B, L, in, nz, hp, coupled, mu, logvar, z, true);
    order(ctx, l0, chain, caller);
    return rc;
}

// Phase 2.  `fused` = called by the single-GPU step right after phase1_impl(join = false): the coupled statistics are
// local and still on lane s (where their consumers run), and the token count is read from the preparation's own counter.
static int phase2_impl(cpg_ctx* ctx, cpg_stream stream, const float* params, float* grads, int V, int B, int L,
                       const cpg_wae_inputs* in, const cpg_loss_noise* nz, const cpg_train_hparams* hp,
                       const float* coupled, float* scalars, float* logits, bool fused_step) {
    int rc = check_dims(V, B, L);
    if (rc) return rc;
    if (!ctx || !params || !grads || !in || !nz || !hp || !coupled) { set_error("cpg_wae_step_phase2: null argument"); return CPG_EINVAL; }
    Workspace& w = ctx->ws;
    if (!ctx->have_stash || w.B != B || w.L != L || w.V != V || w.R < hp->rf_dim) {
        set_error("cpg_wae_step_phase2: phase1 was not run for this shape");
        return CPG_EINVAL;
    }
    if (hp->z_regu == CPG_ZREGU_MMD && (!nz->z_prior_full || (hp->global_batch > 0 && hp->global_batch != B))) {
        set_error("z_regu_loss='mmd' differentiates the full-kernel MMD of the whole batch: it needs z_prior_full and runs on one GPU "
                  "(under data parallelism use mmdrf, whose coupling is two R-float sums)");
        return CPG_EINVAL;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int R = hp->rf_dim;
    const int Bg = hp->global_batch > 0 ? hp->global_batch : B;
    ParamLayout lay = make_layout(V);
    const Lanes ln = lanes(ctx, s);
    flush_deferred_noise(ctx, s);
    dev_memset(grads, 0, (size_t)lay.total * 4, s);
    // RF-MMD from the global feature sums, then the full-kernel MMD: lane s; the caller's lane waits for the RF gradient
    // right before the latent backward that consumes it, for the (logged) full-kernel MMD only at the end of the step
    if (!fused_step) order(ctx, ln, s, ln.s);       // `coupled` as the caller left it on its stream (exchange done)
    const float w_rf = hp->z_regu == CPG_ZREGU_MMDRF ? hp->beta : 0.f;
    const float* dz_rf = nullptr;
    Mark dz_ready = nullptr;
    {
        cudaStream_t q = ln.s;
        launch_rf_loss(q, coupled + 8, coupled + 8 + R, R, Bg, hp->mmd_sigma, w_rf, w.rf_coef, w.mmdrf_out);
        if (w_rf != 0.f) {
            if (rf_uses_tc(B, R)) {
                if ((rc = launch_rf_grad_tc(q, w.rf_pre1, w.rf_tiles, nz->rf_b, w.rf_coef, B, R, hp->mmd_sigma, w.dz_rf))) return rc;
            } else {
                launch_rf_grad_prep(q, w.rf_pre1, nz->rf_b, w.rf_coef, B, R, hp->mmd_sigma);
                launch_sgemm(q, B, ZD, R, 1.f, w.rf_pre1, R, 1, nz->rf_w, 1, R, 0.f, w.dz_rf, ZD, nullptr, 1, nullptr);
            }
            dz_rf = w.dz_rf;
        }
        const bool want_mmd = (hp->compute_full_mmd || hp->z_regu == CPG_ZREGU_MMD) && nz->z_prior_full;
        if (hp->z_regu == CPG_ZREGU_MMD) {                          // the full-kernel MMD is the regulariser: its gradient at z
            launch_mmd_full_grad(q, w.z, nz->z_prior_full, B, hp->mmd_sigma, hp->beta, w.dz_rf);
            dz_rf = w.dz_rf;
        }
        dz_ready = mark(ctx, ln, q);
        if (fused_step) {                           // log-only statistics: after the chain the latent backward waits for
            launch_int_to_float(q, ctx->ints, const_cast<float*>(coupled) + 0, 1);
            launch_latent_stats(q, w.mu, w.logvar, B, w.lat_part, w.lat_nparts, const_cast<float*>(coupled) + 2);
        }
        if (want_mmd) {
            // beside the chain the persistent Gram kernel runs on a part of the SMs only: a CTA per SM holds back CTAs of the
            // BPTT kernels that start under it (measured: 0.648 -> 0.629 ms/step with 64 CTAs instead of 148)
            const int saved = g_opt_mmd_grid;
            if (saved == 0 && ln.on) g_opt_mmd_grid = 64;
            rc = launch_mmd_full(q, w.z, nz->z_prior_full, B, hp->mmd_sigma, w.mmd_ws, w.mmd_out);
            g_opt_mmd_grid = saved;
            if (rc) return rc;
        }
    }
    // reconstruction loss fwd+bwd with the global token count
    DecOutArgs a = dec_out_args(ctx, in, V, B, L);
    a.ntok = coupled + 0;
    a.ntok_i = fused_step ? ctx->ints : nullptr;    // single GPU: the count prep_tokens left on this lane
    a.fused_ce = 1;
    a.logits_out = logits;
    a.dh_out = w.dec_dh_out;
    noise_join(ctx, s);                             // out-dropout mask generated on lane s (normally joined by the latent layers)
    if (!fused_step) order(ctx, ln, ln.t, s);       // the (exchanged) token count lives on lane t
    launch_dec_out(s, a, ctx->sm_count);
    // the ordered reduction of its partials (fc gradients, NLL sum): lane t, under the decoder BPTT
    order(ctx, ln, s, ln.t);
    launch_dec_out_reduce(ln.t, a, ctx->sm_count, grads + lay.off[P_FC_W], grads + lay.off[P_FC_B], w.nll_sum);
    LatentBwdArgs la;
    memset(&la, 0, sizeof(la));
    la.dz_rf = dz_rf;
    la.w_kl = hp->z_regu == CPG_ZREGU_KL ? hp->beta : 0.f;
    la.w_klsm = hp->lambda_logvar_kl;
    la.w_l1 = hp->lambda_logvar_l1;
    la.B_global = Bg;
    if (scalars) {                                  // logging scalars: lane s (after the MMD), once lane t has the NLL sum
        order(ctx, ln, ln.t, ln.s);
        ComposeArgs c;
        memset(&c, 0, sizeof(c));
        c.ntok = coupled + 0; c.nll_sum = w.nll_sum; c.lat_sums = coupled + 2;
        c.mmd = ((hp->compute_full_mmd || hp->z_regu == CPG_ZREGU_MMD) && nz->z_prior_full) ? w.mmd_out : nullptr;
        c.mmdrf = w.mmdrf_out;
        c.beta = hp->beta; c.lambda_l1 = hp->lambda_logvar_l1; c.lambda_kl = hp->lambda_logvar_kl;
        c.z_regu = hp->z_regu; c.B_global = Bg; c.out = scalars;
        launch_compose_scalars(ln.s, c);
    }
    backward_impl(ctx, ln, params, lay, grads, V, B, L, in, la, dz_ready);
    return check_launch("cpg_wae_step_phase2");
}

int cpg_wae_step_phase2(cpg_ctx* ctx, cpg_stream stream, const float* params, float* grads, int V, int B, int L,
                        const cpg_wae_inputs* in, const cpg_loss_noise* nz, const cpg_train_hparams* hp,
                        const float* coupled, float* scalars, float* logits) {
    if (!ctx) { set_error("cpg_wae_step_phase2: null argument"); return CPG_EINVAL; }
    cudaStream_t caller = (cudaStream_t)stream, chain = caller;
    const Lanes l0 = lanes(ctx, caller);
    if (l0.on && g_opt_chain_priority && ctx->chain_stream != nullptr) {
        chain = (cudaStream_t)ctx->chain_stream;
        order(ctx, l0, caller, chain);
    }
    const int rc = phase2_impl(ctx, chain, params, grads, V, B, L, in, nz, hp, coupled, scalars, logits, false);
    order(ctx, l0, chain, caller);
    return rc;
}

int cpg_clip_adam_step(cpg_ctx* ctx, cpg_stream stream, float* params, float* grads, float* m, float* v, int V,
                       const cpg_train_hparams* hp, float* grad_norm_out) {
    if (!ctx || !params || !grads || !m || !v || !hp) { set_error("cpg_clip_adam_step: null argument"); return CPG_EINVAL; }
    if (V < 4 || V > VMAX) { set_error("n_vocab must be in [4, 32]"); return CPG_EINVAL; }
    if (hp->adam_step < 1) { set_error("adam_step is 1-based"); return CPG_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    if (ctx->base == nullptr && (rc = ensure_workspace(ctx, 1, 2, V, 500, s))) return rc;
    Workspace& w = ctx->ws;
    ParamLayout lay = make_layout(V);
    launch_clip_adam_fused(s, params, grads, m, v, lay.total, lay.off[P_EMB], lay.size[P_EMB], hp->clip_norm, w.norm_part,
                           grad_norm_out, w.clip_coef, adam_hyper(hp), ctx->sm_count,
                           g_opt_adam_fused ? reinterpret_cast<unsigned*>(ctx->ints + 8) : nullptr);
    return check_launch("cpg_clip_adam_step");
}

// The fused single-GPU iteration.  The dependent chain runs on the library's highest-priority stream (forked from and joined
// to the caller's), so that its kernels are dispatched ahead of the co-running loss / reduction kernels of the other lanes.
static int train_step_impl(cpg_ctx* ctx, cpg_stream stream, float* params, float* grads, float* m, float* v, int V, int B,
                           int L, const cpg_wae_inputs* in, const cpg_loss_noise* nz, const cpg_train_hparams* hp,
                           float* scalars, float* mu, float* logvar, float* z, float* logits, const StepNoiseArgs* gen) {
    if (!ctx || !hp) { set_error("cpg_wae_train_step: null argument"); return CPG_EINVAL; }
    cpg_train_hparams h = *hp;
    h.global_batch = B;
    int rc = ensure_workspace(ctx, B, L, V, h.rf_dim, (cudaStream_t)stream);
    if (rc) return rc;
    float* cpl = ctx->ws.coupled;
    cudaStream_t caller = (cudaStream_t)stream, chain = caller;
    const Lanes l0 = lanes(ctx, caller);
    if (l0.on && g_opt_chain_priority && ctx->chain_stream != nullptr) {
        chain = (cudaStream_t)ctx->chain_stream;
        noise_join(ctx, caller);
        order(ctx, l0, caller, chain);
    }
    if ((rc = phase1_impl(ctx, chain, params, V, B, L, in, nz, &h, cpl, mu, logvar, z, false, gen))) return rc;
    if ((rc = phase2_impl(ctx, chain, params, grads, V, B, L, in, nz, &h, cpl, scalars, logits, true))) return rc;
    float* gn = scalars ? scalars + SC_GRAD_NORM : nullptr;
    rc = cpg_clip_adam_step(ctx, chain, params, grads, m, v, V, &h, gn);
    order(ctx, l0, chain, caller);
    return rc;
}

int cpg_wae_train_step(cpg_ctx* ctx, cpg_stream stream, float* params, float* grads, float* m, float* v, int V, int B,
                       int L, const cpg_wae_inputs* in, const cpg_loss_noise* nz, const cpg_train_hparams* hp,
                       float* scalars, float* mu, float* logvar, float* z, float* logits) {
    return train_step_impl(ctx, stream, params, grads, m, v, V, B, L, in, nz, hp, scalars, mu, logvar, z, logits, nullptr);
}

// Philox noise + the whole iteration as ONE call; from the third call with the same buffers and settings on, a replay
// of a captured CUDA graph (the ~50 kernels, memsets and the fork / join of the side stream) whose per-step scalars
// (beta, Adam bias corrections, noise counter) are refreshed by one kernel-node parameter update.
int cpg_wae_train_step_philox(cpg_ctx* ctx, cpg_stream stream, float* params, float* grads, float* m, float* v, int V, int B,
                              int L, const int64_t* tokens, const cpg_step_noise_buffers* nb, const cpg_train_hparams* hp,
                              uint64_t seed, uint32_t noise_step, float p_word, float p_out, float* scalars) {
    if (!ctx || !hp || !nb || !tokens || !params || !grads || !m || !v) { set_error("cpg_wae_train_step_philox: null argument"); return CPG_EINVAL; }
    if (!nb->eps || !nb->c || !nb->word_drop || !nb->out_keep || !nb->z_prior_rf || !nb->rf_w || !nb->rf_b) {
        set_error("cpg_wae_train_step_philox: noise buffers missing"); return CPG_EINVAL;
    }
    cpg_wae_inputs in;
    in.tokens = tokens; in.eps = nb->eps; in.c = nb->c; in.word_drop = nb->word_drop; in.out_keep = nb->out_keep;
    in.p_out_dropout = p_out;
    cpg_loss_noise nz;
    nz.z_prior_full = nb->z_prior_full; nz.z_prior_rf = nb->z_prior_rf; nz.rf_w = nb->rf_w; nz.rf_b = nb->rf_b;
    if (B < 1 || L < 1) { set_error("cpg_wae_train_step_philox: bad argument"); return CPG_EINVAL; }
    StepNoiseArgs gen;
    gen.seed = seed; gen.step = noise_step; gen.B = B; gen.L = L; gen.p_word = p_word; gen.p_out = p_out;
    gen.eps = nb->eps; gen.c = nb->c; gen.word_drop = nb->word_drop; gen.out_keep = nb->out_keep;
    gen.zp_full = nb->z_prior_full; gen.zp_rf = nb->z_prior_rf;
    auto body = [&]() -> int {
        return train_step_impl(ctx, stream, params, grads, m, v, V, B, L, &in, &nz, hp, scalars, nullptr, nullptr, nullptr, nullptr, &gen);
    };
#ifdef CPG_EMU
    return body();
#else
    cudaStream_t s = (cudaStream_t)stream;
    if (!g_opt_graph || g_profile_on) return body();
    char kb[512];
    snprintf(kb, sizeof(kb), "%p|%p|%p|%p|%p|%p|%d|%d|%d|%p|%p|%p|%p|%p|%p|%p|%p|%p|%g|%g|%g|%g|%g|%g|%g|%d|%g|%d|%d|%d|%llu|%g|%g|%d|%d|%d|%d|%d",
             (void*)ctx, (void*)s, (void*)params, (void*)grads, (void*)m, (void*)v, V, B, L, (const void*)tokens, (void*)nb->eps, (void*)nb->c,
             (void*)nb->word_drop, (void*)nb->out_keep, (void*)nb->z_prior_full, (void*)nb->z_prior_rf, (void*)nb->rf_w, (void*)nb->rf_b,
             hp->lr, hp->beta1, hp->beta2, hp->adam_eps, hp->clip_norm, hp->lambda_logvar_l1, hp->lambda_logvar_kl, hp->z_regu,
             hp->mmd_sigma, hp->rf_dim, hp->compute_full_mmd, hp->beta != 0.f ? 1 : 0, (unsigned long long)seed, p_word, p_out,
             scalars ? 1 : 0, g_opt_side_stream, g_opt_gru_tc, g_opt_bptt_fused, g_opt_dec_out_tc * 16 + g_opt_wgrad_tc * 4 + g_opt_mmd_tc + 64 * g_opt_latent_tc + 256 * g_opt_latent_rows + 65536 * g_opt_chain_priority + 131072 * g_opt_rf_tc + 524288 * g_opt_adam_fused + 1048576 * g_opt_wgrad_dense_tc + 2097152 * (g_opt_mmd_grid & 255) + g_opt_rf_grid * 7 + g_opt_wd_grid * 13 + g_opt_matmul_terms * 101);
    StepGraph* g = find_graph(kb + std::string(scalars ? std::to_string((uintptr_t)scalars) : ""));
    StepDyn* dyn_dev = reinterpret_cast<StepDyn*>(ctx->ints + 32);
    const StepDyn dv = make_dyn(hp, noise_step);
    if (g->exec == nullptr) {
        if (g->seen < 1) { g->seen++; return body(); }          // first sight: eager (allocations, attribute set-up)
        // capture
        const long long l0 = g_launch_count;
        if (cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed) != cudaSuccess) { cudaGetLastError(); return body(); }
        k_set_dyn<<<1, 32, 0, s>>>(dv, dyn_dev);
        g_dyn = dyn_dev;
        int rc = body();
        g_dyn = nullptr;
        cudaGraph_t graph = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(s, &graph);
        const long long captured = g_launch_count - l0 + 1;
        g_launch_count = l0;
        if (rc || ce != cudaSuccess || graph == nullptr) {
            cudaGetLastError();
            if (graph) cudaGraphDestroy(graph);
            g->seen = -1000000;                                  // do not try again for this key
            return rc ? rc : body();
        }
        size_t nn = 0;
        cudaGraphGetNodes(graph, nullptr, &nn);
        std::vector<cudaGraphNode_t> nodes(nn);
        cudaGraphGetNodes(graph, nodes.data(), &nn);
        for (auto nd : nodes) {
            cudaGraphNodeType ty;
            if (cudaGraphNodeGetType(nd, &ty) != cudaSuccess || ty != cudaGraphNodeTypeKernel) continue;
            cudaKernelNodeParams kp;
            if (cudaGraphKernelNodeGetParams(nd, &kp) == cudaSuccess && kp.func == (void*)k_set_dyn) { g->dyn_node = nd; break; }
        }
        cudaGraphExec_t ex = nullptr;
        if (g->dyn_node == nullptr || cudaGraphInstantiate(&ex, graph, cudaGraphInstantiateFlagUseNodePriority) != cudaSuccess) {
            cudaGetLastError();
            cudaGraphDestroy(graph);
            g->seen = -1000000;
            return body();
        }
        cudaGraphDestroy(graph);
        g->exec = ex;
        g->launches = captured;
    }
    if (g->seen < 0) return body();
    // replay: refresh the per-step scalars through the head node's arguments, then launch
    StepDyn hv = dv;
    void* kargs[2] = {&hv, &dyn_dev};
    cudaKernelNodeParams kp;
    memset(&kp, 0, sizeof(kp));
    kp.func = (void*)k_set_dyn; kp.gridDim = dim3(1); kp.blockDim = dim3(32); kp.sharedMemBytes = 0; kp.kernelParams = kargs; kp.extra = nullptr;
    if (cudaGraphExecKernelNodeSetParams(g->exec, g->dyn_node, &kp) != cudaSuccess || cudaGraphLaunch(g->exec, s) != cudaSuccess) {
        set_error(std::string("cpg_wae_train_step_philox: graph replay failed: ") + cudaGetErrorString(cudaGetLastError()));
        return CPG_ECUDA;
    }
    g_launch_count += g->launches;
    return CPG_OK;
#endif
}

// ---- per-step scalars in device memory, for callers that capture the iteration themselves --------------------
int cpg_step_dyn_write(cpg_ctx* ctx, cpg_stream stream, const cpg_train_hparams* hp, uint32_t noise_step) {
    if (!ctx || !hp) { set_error("cpg_step_dyn_write: null argument"); return CPG_EINVAL; }
    if (hp->adam_step < 1) { set_error("adam_step is 1-based"); return CPG_EINVAL; }
#ifndef CPG_EMU
    k_set_dyn<<<1, 32, 0, (cudaStream_t)stream>>>(make_dyn(hp, noise_step), reinterpret_cast<StepDyn*>(ctx->ints + 32));
#endif
    return check_launch("cpg_step_dyn_write");
}

int cpg_step_dyn_use(cpg_ctx* ctx, int on) {
    if (!ctx) { set_error("cpg_step_dyn_use: null argument"); return CPG_EINVAL; }
    g_dyn = on ? reinterpret_cast<const StepDyn*>(ctx->ints + 32) : nullptr;
    return CPG_OK;
}

// ---- data-parallel helpers (cpg_b200/parallel.py) ----------------------------------------------------------
void* cpg_side_stream(cpg_ctx* ctx) {
    if (!ctx || !side_ready(ctx)) return nullptr;
    return ctx->side_stream;
}
void* cpg_aux_stream(cpg_ctx* ctx) {
    if (!ctx || !side_ready(ctx)) return nullptr;
    return ctx->aux_stream;
}

int cpg_dp_tail_count(void) { return DP_TAIL; }

int cpg_dp_pack_tail(cpg_ctx* ctx, cpg_stream stream, float* tail) {
    if (!ctx || !tail || ctx->base == nullptr) { set_error("cpg_dp_pack_tail: null argument / no step was run"); return CPG_EINVAL; }
    launch_dp_pack_tail((cudaStream_t)stream, ctx->ws.nll_sum, tail);
    return check_launch("cpg_dp_pack_tail");
}

int cpg_dp_apply_tail(cpg_ctx* ctx, cpg_stream stream, const float* tail, float* scalars) {
    if (!ctx || !tail || !scalars) { set_error("cpg_dp_apply_tail: null argument"); return CPG_EINVAL; }
    launch_dp_apply_tail((cudaStream_t)stream, tail, scalars);
    return check_launch("cpg_dp_apply_tail");
}

int cpg_wae_decode_teacher(cpg_ctx* ctx, cpg_stream stream, const float* params, int V, int B, int L,
                           const cpg_wae_inputs* in, const float* z, float* logits) {
    int rc = check_dims(V, B, L);
    if (rc) return rc;
    if (!ctx || !params || !in || !in->tokens || !in->c || !z || !logits) { set_error("cpg_wae_decode_teacher: null argument"); return CPG_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    if ((rc = ensure_workspace(ctx, B, L, V, ctx->ws.R > 0 ? ctx->ws.R : 500, s))) return rc;
    Workspace& w = ctx->ws;
    flush_deferred_noise(ctx, s);
    noise_join(ctx, s);
    launch_prep_tokens(s, in->tokens, in->word_drop, B, L, V, w.tok, w.tokd, w.tgt, ctx->ints, ctx->ints + 1);
    launch_prep_weights(s, params, make_layout(V), V, w.d);
    launch_make_zc(s, z, in->c, B, w.zc);
    launch_sgemm(s, B, 3 * DEC_HP, DEC_HP, 1.f, w.zc, DEC_HP, 1, w.d.wizc_t, 3 * DEC_HP, 1, 0.f, w.rowbias,
                 3 * DEC_HP, nullptr, 1, nullptr);
    GruSeq q;
    memset(&q, 0, sizeof(q));
    q.tok = w.tokd; q.table = w.d.t_dec; q.rowbias = w.rowbias; q.whh_t = w.d.whh_t_dec; q.bhn = w.d.bhn_dec;
    q.h0 = w.zc; q.hs = w.dec_hs;
    launch_gru_fwd_dec(s, q, B, L);
    DecOutArgs a = dec_out_args(ctx, in, V, B, L);
    a.logits_out = logits;
    a.part_nll = nullptr;
    launch_dec_out(s, a, ctx->sm_count);
    ctx->have_stash = false;
    return check_launch("cpg_wae_decode_teacher");
}

}  // extern "C"
