// Context + workspace of the C ABI (internal).
#pragma once
#include <string>
#include "kernels.h"
#include "dec_out.h"
#include "latent.h"
#include "wgrad.h"
#include "adam.h"
#include "../../include/cpg_b200.h"

namespace cpg {

struct Workspace {
    // geometry this layout was made for
    int B = 0, L = 0, V = 0, R = 0;
    Derived d;
    uint8_t *tok, *tokd, *tgt;
    float *enc_hs[2], *enc_gates[2], *enc_dg[2];
    float *hfin, *mu, *logvar, *z, *zc, *rowbias;
    float *dec_hs, *dec_gates, *dec_dg, *dec_dh_out;
    float *drow, *dh0, *dmu, *dlv, *dhfin;
    float *rf_pre1, *rf_pre2, *rf_part, *rf_part2, *rf_sum1, *rf_sum2, *rf_coef, *dz_rf;
    unsigned char* rf_tiles;
    float *lat_part, *lat_sums;
    float *mmd_ws, *mmd_out, *mmdrf_out;
    float *do_part_w, *do_part_b, *do_part_nll, *nll_sum;
    float *wg_part, *dt_part, *wg_part_dec, *dt_part_dec, *wg_part_enc1, *dt_part_enc1, *dT_enc[2], *dT_dec, *dwizc;
    float *gemm_ws, *colsum_ws, *hg_part, *wd_part, *emb_dec;
    float *norm_part, *clip_coef, *scalars, *ntok_f, *coupled;
    int lat_nparts, rf_nchunk, gemm_splits;
};

}  // namespace cpg

struct cpg_ctx {
    int device = 0;
    int sm_count = 148;
    void* base = nullptr;          // one device allocation
    size_t capacity = 0;
    int* ints = nullptr;           // [0] ntok (int), [1] sticky token-range error flag
    cpg::Workspace ws;
    bool have_stash = false;
    int64_t stash_gen = 0;         // bumped by every call that (re)writes the BPTT stash: a backward names the one it expects
    // scratch of the stand-alone loss ops (cpg_mmd_rf / cpg_mmd_full): separate from the step workspace so that a
    // loss evaluated between a forward and its backward can neither re-lay-out nor overwrite the stash
    void* aux = nullptr;
    size_t aux_capacity = 0;
    int64_t launches = 0;
    // Two internal streams next to the caller's: the latency-bound loss kernels (statistics, RF-MMD, full-kernel MMD) on
    // `side_stream`, weight-derived forms / ordered reductions / weight-gradient products on `aux_stream`; they run under
    // the recurrences of the caller's stream.  Dependencies are events from a small rotating pool (api_wae.cu).
    void* side_stream = nullptr;
    void* aux_stream = nullptr;
    void* chain_stream = nullptr;  // highest-priority stream for the dependent chain of the fused step (forked from / joined to the caller's)
    void* ev_pool[32] = {nullptr};
    int ev_next = 0;
    void* ev_noise = nullptr;      // noise generated on the side stream by cpg_fill_step_noise_overlapped, not yet joined
    bool noise_pending = false;
    // cpg_fill_step_noise_overlapped only RECORDS its request; the next forward draws the word-dropout mask inside its token
    // preparation and the rest on the loss lane once the preparation is through (api_wae.cu: forward_impl)
    bool gen_deferred = false;
    cpg::StepNoiseArgs gen_args;
};

namespace cpg {
void set_error(const std::string& msg);
int ensure_workspace(cpg_ctx* ctx, int B, int L, int V, int R, cudaStream_t stream);
int ensure_aux(cpg_ctx* ctx, size_t bytes, cudaStream_t stream);
}
