#pragma once
#include "cpg_common.cuh"

namespace cpg {

struct InputGradArgs {
    const float* emb;               // [V][150]
    const float* enc_wih[2];        // [240][150]
    const float* dec_wih;           // [306][252]
    const float* dT_enc[2];         // [V][4*80]
    const float* dT_dec;            // [V][4*104]
    const float* dwizc;             // [312][104] padded gradient of W_ih[:,150:]
    float* g_emb;
    float* emb_dec;                 // [V][150] scratch: the decoder table's share of the embedding gradient (computed early)
    float* g_enc_wih[2]; float* g_enc_bih[2]; float* g_enc_bhh[2];
    float* g_dec_wih; float* g_dec_bih; float* g_dec_bhh;
    int V;
};

int wgrad_splits(int B, int L, int sm_count);
int dtable_splits(int B, int L, int sm_count);
// dW_hh (and, on the tensor-core path, the token-table gradient dT as well: returns true then).
bool launch_wgrad_hh(cudaStream_t s, int HP, int H, const float* dg, const float* hs, const float* h0,
                     const uint8_t* tok, int reverse, int V, int B, int L, int sm_count, float* part, float* dt_part,
                     float* dW, float* dT,
                     cudaStream_t reduce_stream = nullptr, void* reduce_event = nullptr,
                     int dg_rounded = 0);   // reductions on another stream (event-ordered)
void launch_wgrad_hh_simt(cudaStream_t s, int HP, int H, const float* dg, const float* hs, const float* h0, int B, int L,
                          int sm_count, float* part, float* dW);
// tcgen05 version (wgrad_tc.cu): partials part_w [nsplit][3*HP][HP] and part_t [nsplit][V][4*HP]
int wgrad_tc_splits(int sm_count);
int launch_wgrad_tc(cudaStream_t s, int HP, const float* dg, const float* hs, const float* h0, const uint8_t* tok,
                    int reverse, int B, int L, int V, int sm_count, float* part_w, float* part_t, int* nsplit_out, int dg_rounded);
// ordered sums of per-CTA partials in the layouts above -> dW_hh [3H][H] (unpadded) and dT [V][4*HP]
void launch_wgrad_partial_reduce(cudaStream_t s, int HP, int H, int V, const float* part_w, const float* part_t, int nsplit,
                                 float* dW, float* dT);
extern int g_opt_wgrad_tc;
bool wgrad_uses_tc(int nrows);      // would launch_wgrad_hh take the tensor-core path for this many rows?
void launch_dtable(cudaStream_t s, int HP, const float* dg, const uint8_t* tok, int B, int L, int reverse, int V,
                   int sm_count, float* part, float* dT);
// s_emb: stream for the embedding gradient; parts: 1 = encoder W_ih / biases, 2 = decoder W_ih / biases,
// 8 = the decoder table's share of the embedding gradient (-> emb_dec), 4 = embedding gradient (encoder tables + emb_dec)
void launch_input_grads(cudaStream_t s, const InputGradArgs& a, cudaStream_t s_emb = nullptr, int parts = 15);

}  // namespace cpg
