// Token preparation and per-step derived weight forms.
//
// Reference call sites replaced: nn.Embedding lookups (models/model.py:102,
// models/decoder.py:67), WordDropout (models/decoder.py:117-133) and the input
// half of nn.GRU (models/encoder.py:42, models/decoder.py:77).  Because the
// vocabulary is tiny (V <= 32), the embedding gather and the input projection
// x_t W_ih^T + b_ih collapse into a token -> gate table T = E W_ih[:, :150]^T + b
// that is rebuilt from the current weights every step (2 V 150 3H flop).
#include "latent.h"
#ifndef CPG_EMU
#include "tc_gru.cuh"
#endif

namespace cpg {

ParamLayout make_layout(int V) {
    ParamLayout l;
    const int64_t ge = 3 * ENC_H, gd = 3 * DEC_H;
    int64_t sz[P_COUNT] = {
        (int64_t)V * EMB,
        ge * EMB, ge * ENC_H, ge, ge,
        ge * EMB, ge * ENC_H, ge, ge,
        (int64_t)ZD * 2 * ENC_H, ZD, (int64_t)ZD * 2 * ENC_H, ZD,
        gd * DEC_IN, gd * DEC_H, gd, gd,
        (int64_t)V * DEC_H, V};
    int64_t o = 0;
    for (int i = 0; i < P_COUNT; ++i) {
        l.off[i] = o;
        l.size[i] = sz[i];
        o += (sz[i] + 3) / 4 * 4;
    }
    l.total = o;
    return l;
}

size_t derived_floats(int V) {
    size_t n = 0;
    n += 2 * (size_t)V * 3 * ENC_H;          // t_enc
    n += (size_t)V * 3 * DEC_HP;             // t_dec
    n += 2 * (size_t)ENC_H * 3 * ENC_H;      // whh_t_enc
    n += 4 * (size_t)DEC_HP * 3 * DEC_HP;    // whh_t_dec, whh_dec, wizc_t, wizc
    n += 2 * ENC_H + DEC_HP;                 // bhn
    n += (size_t)VMAX * DEC_HP + VMAX;       // fc
    return n + 64;
}

// tokens int64 [B,L] (data_processing/dataset.py batch.text) -> uint8 planes.
//   tok  : encoder input
//   tokd : decoder input after word dropout (mask==1 -> <unk>)
//   tgt  : next-token targets, <pad> appended (losses.py:25-26)
//   ntok : number of non-<pad> targets of this batch (CE denominator, losses.py:27-30)
// gen != 0: the word-dropout mask is drawn here (the arithmetic of k_step_noise, stream id 16 * step + 2) and also
// written to word_drop_out, instead of being read from word_drop.
__global__ void k_prep_tokens(const int64_t* __restrict__ tokens, const uint8_t* __restrict__ word_drop,
                              int B, int L, int V, uint8_t* __restrict__ tok, uint8_t* __restrict__ tokd,
                              uint8_t* __restrict__ tgt, int* __restrict__ ntok, int* __restrict__ err,
                              int gen, uint64_t seed, uint32_t step, float p_word, const StepDyn* __restrict__ dyn,
                              uint8_t* __restrict__ word_drop_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int n = B * L;
    int cnt = 0;
    if (i < n) {
        int64_t w = tokens[i];
        if (w < 0 || w >= V) { atomicExch(err, 1); w = UNK; }
        int t = i % L;
        int64_t nx = (t + 1 < L) ? tokens[i + 1] : (int64_t)PAD;
        if (nx < 0 || nx >= V) nx = UNK;
        bool drop;
        if (gen) {
            uint32_t r[4];
            Philox::gen(seed, (uint64_t)i, (dyn != nullptr ? dyn->noise_step : step) * 16u + 2, r);
            drop = u32_to_unit_open(r[0]) < p_word;      // decoder.py:124-127
            word_drop_out[i] = drop ? 1 : 0;
        } else {
            drop = word_drop != nullptr && word_drop[i];
        }
        tok[i] = (uint8_t)w;
        tokd[i] = drop ? (uint8_t)UNK : (uint8_t)w;
        tgt[i] = (uint8_t)nx;
        cnt = (nx != PAD) ? 1 : 0;
    }
    unsigned m = __ballot_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(ntok, __popc(m));
}

void launch_prep_tokens(cudaStream_t s, const int64_t* tokens, const uint8_t* word_drop, int B, int L, int V,
                        uint8_t* tok, uint8_t* tokd, uint8_t* tgt, int* ntok, int* err, const StepNoiseArgs* gen) {
    cudaMemsetAsync(ntok, 0, sizeof(int), s);
    int n = B * L;
    CPG_LAUNCH(k_prep_tokens, ceil_div(n, 256), 256, 0, s, tokens, word_drop, B, L, V, tok, tokd, tgt, ntok, err,
               gen != nullptr ? 1 : 0, gen != nullptr ? gen->seed : 0ull, gen != nullptr ? gen->step : 0u,
               gen != nullptr ? gen->p_word : 0.f, g_dyn, gen != nullptr ? gen->word_drop : nullptr);
}

struct PrepArgs {
    const float* emb;
    const float* enc_wih[2]; const float* enc_whh[2]; const float* enc_bih[2]; const float* enc_bhh[2];
    const float* dec_wih; const float* dec_whh; const float* dec_bih; const float* dec_bhh;
    const float* fc_w; const float* fc_b;
    const float* wmu; const float* wlv;
    Derived d;
    int V;
};

// blockIdx.y selects the task; blockIdx.x strides over its elements.
// task_map: 4 bits per blockIdx.y = the task it runs (the encoder's forms are a launch of their own: its recurrence starts first)
__global__ void k_prep_weights(PrepArgs a, unsigned long long task_map) {
    const int task = (int)((task_map >> (4 * blockIdx.y)) & 15);
    const int stride = gridDim.x * blockDim.x;
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int V = a.V;
    // token tables: one warp per output, lanes over the 150 embedding columns (both operands read coalesced; the
    // one-thread-per-output version walked W_ih rows with a 600-byte stride and took 22 us for 3 MFLOP)
    const int lane = threadIdx.x & 31;
    const int warp_g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = stride >> 5;
    if (task < 2) {                     // encoder token tables
        const int d = task, G = 3 * ENC_H;
        for (int i = warp_g; i < V * G; i += nwarp) {
            int v = i / G, g = i % G;
            const float* e = a.emb + (size_t)v * EMB;
            const float* w = a.enc_wih[d] + (size_t)g * EMB;
            float acc = 0.f;
            for (int k = lane; k < EMB; k += 32) acc = fmaf(e[k], w[k], acc);
            acc = warp_sum(acc);
            if (lane == 0) {
                acc += a.enc_bih[d][g];
                if (g < 2 * ENC_H) acc += a.enc_bhh[d][g];      // b_hr, b_hz fold into the table; b_hn cannot
                a.d.t_enc[d][i] = acc;
            }
        }
    } else if (task == 2) {             // decoder token table (embedding columns of W_ih only)
        const int G = 3 * DEC_HP;
        for (int i = warp_g; i < V * G; i += nwarp) {
            int v = i / G, gp = i % G, gate = gp / DEC_HP, j = gp % DEC_HP;
            float acc = 0.f;
            if (j < DEC_H) {
                int g = gate * DEC_H + j;
                const float* e = a.emb + (size_t)v * EMB;
                const float* w = a.dec_wih + (size_t)g * DEC_IN;
                for (int k = lane; k < EMB; k += 32) acc = fmaf(e[k], w[k], acc);
                acc = warp_sum(acc);
                acc += a.dec_bih[g];
                if (gate < 2) acc += a.dec_bhh[g];
            }
            if (lane == 0) a.d.t_dec[i] = acc;
        }
    } else if (task < 5) {              // encoder W_hh^T  [k][g]
        const int d = task - 3, G = 3 * ENC_H;
        for (int i = t0; i < ENC_H * G; i += stride) {
            int k = i / G, g = i % G;
            a.d.whh_t_enc[d][i] = a.enc_whh[d][(size_t)g * ENC_H + k];
        }
        for (int i = t0; i < ENC_H; i += stride) a.d.bhn_enc[d][i] = a.enc_bhh[d][2 * ENC_H + i];
    } else if (task == 5) {             // decoder W_hh^T and (W_ih[:,150:])^T, padded [104][312]
        const int G = 3 * DEC_HP;
        for (int i = t0; i < DEC_HP * G; i += stride) {
            int k = i / G, gp = i % G, gate = gp / DEC_HP, j = gp % DEC_HP;
            bool ok = (k < DEC_H) && (j < DEC_H);
            int g = gate * DEC_H + j;
            a.d.whh_t_dec[i] = ok ? a.dec_whh[(size_t)g * DEC_H + k] : 0.f;
            a.d.wizc_t[i] = ok ? a.dec_wih[(size_t)g * DEC_IN + EMB + k] : 0.f;
        }
        for (int i = t0; i < DEC_HP; i += stride) a.d.bhn_dec[i] = (i < DEC_H) ? a.dec_bhh[2 * DEC_H + i] : 0.f;
    } else if (task == 6) {             // decoder natural layouts, padded [312][104]
        const int G = 3 * DEC_HP;
        for (int i = t0; i < G * DEC_HP; i += stride) {
            int gp = i / DEC_HP, k = i % DEC_HP, gate = gp / DEC_HP, j = gp % DEC_HP;
            bool ok = (k < DEC_H) && (j < DEC_H);
            int g = gate * DEC_H + j;
            a.d.whh_dec[i] = ok ? a.dec_whh[(size_t)g * DEC_H + k] : 0.f;
            a.d.wizc[i] = ok ? a.dec_wih[(size_t)g * DEC_IN + EMB + k] : 0.f;
        }
    } else if (task == 7) {             // fc padded [VMAX][104]
        for (int i = t0; i < VMAX * DEC_HP; i += stride) {
            int v = i / DEC_HP, j = i % DEC_HP;
            a.d.fc_w[i] = (v < V && j < DEC_H) ? a.fc_w[(size_t)v * DEC_H + j] : 0.f;
        }
        for (int i = t0; i < VMAX; i += stride) a.d.fc_b[i] = (i < V) ? a.fc_b[i] : 0.f;
    }
#ifndef CPG_EMU
    // ---- pre-split operand tiles of the tcgen05 dense layers around the latent code (latent_tc.cu; geometry in latent.h):
    // one thread per 16-byte chunk (8 consecutive K, or 8 consecutive N of an MN-major tile) of the shared-memory image
    else if (task == 8) {               // heads, K-major [208][160]: row n = 2 j (q_mu) | 2 j + 1 (q_logvar); fp16 split
        for (int i = t0; i < LT_F1_ROWS * (LT_F1_K / 8); i += stride) {
            const int kc = i % (LT_F1_K / 8), r = i / (LT_F1_K / 8), j = r >> 1;
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = j < ZD ? ((r & 1) ? a.wlv : a.wmu)[(size_t)j * (2 * ENC_H) + kc * 8 + e] : 0.f;
            uint4 h, l;
            split2h(x[0], x[1], h.x, l.x); split2h(x[2], x[3], h.y, l.y); split2h(x[4], x[5], h.z, l.z); split2h(x[6], x[7], h.w, l.w);
            const size_t off = LT_F1_OFF + (size_t)kc * (LT_F1_ROWS * 16) + (r >> 3) * 128 + (r & 7) * 16;
            *reinterpret_cast<uint4*>(a.d.lat_tiles + off) = h;
            *reinterpret_cast<uint4*>(a.d.lat_tiles + off + LT_F1_TERM) = l;
        }
    } else if (task == 9) {             // W_ih[:,150:], K-major [320][112]: row = padded gate index, K = [z;c] index; fp16 split
        for (int i = t0; i < LT_F2_ROWS * (LT_F2_K / 8); i += stride) {
            const int kc = i % (LT_F2_K / 8), r = i / (LT_F2_K / 8), gate = r / DEC_HP, j = r % DEC_HP;
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int k = kc * 8 + e;
                x[e] = (gate < 3 && j < DEC_H && k < DEC_H) ? a.dec_wih[(size_t)(gate * DEC_H + j) * DEC_IN + EMB + k] : 0.f;
            }
            uint4 h, l;
            split2h(x[0], x[1], h.x, l.x); split2h(x[2], x[3], h.y, l.y); split2h(x[4], x[5], h.z, l.z); split2h(x[6], x[7], h.w, l.w);
            const size_t off = LT_F2_OFF + (size_t)kc * (LT_F2_ROWS * 16) + (r >> 3) * 128 + (r & 7) * 16;
            *reinterpret_cast<uint4*>(a.d.lat_tiles + off) = h;
            *reinterpret_cast<uint4*>(a.d.lat_tiles + off + LT_F2_TERM) = l;
        }
    } else if (task == 10) {            // (W_ih[:,150:])^T, K-major [112][320]: row = [z;c] index, K = padded gate index; bf16 split
        for (int i = t0; i < LT_B1_ROWS * (LT_B1_K / 8); i += stride) {
            const int r = i % LT_B1_ROWS, kc = i / LT_B1_ROWS;
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int g = kc * 8 + e, gate = g / DEC_HP, j = g % DEC_HP;
                x[e] = (gate < 3 && j < DEC_H && r < DEC_H) ? a.dec_wih[(size_t)(gate * DEC_H + j) * DEC_IN + EMB + r] : 0.f;
            }
            uint4 h, l;
            split8(x, h, l);
            const size_t off = LT_B1_OFF + (size_t)kc * (LT_B1_ROWS * 16) + (r >> 3) * 128 + (r & 7) * 16;
            *reinterpret_cast<uint4*>(a.d.lat_tiles + off) = h;
            *reinterpret_cast<uint4*>(a.d.lat_tiles + off + LT_B1_TERM) = l;
        }
    } else if (task == 11) {            // heads^T, MN-major: element (n = hfin column, k = 2 j | 2 j + 1) = W_mu[j][n] | W_logvar[j][n]; bf16 split
        for (int i = t0; i < (LT_B2_N / 8) * LT_B2_K; i += stride) {
            const int nc = i % (LT_B2_N / 8), k = i / (LT_B2_N / 8), j = k >> 1;
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = j < ZD ? ((k & 1) ? a.wlv : a.wmu)[(size_t)j * (2 * ENC_H) + nc * 8 + e] : 0.f;
            uint4 h, l;
            split8(x, h, l);
            const size_t off = LT_B2_OFF + (size_t)nc * (LT_B2_K * 16) + (k >> 3) * 128 + (k & 7) * 16;
            *reinterpret_cast<uint4*>(a.d.lat_tiles + off) = h;
            *reinterpret_cast<uint4*>(a.d.lat_tiles + off + LT_B2_TERM) = l;
        }
    }
#endif
}

void launch_prep_weights(cudaStream_t s, const float* p, const ParamLayout& l, int V, const Derived& d, int part) {
    PrepArgs a;
    a.emb = p + l.off[P_EMB];
    a.enc_wih[0] = p + l.off[P_ENC_WIH_F]; a.enc_whh[0] = p + l.off[P_ENC_WHH_F];
    a.enc_bih[0] = p + l.off[P_ENC_BIH_F]; a.enc_bhh[0] = p + l.off[P_ENC_BHH_F];
    a.enc_wih[1] = p + l.off[P_ENC_WIH_R]; a.enc_whh[1] = p + l.off[P_ENC_WHH_R];
    a.enc_bih[1] = p + l.off[P_ENC_BIH_R]; a.enc_bhh[1] = p + l.off[P_ENC_BHH_R];
    a.dec_wih = p + l.off[P_DEC_WIH]; a.dec_whh = p + l.off[P_DEC_WHH];
    a.dec_bih = p + l.off[P_DEC_BIH]; a.dec_bhh = p + l.off[P_DEC_BHH];
    a.fc_w = p + l.off[P_FC_W]; a.fc_b = p + l.off[P_FC_B];
    a.wmu = p + l.off[P_QMU_W]; a.wlv = p + l.off[P_QLV_W];
    a.d = d;
    a.V = V;
    // 960 warps per table task: ~8 outputs per warp; tasks 8-11 (operand tiles of latent_tc.cu) only where they are used.
    // part 1 = what the encoder recurrence needs (token tables and W_hh^T of both directions), part 2 = the rest
    int tasks[12], n = 0;
    if (part & 1) { tasks[n++] = 0; tasks[n++] = 1; tasks[n++] = 3; tasks[n++] = 4; }
    if (part & 2) {
        tasks[n++] = 2; tasks[n++] = 5; tasks[n++] = 6; tasks[n++] = 7;
        if (d.lat_tiles != nullptr) { tasks[n++] = 8; tasks[n++] = 9; tasks[n++] = 10; tasks[n++] = 11; }
    }
    unsigned long long map = 0;
    for (int i = 0; i < n; ++i) map |= (unsigned long long)tasks[i] << (4 * i);
    CPG_LAUNCH(k_prep_weights, dim3(720, n), 256, 0, s, a, map);      // 5760 warps per task: one token-table output per warp
}

}  // namespace cpg
