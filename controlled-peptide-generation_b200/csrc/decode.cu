// Step-wise decoding from (z, c): batched beam search, greedy and categorical sampling.
//
// Replaces the Python double loop of RNN_VAE.sample_G (models/model.py:225-385), which per step
// calls GRUDecoder.forward_sample (models/decoder.py:86-109) and then, for beam search, runs
// log_softmax + Beam.advance (models/Beam.py:56-105) + _update_hidden (model.py:387-404) ONCE PER
// SAMPLE in Python.  Here a CTA owns 32 decoder rows (6 samples x 5 beams, or 32 samples) for all
// <= 25 steps: W_hh^T stays in shared memory, the input projection is the token table plus the
// per-sample [z;c] projection, the fc + log-softmax is one warp per row (lane = class), and the
// top-5 over 5*V candidates, back-pointers, finished list and hidden-state permutation stay on chip.
//
// Tie rule (torch.topk's CPU order among equal scores is unspecified): larger score first, then
// lower flat index.  Sentinel -1e20 as in Beam.py.
#include "ctx.h"

namespace cpg {
int check_launch(const char* where);

constexpr int DR = 32;                 // rows per CTA
constexpr int BEAM_K = 5;              // beam width supported by the fused kernel
constexpr int BEAM_S = DR / BEAM_K;    // 6 samples per CTA
constexpr int NBEST_MAX = 5;
constexpr int FIN_MAX = BEAM_K * LMAX; // finished-list capacity per sample
constexpr float NEG_SENT = -1e20f;
constexpr int D_NU = DEC_HP / 4, D_NT = D_NU * (DR / 4), D_G = 3 * DEC_HP;
constexpr int FSTR = DEC_HP + 1;

// 3..5: soft sampling (models/model.py:330-341): the softmax of every step is returned and fed back as a SOFT embedding
enum DecodeMode { MODE_BEAM = 0, MODE_GREEDY = 1, MODE_CATEGORICAL = 2, MODE_NONE_SOFTMAX = 3, MODE_GREEDY_SOFTMAX = 4,
                  MODE_CATEGORICAL_SOFTMAX = 5 };

struct DecodeArgs {
    const float* table;     // [V][312]
    const float* rowbias;   // [n][312]  per-sample [z;c] projection
    const float* whh_t;     // [104][312]
    const float* bhn;       // [104]
    const float* zc;        // [n][104]  h0
    const float* fc_w;      // [VMAX][104]
    const float* fc_b;      // [VMAX]
    int n, L, V, n_best;
    float temp; uint64_t seed;
    int* out_tok;           // beam: [n][n_best][L+1] (-1 padded) ; sampling: [n][L+1]
    int* out_len;           // beam: [n][n_best]
    float* out_score;       // beam: [n][n_best]
    int* out_steps;         // sampling: max over samples of steps executed (atomicMax)
    float* out_soft;        // soft sampling: [n][L+1][V] softmax(logits / temp) per step (0 after <eos>), one-hot <start> first
};

template <int R>
__device__ __forceinline__ int dswz(int k, int ty) { return k * R + ((ty ^ ((k >> 2) & (R / 4 - 1))) << 2); }
// scalar element (k, row) of a swizzled [K][R] tile
template <int R>
__device__ __forceinline__ int dswz_elem(int k, int row) { return dswz<R>(k, row >> 2) + (row & 3); }

struct BeamState {                      // per sample, in shared memory
    float score[BEAM_K];
    int tok[BEAM_K];
    int prev[BEAM_K];
    unsigned char ys[LMAX + 1][BEAM_K];     // next_ys history
    unsigned char pk[LMAX][BEAM_K];         // prev_ks history
    float fin_score[FIN_MAX];
    unsigned char fin_t[FIN_MAX], fin_k[FIN_MAX];
    int n_fin, n_steps, eos_top, done;
};

template <int MODE>
__global__ void __launch_bounds__(D_NT)
k_decode(DecodeArgs a) {
    constexpr int HP = DEC_HP, G = D_G, R = DR;
    CPG_DYN_SMEM(float, smem);
    float* Wt = smem;                         // [HP][G]
    float* hA = Wt + HP * G;                  // [HP][R] current hidden (swizzled)
    float* hB = hA + HP * R;                  // [HP][R] new hidden
    float* Fw = hB + HP * R;                  // [VMAX][FSTR]
    float* cand = Fw + VMAX * FSTR;           // [R][VMAX]  log-probs (+ scores for beam)
    int* rtok = reinterpret_cast<int*>(cand + R * VMAX);     // [R] current input token per row
    int* rfin = rtok + R;                                    // [R] sampling: finished flag
    int* misc = rfin + R;                                    // [4]
    BeamState* bs = reinterpret_cast<BeamState*>(misc + 4);  // [BEAM_S] (beam mode only)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = D_NT / 32;          // FULL warps only take part in the warp-collective phases
    const bool full_warp = warp < nwarps;
    const int tx = tid % D_NU, ty = tid / D_NU;
    const int j0 = 4 * tx, r0 = 4 * ty;
    constexpr bool SOFT = MODE >= MODE_NONE_SOFTMAX;
    const int V = a.V, L = a.L;
    const int spc = MODE == MODE_BEAM ? BEAM_S : R;          // samples per CTA
    const int samp0 = blockIdx.x * spc;
    auto row_sample = [&](int row) { return MODE == MODE_BEAM ? samp0 + row / BEAM_K : samp0 + row; };
    auto row_valid = [&](int row) { return (MODE != MODE_BEAM || row < BEAM_S * BEAM_K) && row_sample(row) < a.n; };

    for (int i = tid * 4; i < HP * G; i += D_NT * 4) st4(Wt + i, ld4(a.whh_t + i));
    for (int i = tid; i < VMAX * HP; i += D_NT) Fw[(i / HP) * FSTR + (i % HP)] = a.fc_w[i];
    if (tid < R) {
        rtok[tid] = (MODE == MODE_BEAM) ? ((tid % BEAM_K) == 0 ? START : PAD) : START;   // model.py:272-276
        rfin[tid] = 0;
    }
    if (MODE == MODE_BEAM && tid < BEAM_S) {
        BeamState& b = bs[tid];
        for (int k = 0; k < BEAM_K; ++k) { b.score[k] = 0.f; b.tok[k] = k == 0 ? START : PAD; b.prev[k] = k; b.ys[0][k] = (unsigned char)b.tok[k]; }
        b.n_fin = 0; b.n_steps = 0; b.eos_top = 0;
        b.done = (samp0 + tid < a.n) ? 0 : 1;
    }
    if (tid == 0) misc[0] = 0;
    // h0 = [z;c] for every row of the sample
    float hprev[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int row = r0 + i;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row_valid(row)) v = ld4(a.zc + (size_t)row_sample(row) * HP + j0);
        hprev[i][0] = v.x; hprev[i][1] = v.y; hprev[i][2] = v.z; hprev[i][3] = v.w;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
        st4(hA + dswz<R>(j0 + u, ty), make_float4(hprev[0][u], hprev[1][u], hprev[2][u], hprev[3][u]));
    const float4 bhn4 = ld4(a.bhn + j0);
    const float bhn[4] = {bhn4.x, bhn4.y, bhn4.z, bhn4.w};
    __syncthreads();

    int steps_done = 0;
    for (int s = 0; s < L; ++s) {
        // ---- 1) GRU cell on the 32-row tile (same tiling as k_gru_fwd)
        float gi[4][3][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int row = r0 + i;
            int sm_ = row_valid(row) ? row_sample(row) : 0;
            const float* base = a.table + (size_t)rtok[row] * G + j0;
            const float* rb = a.rowbias + (size_t)sm_ * G + j0;
            if (SOFT && s > 0) {
                // soft embedding of the previous step's (masked) softmax: sum_v p_v T[v]; the table rows carry the input
                // biases, so the missing mass 1 - sum_v p_v (all of it after <eos>) takes the <pad> row = biases only
                float4 acc3[3];
                float rest = 1.0f;
#pragma unroll
                for (int g = 0; g < 3; ++g) acc3[g] = __ldg(reinterpret_cast<const float4*>(rb + g * HP));
                for (int v = 0; v <= V; ++v) {
                    const float p = v < V ? cand[row * VMAX + v] : rest;
                    rest -= p;
                    const float* tv = a.table + (size_t)(v < V ? v : PAD) * G + j0;
#pragma unroll
                    for (int g = 0; g < 3; ++g) {
                        const float4 t4 = __ldg(reinterpret_cast<const float4*>(tv + g * HP));
                        acc3[g].x = fmaf(p, t4.x, acc3[g].x); acc3[g].y = fmaf(p, t4.y, acc3[g].y);
                        acc3[g].z = fmaf(p, t4.z, acc3[g].z); acc3[g].w = fmaf(p, t4.w, acc3[g].w);
                    }
                }
#pragma unroll
                for (int g = 0; g < 3; ++g) { gi[i][g][0] = acc3[g].x; gi[i][g][1] = acc3[g].y; gi[i][g][2] = acc3[g].z; gi[i][g][3] = acc3[g].w; }
                continue;
            }
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                float4 v = __ldg(reinterpret_cast<const float4*>(base + g * HP));
                float4 b = __ldg(reinterpret_cast<const float4*>(rb + g * HP));
                gi[i][g][0] = v.x + b.x; gi[i][g][1] = v.y + b.y; gi[i][g][2] = v.z + b.z; gi[i][g][3] = v.w + b.w;
            }
        }
        float acc[4][3][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int g = 0; g < 3; ++g)
#pragma unroll
                for (int u = 0; u < 4; ++u) acc[i][g][u] = 0.f;
#pragma unroll 4
        for (int k = 0; k < HP; ++k) {
            const float4 hv = ld4(hA + dswz<R>(k, ty));
            const float h4[4] = {hv.x, hv.y, hv.z, hv.w};
            const float* wrow = Wt + k * G + j0;
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                const float4 wv = ld4(wrow + g * HP);
                const float w4[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int u = 0; u < 4; ++u) acc[i][g][u] = fmaf(h4[i], w4[u], acc[i][g][u]);
            }
        }
        float hnew[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float rr = sigmoidf_acc(gi[i][0][u] + acc[i][0][u]);
                float zz = sigmoidf_acc(gi[i][1][u] + acc[i][1][u]);
                float nn = tanhf(gi[i][2][u] + rr * (acc[i][2][u] + bhn[u]));
                hnew[i][u] = (1.0f - zz) * nn + zz * hprev[i][u];
            }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            st4(hB + dswz<R>(j0 + u, ty), make_float4(hnew[0][u], hnew[1][u], hnew[2][u], hnew[3][u]));
        __syncthreads();

        // ---- 2) fc + log-softmax, one warp per row, lane = class
        for (int row = warp; full_warp && row < R; row += nwarps) {
            float logit = -INFINITY;
            if (lane < V) {
                float sacc = 0.f;
                const float* w = Fw + lane * FSTR;
#pragma unroll 6
                for (int j = 0; j < DEC_H; ++j) sacc = fmaf(hB[dswz_elem<R>(j, row)], w[j], sacc);
                logit = sacc + a.fc_b[lane];
            }
            if (MODE == MODE_BEAM) {
                float mx = warp_max(logit);
                float e = lane < V ? expf(logit - mx) : 0.f;
                float se = warp_sum(e);
                float lp = (logit - mx) - logf(se);                     // F.log_softmax (model.py:322)
                if (lane == START) lp = NEG_SENT;                       // Beam.py:69
                const BeamState& b = bs[row / BEAM_K < BEAM_S ? row / BEAM_K : 0];
                int k = row % BEAM_K;
                float v = lp;
                if (b.n_steps > 0) {
                    v = lp + b.score[k];                                 // Beam.py:71-72
                    if (b.tok[k] == EOS) v = NEG_SENT;                   // Beam.py:74-75
                } else if (k != 0) {
                    v = -INFINITY;                                       // step 0: beam 0 only (Beam.py:77)
                }
                cand[row * VMAX + lane] = lane < V ? v : -INFINITY;
            } else {
                int nxt;
                float psoft = 0.f;
                if (SOFT) {                                              // F.softmax(logits / temp, dim=1)
                    const float sc = logit / a.temp;
                    const float mx = warp_max(sc);
                    const float e = lane < V ? expf(sc - mx) : 0.f;
                    psoft = e / warp_sum(e);
                }
                if (MODE == MODE_NONE_SOFTMAX) {
                    nxt = rtok[row];                                     // the reference never updates sampleIx in this mode
                } else if (MODE == MODE_GREEDY || MODE == MODE_GREEDY_SOFTMAX) {
                    // torch.argmax: first maximal index
                    float best = logit; int bi = lane;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        float ov = __shfl_xor_sync(0xffffffffu, best, o);
                        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
                    }
                    nxt = bi;
                } else {
                    // Categorical(logits / temp): inverse CDF on the softmax with a Philox uniform
                    float sc = logit / a.temp;
                    float mx = warp_max(sc);
                    float e = lane < V ? expf(sc - mx) : 0.f;
                    float cum = e;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        float t = __shfl_up_sync(0xffffffffu, cum, o);
                        if (lane >= o) cum += t;
                    }
                    float total = __shfl_sync(0xffffffffu, cum, 31);
                    uint32_t rnd[4];
                    Philox::gen(a.seed, (uint64_t)row_sample(row), (uint32_t)s, rnd);
                    float u = u32_to_unit_open(rnd[0]) * total;
                    unsigned m = __ballot_sync(0xffffffffu, lane < V && cum >= u);
                    nxt = m ? (__ffs((int)m) - 1) : (V - 1);
                }
                {
                    const int fin = rfin[row];
                    if (fin) nxt = PAD;                                  // model.py:349 masked_fill_(finished, PAD)
                    const int fin_now = fin || nxt == EOS;               // model.py:350
                    if (SOFT) {
                        if (fin_now) psoft = 0.f;                        // model.py:354: zeroed from the <eos> step on
                        cand[row * VMAX + lane] = lane < V ? psoft : 0.f;
                        if (lane < V && row_valid(row)) a.out_soft[((size_t)row_sample(row) * (L + 1) + s + 1) * V + lane] = psoft;
                    }
                    __syncwarp();
                    if (lane == 0) {
                        rfin[row] = fin_now;
                        rtok[row] = nxt;
                        if (row_valid(row)) a.out_tok[(size_t)row_sample(row) * (L + 1) + s + 1] = nxt;
                    }
                }
            }
        }
        __syncthreads();

        if (MODE == MODE_BEAM) {
            // ---- 3) per-sample top-K over K*V candidates, Beam.advance bookkeeping (one warp per sample)
            for (int sl = warp; full_warp && sl < BEAM_S; sl += nwarps) {
                BeamState& b = bs[sl];
                if (b.done) continue;                                   // warp-uniform
                const int ncand = BEAM_K * VMAX;                         // padded classes hold -inf
                float sel_v[BEAM_K]; int sel_i[BEAM_K];
                unsigned taken = 0;                                      // per-lane bitmask over its slots
                for (int pick = 0; pick < BEAM_K; ++pick) {
                    float best = -INFINITY; int bi = 0x7fffffff;
#pragma unroll
                    for (int q = 0; q < ncand / 32; ++q) {
                        int idx = q * 32 + lane;                         // idx = k * VMAX + v
                        int k = idx / VMAX, v = idx % VMAX;
                        float val = cand[(sl * BEAM_K + k) * VMAX + v];
                        int flat = k * V + v;                            // reference flattening (K x V)
                        if (v >= V || ((taken >> q) & 1u)) continue;
                        if (val > best || (val == best && flat < bi)) { best = val; bi = flat; }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        float ov = __shfl_xor_sync(0xffffffffu, best, o);
                        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
                    }
                    sel_v[pick] = best; sel_i[pick] = bi;
                    int k = bi / V, v = bi % V;
                    int idx = k * VMAX + v;
                    if ((idx & 31) == lane) taken |= 1u << (idx >> 5);
                }
                if (lane == 0) {
                    int t = b.n_steps;                                   // this advance produces next_ys[t+1]
                    for (int k = 0; k < BEAM_K; ++k) {
                        int pk = sel_i[k] / V, y = sel_i[k] - pk * V;
                        b.score[k] = sel_v[k]; b.prev[k] = pk; b.tok[k] = y;
                        b.pk[t][k] = (unsigned char)pk;
                        b.ys[t + 1][k] = (unsigned char)y;
                    }
                    for (int k = 0; k < BEAM_K; ++k)                     // Beam.py:95-98
                        if (b.tok[k] == EOS && b.n_fin < FIN_MAX) {
                            b.fin_score[b.n_fin] = b.score[k];
                            b.fin_t[b.n_fin] = (unsigned char)(t + 1);
                            b.fin_k[b.n_fin] = (unsigned char)k;
                            b.n_fin++;
                        }
                    if (b.tok[0] == EOS) b.eos_top = 1;                  // Beam.py:101-103
                    b.n_steps = t + 1;
                    if (b.eos_top && b.n_fin >= a.n_best) b.done = 1;    // Beam.py:107-108
                }
            }
            __syncthreads();
            // ---- 4) hidden-state permutation by back-pointers (model.py:325,387-404) and next tokens
            for (int idx = tid; idx < HP * R; idx += D_NT) {
                int j = idx / R, row = idx % R;
                int src = row;
                if (row < BEAM_S * BEAM_K) {
                    const BeamState& b = bs[row / BEAM_K];
                    src = (row / BEAM_K) * BEAM_K + b.prev[row % BEAM_K];
                }
                hA[dswz_elem<R>(j, row)] = hB[dswz_elem<R>(j, src)];
            }
            if (tid < BEAM_S * BEAM_K) rtok[tid] = bs[tid / BEAM_K].tok[tid % BEAM_K];
            if (tid == 0) {
                int all = 1;
                for (int sl = 0; sl < BEAM_S; ++sl) all &= bs[sl].done;
                misc[0] = all;
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float4 v = ld4(hA + dswz<R>(j0 + u, ty));
                hprev[0][u] = v.x; hprev[1][u] = v.y; hprev[2][u] = v.z; hprev[3][u] = v.w;
            }
            if (misc[0]) break;
        } else {
            // swap roles of the two hidden buffers by copying registers (h' of this thread's tile)
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int u = 0; u < 4; ++u) hprev[i][u] = hnew[i][u];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                st4(hA + dswz<R>(j0 + u, ty), make_float4(hnew[0][u], hnew[1][u], hnew[2][u], hnew[3][u]));
            if (tid == 0) {
                int all = 1;
                for (int r = 0; r < R; ++r) if (row_valid(r)) all &= rfin[r];
                misc[0] = all;
            }
            steps_done = s + 1;
            __syncthreads();
            if (misc[0]) break;
        }
    }

    if (MODE != MODE_BEAM) {
        // rows that finished early keep <pad> in the remaining columns
        for (int row = warp; full_warp && row < R; row += nwarps) {
            if (!row_valid(row)) continue;
            int* o = a.out_tok + (size_t)row_sample(row) * (L + 1);
            if (lane == 0) o[0] = START;
            for (int c = steps_done + 1 + lane; c <= L; c += 32) o[c] = PAD;
            if (SOFT) {
                float* os = a.out_soft + (size_t)row_sample(row) * (L + 1) * V;
                if (lane < V) os[lane] = lane == START ? 1.f : 0.f;      // onehot_embed(<start>) (model.py:293)
                for (int c = (steps_done + 1) * V + lane; c < (L + 1) * V; c += 32) os[c] = 0.f;
            }
        }
        if (tid == 0 && a.out_steps != nullptr) atomicMax(a.out_steps, steps_done);
        return;
    }

    // ---- 5) sort_finished(minimum=n_best) + get_hyp (Beam.py:110-132), one thread per sample
    if (tid < BEAM_S && samp0 + tid < a.n) {
        BeamState& b = bs[tid];
        int i = 0;
        while (b.n_fin < a.n_best && b.n_fin < FIN_MAX) {            // pad from the live beam in rank order
            b.fin_score[b.n_fin] = b.score[i];
            b.fin_t[b.n_fin] = (unsigned char)b.n_steps;
            b.fin_k[b.n_fin] = (unsigned char)i;
            b.n_fin++; i++;
        }
        const int samp = samp0 + tid;
        unsigned char used[FIN_MAX];
        for (int q = 0; q < b.n_fin; ++q) used[q] = 0;
        for (int rank = 0; rank < a.n_best; ++rank) {                 // stable selection by descending score
            int best = -1;
            for (int q = 0; q < b.n_fin; ++q)
                if (!used[q] && (best < 0 || b.fin_score[q] > b.fin_score[best])) best = q;
            used[best] = 1;
            int t = b.fin_t[best], k = b.fin_k[best];
            int* o = a.out_tok + ((size_t)samp * a.n_best + rank) * (L + 1);
            for (int c = 0; c <= L; ++c) o[c] = -1;
            for (int pos = t; pos >= 0; --pos) {                      // walk the back-pointers
                o[pos] = b.ys[pos][k];
                if (pos > 0) k = b.pk[pos - 1][k];
            }
            a.out_len[samp * a.n_best + rank] = t + 1;
            a.out_score[samp * a.n_best + rank] = b.fin_score[best];
        }
    }
}

static size_t decode_smem_bytes() {
    size_t f = (size_t)DEC_HP * D_G + 2 * (size_t)DEC_HP * DR + (size_t)VMAX * FSTR + (size_t)DR * VMAX;
    size_t bytes = f * sizeof(float) + (2 * DR + 4) * sizeof(int);
    bytes = align_up(bytes, 16) + BEAM_S * sizeof(BeamState);
    return bytes;
}

}  // namespace cpg

using namespace cpg;

extern "C" {

// shared front end: derived weights, [z;c] rows and their input projection
static int decode_prepare(cpg_ctx* ctx, cudaStream_t s, const float* params, int V, int n, int L, const float* z,
                          const float* c, DecodeArgs& a) {
    if (!ctx || !params || !z || !c) { set_error("decode: null argument"); return CPG_EINVAL; }
    if (V < 4 || V > VMAX || n < 1 || L < 1 || L > LMAX) { set_error("decode: bad shape"); return CPG_EINVAL; }
    int rc = ensure_workspace(ctx, n, ctx->ws.L >= 2 ? ctx->ws.L : 2, V, ctx->ws.R > 0 ? ctx->ws.R : 500, s);
    if (rc) return rc;
    Workspace& w = ctx->ws;
    ParamLayout lay = make_layout(V);
    launch_prep_weights(s, params, lay, V, w.d);
    launch_make_zc(s, z, c, n, w.zc);
    launch_sgemm(s, n, 3 * DEC_HP, DEC_HP, 1.f, w.zc, DEC_HP, 1, w.d.wizc_t, 3 * DEC_HP, 1, 0.f, w.rowbias, 3 * DEC_HP,
                 nullptr, 1, nullptr);
    memset(&a, 0, sizeof(a));
    a.table = w.d.t_dec; a.rowbias = w.rowbias; a.whh_t = w.d.whh_t_dec; a.bhn = w.d.bhn_dec; a.zc = w.zc;
    a.fc_w = w.d.fc_w; a.fc_b = w.d.fc_b;
    a.n = n; a.L = L; a.V = V;
    ctx->have_stash = false;
    return CPG_OK;
}

int cpg_beam_decode(cpg_ctx* ctx, cpg_stream stream, const float* params, int V, int n, int L, const float* z,
                    const float* c, int beam_size, int n_best, int* out_tokens, int* out_len, float* out_score) {
    if (beam_size != BEAM_K) { set_error("cpg_beam_decode: the fused kernel supports beam_size == 5 (cfg.evals default)"); return CPG_EINVAL; }
    if (n_best < 1 || n_best > NBEST_MAX) { set_error("cpg_beam_decode: n_best must be in [1, 5]"); return CPG_EINVAL; }
    if (!out_tokens || !out_len || !out_score) { set_error("cpg_beam_decode: null output"); return CPG_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    DecodeArgs a;
    int rc = decode_prepare(ctx, s, params, V, n, L, z, c, a);
    if (rc) return rc;
    a.n_best = n_best; a.out_tok = out_tokens; a.out_len = out_len; a.out_score = out_score;
    auto kfn = k_decode<MODE_BEAM>;
    size_t smem = decode_smem_bytes();
    CPG_SET_MAX_SMEM(kfn, smem);
    CPG_LAUNCH_NAMED("k_decode_beam", kfn, ceil_div(n, BEAM_S), D_NT, smem, s, a);
    return check_launch("cpg_beam_decode");
}

int cpg_sample_decode(cpg_ctx* ctx, cpg_stream stream, const float* params, int V, int n, int L, const float* z,
                      const float* c, int mode, float temp, uint64_t seed, int* out_tokens, int* out_steps) {
    if (mode != MODE_GREEDY && mode != MODE_CATEGORICAL) { set_error("cpg_sample_decode: mode must be 1 (greedy) or 2 (categorical)"); return CPG_EINVAL; }
    if (!out_tokens || !out_steps) { set_error("cpg_sample_decode: null output"); return CPG_EINVAL; }
    if (!(temp > 0.f)) { set_error("cpg_sample_decode: temp must be > 0"); return CPG_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    DecodeArgs a;
    int rc = decode_prepare(ctx, s, params, V, n, L, z, c, a);
    if (rc) return rc;
    a.temp = temp; a.seed = seed; a.out_tok = out_tokens; a.out_steps = out_steps;
    cudaMemsetAsync(out_steps, 0, sizeof(int), s);
    size_t smem = decode_smem_bytes();
    if (mode == MODE_GREEDY) {
        auto kfn = k_decode<MODE_GREEDY>;
        CPG_SET_MAX_SMEM(kfn, smem);
        CPG_LAUNCH_NAMED("k_decode_greedy", kfn, ceil_div(n, DR), D_NT, smem, s, a);
    } else {
        auto kfn = k_decode<MODE_CATEGORICAL>;
        CPG_SET_MAX_SMEM(kfn, smem);
        CPG_LAUNCH_NAMED("k_decode_categorical", kfn, ceil_div(n, DR), D_NT, smem, s, a);
    }
    return check_launch("cpg_sample_decode");
}

int cpg_soft_decode(cpg_ctx* ctx, cpg_stream stream, const float* params, int V, int n, int L, const float* z, const float* c,
                    int mode, float temp, uint64_t seed, int* out_tokens, float* out_soft, int* out_steps) {
    if (mode < MODE_NONE_SOFTMAX || mode > MODE_CATEGORICAL_SOFTMAX) { set_error("cpg_soft_decode: mode must be 3 (none_softmax), 4 (greedy_softmax) or 5 (categorical_softmax)"); return CPG_EINVAL; }
    if (!out_tokens || !out_soft || !out_steps) { set_error("cpg_soft_decode: null output"); return CPG_EINVAL; }
    if (!(temp > 0.f)) { set_error("cpg_soft_decode: temp must be > 0"); return CPG_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    DecodeArgs a;
    int rc = decode_prepare(ctx, s, params, V, n, L, z, c, a);
    if (rc) return rc;
    a.temp = temp; a.seed = seed; a.out_tok = out_tokens; a.out_steps = out_steps; a.out_soft = out_soft;
#ifdef CPG_EMU
    *out_steps = 0;
#else
    cudaMemsetAsync(out_steps, 0, sizeof(int), s);
#endif
    size_t smem = decode_smem_bytes();
    if (mode == MODE_NONE_SOFTMAX) {
        auto kfn = k_decode<MODE_NONE_SOFTMAX>;
        CPG_SET_MAX_SMEM(kfn, smem);
        CPG_LAUNCH_NAMED("k_decode_none_softmax", kfn, ceil_div(n, DR), D_NT, smem, s, a);
    } else if (mode == MODE_GREEDY_SOFTMAX) {
        auto kfn = k_decode<MODE_GREEDY_SOFTMAX>;
        CPG_SET_MAX_SMEM(kfn, smem);
        CPG_LAUNCH_NAMED("k_decode_greedy_softmax", kfn, ceil_div(n, DR), D_NT, smem, s, a);
    } else {
        auto kfn = k_decode<MODE_CATEGORICAL_SOFTMAX>;
        CPG_SET_MAX_SMEM(kfn, smem);
        CPG_LAUNCH_NAMED("k_decode_categorical_softmax", kfn, ceil_div(n, DR), D_NT, smem, s, a);
    }
    return check_launch("cpg_soft_decode");
}

}  // extern "C"
