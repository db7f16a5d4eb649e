// BPTT of the three GRUs with the weight-gradient contractions fused in (train_vae.py:39-40: loss.backward()
// through nn.GRU of models/encoder.py:25-30,42 and models/decoder.py:40,77).
//
// The first tcgen05 BPTT (gru_tc.cu: k_gru_bwd_tc) writes the four gate-gradient planes dg = (dr, dz, dn, dhn) to
// HBM for a separate weight-gradient kernel (wgrad_tc.cu) to read back: 8x the bytes the path needs.  Here dg never
// leaves the SM.  Per step s (reverse order) one CTA owns NB batch rows (encoder 64, decoder 32) of one direction:
//
//   dh^T[H x NB]        = W_hh^T[H x 3H] . (dr, dz, dhn)^T[3H x NB]            (the recurrence; critical path)
//   dW_hh[3H x H]      += (dr, dz, dhn)^T[3H x NB] . h_prev[NB x H]             (K = the CTA's batch rows)
//   dT[4H x 32]        += (dr, dz, dhn, dn)^T[4H x NB] . onehot(token)[NB x 32]  (token-table gradient: embedding,
//                                                                                W_ih, b_ih, b_hh all follow from it)
//
// all three on the tensor cores (kind::f16, bf16 operands, fp32 accumulation in TMEM) from ONE operand tile X that
// the gate-derivative math writes once per step: X[b][g] as split bf16 (x = x1 + x2; products x1 w1 + x1 w2 + x2 w1:
// 2^-16 relative, fp32 grade).  The tile is K-major for the recurrence (B operand, K = gate index) and, read through
// an MN-major descriptor, the A operand of the two gradient contractions (M = gate index, K = batch row).
// The dW_hh / dT accumulators (encoder 2 x 80 + 3 x 32, decoder 3 x 112 + 4 x 32 columns) stay in tensor memory
// for all L steps; W_hh^T therefore lives in SHARED memory (both bf16 terms) as the A operand of the recurrence,
// and the batch rows of a CTA form one wide chain (N = NB) so that its 4 KB read per MMA is amortised.
// At the end every CTA writes its partial dW_hh / dT; the ordered reductions of wgrad.cu sum them (no atomics).
//
// Warp roles: NWE epilogue warps (TMEM read-out of dh -> P[b][j] in shared memory -> gate derivatives for
// (row, 4 units) items -> operand tiles) + one MMA-issue warp.  Barriers: bar_x (tiles of step i written),
// bar_d (dh MMAs of step i done), bar_w (gradient MMAs of step i done: the tiles may be overwritten).
#include "ctx.h"
#ifndef CPG_EMU
#include "tc_gru.cuh"

namespace cpg {
int check_launch(const char* where);

namespace {

// D[tmem] (+)= A[smem desc] * B[smem desc]: kind::f16 (bf16 operands), fp32 accumulate
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// kind::f16 / bf16 operands / fp32 accumulate; a_mn, b_mn: 1 = MN-major operand
__host__ __device__ constexpr uint32_t make_idesc_bf16_mn(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int HP_, int NB_, int NWE_, bool DEC_>
struct FCfg {
    static constexpr int HP = HP_, NB = NB_, NWE = NWE_;
    static constexpr bool DEC = DEC_;
    static constexpr int K3 = 3 * HP;
    static constexpr int KPAD = (K3 + 15) / 16 * 16;          // K extent of the recurrence MMAs: 240 | 320
    static constexpr int KSTEPS = KPAD / 16;
    static constexpr int KX = KPAD + HP;                        // gate index extent of X incl. the dn plane: 320 | 424
    static constexpr int KXA = (KX + 63) / 64 * 64;             // allocated: 320 | 448
    static constexpr int KC = KXA / 8;
    static constexpr int NQ = HP / 4;
    static constexpr int NITEMS = NB * NQ;
    static constexpr int NT_E = NWE * 32;
    static constexpr int ITEMS = (NITEMS + NT_E - 1) / NT_E;
    static constexpr int NTHREADS = NT_E + 32;
    static constexpr int X_LBO = (NB / 8) * 128 + 16;           // stride of 8-gate chunks (padded: conflict-free 8-byte stores)
    static constexpr int X_TERM = KC * X_LBO;
    static constexpr int NH = (HP + 15) / 16 * 16;              // N of the dW_hh MMAs: 80 | 112
    static constexpr int HC_HI = (NH + 32) / 8, HC_LO = NH / 8; // chunks of [h_prev | onehot] (hi) and h_prev (lo)
    static constexpr int H_LBO = X_LBO;
    static constexpr int W_ROWS = (HP + 7) / 8 * 8;
    static constexpr int W_LBO = (W_ROWS / 8) * 128;            // K-adjacent core matrices of W_hh^T (rows j, K = gate)
    static constexpr int W_TERM = (KPAD / 8) * W_LBO;
    static constexpr int NT_W = (K3 + 127) / 128;               // M tiles of dW_hh: 2 | 3
    static constexpr int NT_T = (KX + 127) / 128;               // M tiles of dT:    3 | 4
    static constexpr int KS_B = NB / 16;                        // K steps (batch rows) of the gradient MMAs
    // tensor memory columns
    static constexpr int COL_DH = 0, COL_W = NB, COL_T = COL_W + NT_W * NH, COL_END = COL_T + NT_T * 32;
    static_assert(COL_END <= 512, "TMEM columns");
    static_assert(NB % 16 == 0 && HP % 8 == 0 && NWE >= 4, "geometry");
    // shared memory (bytes): W (2 terms) | X (2 terms) | Hx hi | Hx lo | P | tokens
    static constexpr int OFF_W = 0;
    static constexpr int OFF_X = OFF_W + 2 * W_TERM;
    static constexpr int OFF_H = OFF_X + 2 * X_TERM;
    static constexpr int OFF_HL = OFF_H + HC_HI * H_LBO;
    static constexpr int OFF_P = OFF_HL + HC_LO * H_LBO;
    static constexpr int OFF_TOK = OFF_P + NB * HP * 4;
    static size_t smem_bytes(int L) { return (size_t)OFF_TOK + (size_t)((NB * L + 15) & ~15) + 64; }
};

struct FArgs {
    const float* whh[2];       // [3*HP][HP] natural (zero padded for the decoder)
    const float* hs[2];        // [B][L][HP] by step
    const float* gates[2];     // [ceil(B/32)][L][4][32][HP] tiled gate stash of gru_tc.cu's forward
    const uint8_t* tok;        // [B][L] tokens that fed this GRU
    const float* h0;           // decoder: [B][HP] (null = zeros)
    const float* dh_out;       // decoder: [B][L][HP]
    const float* dh_fin;       // encoder: [B][2*HP]
    float* dh0;                // decoder: [B][HP]
    float* drow;               // decoder: [B][3*HP]
    float* part_w[2];          // [gridDim.x][3*HP][HP] per direction
    float* part_t[2];          // [gridDim.x][V][4*HP]
    int B, L, V;
};

template <class C>
__global__ void __launch_bounds__(C::NTHREADS, 1)
k_gru_bwd_fused(FArgs a) {
    constexpr int HP = C::HP, NB = C::NB, NQ = C::NQ, K3 = C::K3, KPAD = C::KPAD;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* Wb = smem + C::OFF_W;
    unsigned char* Xb = smem + C::OFF_X;
    unsigned char* Hh = smem + C::OFF_H;
    unsigned char* Hl = smem + C::OFF_HL;
    float* P = reinterpret_cast<float*>(smem + C::OFF_P);
    uint8_t* toks = smem + C::OFF_TOK;
    __shared__ __align__(8) uint64_t bar_x, bar_d, bar_w;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dir = blockIdx.y;
    const int row0 = blockIdx.x * NB;
    const int B = a.B, L = a.L;

    // ---- one-time setup
    // operand tiles zeroed: K padding of X, the unused one-hot / padding columns of Hx
    for (int i = tid; i < (C::OFF_P - C::OFF_X) / 16; i += C::NTHREADS) reinterpret_cast<uint4*>(Xb)[i] = make_uint4(0, 0, 0, 0);
    // W_hh^T as the K-major A operand: A[j][k] = W_hh[k][j], 16 bytes = 8 consecutive k of one j
    {
        const float* whh = dir ? a.whh[1] : a.whh[0];
        for (int idx = tid; idx < (KPAD / 8) * C::W_ROWS; idx += C::NTHREADS) {
            const int kc = idx / C::W_ROWS, j = idx % C::W_ROWS;
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int k = kc * 8 + e;
                x[e] = (k < K3 && j < HP) ? __ldg(whh + (size_t)k * HP + j) : 0.f;
            }
            uint4 hi, lo;
            split8(x, hi, lo);
            const int off = kc * C::W_LBO + (j >> 3) * 128 + (j & 7) * 16;
            *reinterpret_cast<uint4*>(Wb + off) = hi;
            *reinterpret_cast<uint4*>(Wb + C::W_TERM + off) = lo;
        }
    }
    for (int i = tid; i < NB * L; i += C::NTHREADS) {
        const int row = min(row0 + i / L, B - 1);
        toks[i] = a.tok[(size_t)row * L + i % L];
    }
    if (warp == C::NWE) {
        if (lane == 0) {
            tc::mbar_init(&bar_x, C::NT_E);
            tc::mbar_init(&bar_d, 1);
            tc::mbar_init(&bar_w, 1);
            tc::fence_barrier_init();
        }
        __syncwarp();
        tc::tmem_alloc<512>(&tmem_slot);
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == C::NWE) {
        // ---------------- MMA issuer: iteration i handles step s = L-1-i
        constexpr uint32_t id_dh = make_idesc_bf16_mn(128, NB, 0, 0);
        constexpr uint32_t id_w = make_idesc_bf16_mn(128, C::NH, 1, 1);
        constexpr uint32_t id_t = make_idesc_bf16_mn(128, 32, 1, 1);
        const uint32_t w0 = tc::smem_u32(Wb), x0 = tc::smem_u32(Xb), hh0 = tc::smem_u32(Hh), hl0 = tc::smem_u32(Hl);
        for (int i = 0; i < L; ++i) {
            tc::mbar_wait(&bar_x, i & 1);                 // X / Hx of this step are in shared memory
            tc::tc_fence_after();
            if (elect_one()) {
                // recurrence: dh^T = W_hh^T . (dr, dz, dhn)^T   (A, B K-major)
                uint32_t acc = 0;
#pragma unroll
                for (int p = 0; p < 3; ++p) {
#pragma unroll
                    for (int ks = 0; ks < C::KSTEPS; ++ks) {
                        const uint64_t da = tc::make_smem_desc(w0 + WS[p] * C::W_TERM + ks * 2 * C::W_LBO, C::W_LBO, 128, 0);
                        const uint64_t db = tc::make_smem_desc(x0 + XS[p] * C::X_TERM + ks * 2 * C::X_LBO, C::X_LBO, 128, 0);
                        umma_bf16_ss(tmem + C::COL_DH, da, db, id_dh, acc);
                        acc = 1;
                    }
                }
                tc::umma_commit(&bar_d);
                // weight gradients: A = X read MN-major (M = gate index, K = batch row), B = [h_prev | onehot] MN-major
                const uint32_t accw = i > 0 ? 1u : 0u;
#pragma unroll
                for (int t = 0; t < C::NT_W; ++t) {
#pragma unroll
                    for (int p = 0; p < 3; ++p) {
#pragma unroll
                        for (int ks = 0; ks < C::KS_B; ++ks) {
                            const uint64_t da = tc::make_smem_desc(x0 + XS[p] * C::X_TERM + t * 16 * C::X_LBO + ks * 256, 128, C::X_LBO, 0);
                            const uint64_t db = tc::make_smem_desc((WS[p] ? hl0 : hh0) + ks * 256, 128, C::H_LBO, 0);
                            umma_bf16_ss(tmem + C::COL_W + t * C::NH, da, db, id_w, (p > 0 || ks > 0) ? 1u : accw);
                        }
                    }
                }
#pragma unroll
                for (int t = 0; t < C::NT_T; ++t) {
#pragma unroll
                    for (int p = 0; p < 2; ++p) {                  // the one-hot operand is exact: (x1 + x2) . onehot
#pragma unroll
                        for (int ks = 0; ks < C::KS_B; ++ks) {
                            const uint64_t da = tc::make_smem_desc(x0 + p * C::X_TERM + t * 16 * C::X_LBO + ks * 256, 128, C::X_LBO, 0);
                            const uint64_t db = tc::make_smem_desc(hh0 + (C::NH / 8) * C::H_LBO + ks * 256, 128, C::H_LBO, 0);
                            umma_bf16_ss(tmem + C::COL_T + t * 32, da, db, id_t, (p > 0 || ks > 0) ? 1u : accw);
                        }
                    }
                }
                tc::umma_commit(&bar_w);
            }
            __syncwarp();
        }
    } else {
        // ---------------- epilogue
        unsigned char* X0 = Xb;
        unsigned char* X1 = Xb + C::X_TERM;
        const float* hs_g = dir ? a.hs[1] : a.hs[0];
        const float* gates_g = dir ? a.gates[1] : a.gates[0];

        int ib[C::ITEMS], ij[C::ITEMS];
        float carry[C::ITEMS][4];
        float rs[C::DEC ? C::ITEMS : 1][3][4];
#pragma unroll
        for (int it = 0; it < C::ITEMS; ++it) {
            const int idx = tid + it * C::NT_E;
            const bool valid = idx < C::NITEMS;
            ib[it] = valid ? idx / NQ : -1;
            ij[it] = valid ? (idx % NQ) * 4 : 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) carry[it][e] = 0.f;
            if (C::DEC) {
#pragma unroll
                for (int g = 0; g < 3; ++g)
#pragma unroll
                    for (int e = 0; e < 4; ++e) rs[C::DEC ? it : 0][g][e] = 0.f;
            }
        }
        // prefetch registers for the step about to be processed: gate planes (r, z, n, hn), h_prev, dh_out
        float4 pg[C::ITEMS][4], ph[C::ITEMS], pd[C::ITEMS];
        auto prefetch = [&](int s) {
#pragma unroll
            for (int it = 0; it < C::ITEMS; ++it) {
                if (ib[it] < 0) continue;
                const int j0 = ij[it];
                const int row = min(row0 + ib[it], B - 1);
                const size_t bs = (size_t)row * L + s;
                const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
                const float* g = gates_g + gate_stash_offset<HP>(row, s, L) + j0;
#pragma unroll
                for (int pl = 0; pl < 4; ++pl) pg[it][pl] = ld_stream4(g + pl * 32 * HP);
                if (s > 0) ph[it] = ld_stream4(hs_g + (bs - 1) * HP + j0);
                else ph[it] = (C::DEC && a.h0 != nullptr) ? ldg4(a.h0 + (size_t)row * HP + j0) : zero;
                if (C::DEC) pd[it] = ld_stream4(a.dh_out + bs * HP + j0);
                else pd[it] = (s == L - 1) ? ldg4(a.dh_fin + (size_t)row * (2 * HP) + dir * HP + j0) : zero;
            }
        };
        prefetch(L - 1);

        for (int i = 0; i <= L; ++i) {
            const int s = L - 1 - i;                             // i == L: only collects the last contraction (dh0)
            if (i > 0) {
                tc::mbar_wait(&bar_d, (i - 1) & 1);
                tc::tc_fence_after();
                // dh accumulator: lane = hidden unit j, NB batch columns; 16-column slices dealt over the warps of a quadrant
                const int q = warp & 3;
                const int slice = warp >> 2;
                if (q * 32 < HP && slice < NB / 16) {
                    float v[16];
                    tmem_ld_cols<16>(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(C::COL_DH + slice * 16), v);
                    const int j = q * 32 + lane;
                    if (j < HP) {
#pragma unroll
                        for (int c = 0; c < 16; ++c) P[(slice * 16 + c) * HP + j] = v[c];
                    }
                }
                tc::tc_fence_before();
                group_bar_sync(1, C::NT_E);
            }
            float o_r[C::ITEMS][4], o_z[C::ITEMS][4], o_n[C::ITEMS][4], o_hn[C::ITEMS][4];
#pragma unroll
            for (int it = 0; it < C::ITEMS; ++it) {
                const int b = ib[it], j0 = ij[it];
                if (b < 0) continue;
                const int row = row0 + b;
                float dh[4] = {carry[it][0], carry[it][1], carry[it][2], carry[it][3]};
                if (i > 0) {
                    const float4 p4 = ld4(P + b * HP + j0);
                    dh[0] += p4.x; dh[1] += p4.y; dh[2] += p4.z; dh[3] += p4.w;
                }
                if (i == L) {
                    if (C::DEC && row < B) {
                        if (a.dh0 != nullptr) st4(a.dh0 + (size_t)row * HP + j0, make_float4(dh[0], dh[1], dh[2], dh[3]));
                        if (a.drow != nullptr) {
#pragma unroll
                            for (int g = 0; g < 3; ++g)
                                st4(a.drow + (size_t)row * K3 + g * HP + j0,
                                    make_float4(rs[C::DEC ? it : 0][g][0], rs[C::DEC ? it : 0][g][1],
                                                rs[C::DEC ? it : 0][g][2], rs[C::DEC ? it : 0][g][3]));
                        }
                    }
                    continue;
                }
                const float r4[4] = {pg[it][0].x, pg[it][0].y, pg[it][0].z, pg[it][0].w};
                const float z4[4] = {pg[it][1].x, pg[it][1].y, pg[it][1].z, pg[it][1].w};
                const float n4[4] = {pg[it][2].x, pg[it][2].y, pg[it][2].z, pg[it][2].w};
                const float hn4[4] = {pg[it][3].x, pg[it][3].y, pg[it][3].z, pg[it][3].w};
                const float hp4[4] = {ph[it].x, ph[it].y, ph[it].z, ph[it].w};
                const float do4[4] = {pd[it].x, pd[it].y, pd[it].z, pd[it].w};
                const float live = row < B ? 1.0f : 0.0f;        // rows past the batch must add nothing to dW / dT
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float dht = (dh[e] + do4[e]) * live;
                    const float dn = dht * (1.0f - z4[e]);
                    const float dz = dht * (hp4[e] - n4[e]);
                    carry[it][e] = dht * z4[e];
                    const float dn_pre = dn * (1.0f - n4[e] * n4[e]);
                    const float dr = dn_pre * hn4[e];
                    o_hn[it][e] = dn_pre * r4[e];
                    o_r[it][e] = dr * r4[e] * (1.0f - r4[e]);
                    o_z[it][e] = dz * z4[e] * (1.0f - z4[e]);
                    o_n[it][e] = dn_pre;
                    if (C::DEC) {
                        rs[C::DEC ? it : 0][0][e] += o_r[it][e];
                        rs[C::DEC ? it : 0][1][e] += o_z[it][e];
                        rs[C::DEC ? it : 0][2][e] += o_n[it][e];
                    }
                }
            }
            if (i == L) break;
            // the gradient MMAs of the previous step still read X / Hx: wait for them before overwriting the tiles
            if (i > 0) tc::mbar_wait(&bar_w, (i - 1) & 1);
#pragma unroll
            for (int it = 0; it < C::ITEMS; ++it) {
                const int b = ib[it], j0 = ij[it];
                if (b < 0) continue;
                const int boff = (b >> 3) * 128 + (b & 7) * 16;
                uint2 hi, lo;
                int k = j0;
                split4(o_r[it], hi, lo);
                *reinterpret_cast<uint2*>(X0 + (k >> 3) * C::X_LBO + boff + (k & 7) * 2) = hi;
                *reinterpret_cast<uint2*>(X1 + (k >> 3) * C::X_LBO + boff + (k & 7) * 2) = lo;
                k = HP + j0;
                split4(o_z[it], hi, lo);
                *reinterpret_cast<uint2*>(X0 + (k >> 3) * C::X_LBO + boff + (k & 7) * 2) = hi;
                *reinterpret_cast<uint2*>(X1 + (k >> 3) * C::X_LBO + boff + (k & 7) * 2) = lo;
                k = 2 * HP + j0;
                split4(o_hn[it], hi, lo);
                *reinterpret_cast<uint2*>(X0 + (k >> 3) * C::X_LBO + boff + (k & 7) * 2) = hi;
                *reinterpret_cast<uint2*>(X1 + (k >> 3) * C::X_LBO + boff + (k & 7) * 2) = lo;
                k = KPAD + j0;
                split4(o_n[it], hi, lo);
                *reinterpret_cast<uint2*>(X0 + (k >> 3) * C::X_LBO + boff + (k & 7) * 2) = hi;
                *reinterpret_cast<uint2*>(X1 + (k >> 3) * C::X_LBO + boff + (k & 7) * 2) = lo;
                // h_prev as the N side of the dW_hh contraction
                const float hp4[4] = {ph[it].x, ph[it].y, ph[it].z, ph[it].w};
                split4(hp4, hi, lo);
                *reinterpret_cast<uint2*>(Hh + (j0 >> 3) * C::H_LBO + boff + (j0 & 7) * 2) = hi;
                *reinterpret_cast<uint2*>(Hl + (j0 >> 3) * C::H_LBO + boff + (j0 & 7) * 2) = lo;
            }
            // one-hot token columns (bf16 1.0 = 0x3F80): thread = (row, chunk of 8 tokens)
            if (tid < NB * 4) {
                const int b = tid >> 2, ch = tid & 3;
                const int t = dir ? (L - 1 - s) : s;
                const int tk = toks[b * L + t] - ch * 8;
                uint32_t w[4] = {0u, 0u, 0u, 0u};
                if (tk >= 0 && tk < 8 && row0 + b < B) w[tk >> 1] = (tk & 1) ? 0x3F800000u : 0x00003F80u;
                *reinterpret_cast<uint4*>(Hh + (C::NH / 8 + ch) * C::H_LBO + (b >> 3) * 128 + (b & 7) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
            }
            tc::fence_proxy_async();
            tc::mbar_arrive(&bar_x);
            if (s > 0) prefetch(s - 1);                          // in flight under the MMAs
        }
        // ---- the CTA's partial weight gradients: TMEM -> global
        tc::mbar_wait(&bar_w, (L - 1) & 1);
        tc::tc_fence_after();
        {
            float* pw = (dir ? a.part_w[1] : a.part_w[0]) + (size_t)blockIdx.x * K3 * HP;
            float* pt = (dir ? a.part_t[1] : a.part_t[0]) + (size_t)blockIdx.x * a.V * 4 * HP;
            const int q = warp & 3;
            const int wq = warp >> 2, nwq = (C::NWE - q + 3) >> 2;      // this warp's rank / count among the warps of its quadrant
            // dW_hh: tile t, lane = gate row g - 128 t, columns = h index k
            int task = 0;
            for (int t = 0; t < C::NT_W; ++t) {
                for (int c0 = 0; c0 < C::NH; c0 += 16, ++task) {
                    if (task % nwq != wq) continue;
                    float v[16];
                    tmem_ld_cols<16>(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(C::COL_W + t * C::NH + c0), v);
                    const int g = t * 128 + q * 32 + lane;
                    if (g < K3) {
#pragma unroll
                        for (int c = 0; c < 16; c += 4)
                            if (c0 + c < HP) st4(pw + (size_t)g * HP + c0 + c, make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]));
                    }
                }
            }
            // dT: tile t, lane = X gate index g - 128 t, columns = token v.  X planes (dr, dz, dhn | pad | dn) -> (0, 1, 3 | - | 2)
            for (int t = 0; t < C::NT_T; ++t) {
                for (int c0 = 0; c0 < 32; c0 += 16, ++task) {
                    if (task % nwq != wq) continue;
                    float v[16];
                    tmem_ld_cols<16>(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(C::COL_T + t * 32 + c0), v);
                    const int g = t * 128 + q * 32 + lane;
                    int col = -1;
                    if (g < K3) { const int pl = g / HP; col = (pl == 2 ? 3 : pl) * HP + (g - pl * HP); }
                    else if (g >= KPAD && g < C::KX) col = 2 * HP + (g - KPAD);
                    if (col >= 0) {
#pragma unroll
                        for (int c = 0; c < 16; ++c)
                            if (c0 + c < a.V) pt[(size_t)(c0 + c) * 4 * HP + col] = v[c];
                    }
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == C::NWE) tc::tmem_dealloc<512>(tmem);
}

using EncF = FCfg<ENC_H, 64, 20, false>;
using DecF = FCfg<DEC_HP, 32, 13, true>;

template <class K>
int set_smem_f(K kfn, size_t bytes, size_t& set_for) {
    if (set_for < bytes) {
        if (cudaFuncSetAttribute((const void*)kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
            cudaGetLastError();
            return CPG_ECUDA;
        }
        set_for = bytes;
    }
    return CPG_OK;
}
}  // namespace

int bptt_fused_ctas_enc(int B) { return ceil_div(B, EncF::NB); }
int bptt_fused_ctas_dec(int B) { return ceil_div(B, DecF::NB); }

int launch_gru_bwd_enc_fused(cudaStream_t s, const GruSeq* two, const uint8_t* tok, int B, int L, int V, float* const part_w[2],
                             float* const part_t[2]) {
    FArgs a;
    memset(&a, 0, sizeof(a));
    for (int d = 0; d < 2; ++d) {
        a.whh[d] = two[d].whh; a.hs[d] = two[d].hs; a.gates[d] = two[d].gates;
        a.part_w[d] = part_w[d]; a.part_t[d] = part_t[d];
    }
    a.tok = tok;
    a.dh_fin = two[0].dh_fin;
    a.B = B; a.L = L; a.V = V;
    const size_t smem = EncF::smem_bytes(L);
    static size_t set_for = 0;
    if (set_smem_f(k_gru_bwd_fused<EncF>, smem, set_for)) return CPG_ECUDA;
    CPG_LAUNCH_NAMED("k_gru_bwd_enc_fused", k_gru_bwd_fused<EncF>, dim3(ceil_div(B, EncF::NB), 2), EncF::NTHREADS, smem, s, a);
    return CPG_OK;
}

int launch_gru_bwd_dec_fused(cudaStream_t s, const GruSeq& q, const uint8_t* tok, int B, int L, int V, float* part_w, float* part_t) {
    FArgs a;
    memset(&a, 0, sizeof(a));
    a.whh[0] = q.whh; a.hs[0] = q.hs; a.gates[0] = q.gates;
    a.part_w[0] = part_w; a.part_t[0] = part_t;
    a.tok = tok;
    a.h0 = q.h0; a.dh_out = q.dh_out; a.dh0 = q.dh0; a.drow = q.drow;
    a.B = B; a.L = L; a.V = V;
    const size_t smem = DecF::smem_bytes(L);
    static size_t set_for = 0;
    if (set_smem_f(k_gru_bwd_fused<DecF>, smem, set_for)) return CPG_ECUDA;
    CPG_LAUNCH_NAMED("k_gru_bwd_dec_fused", k_gru_bwd_fused<DecF>, dim3(ceil_div(B, DecF::NB), 1), DecF::NTHREADS, smem, s, a);
    return CPG_OK;
}

}  // namespace cpg
#else   // CPG_EMU
namespace cpg {
int bptt_fused_ctas_enc(int) { return 1; }
int bptt_fused_ctas_dec(int) { return 1; }
int launch_gru_bwd_enc_fused(cudaStream_t, const GruSeq*, const uint8_t*, int, int, int, float* const*, float* const*) { return CPG_ECUDA; }
int launch_gru_bwd_dec_fused(cudaStream_t, const GruSeq&, const uint8_t*, int, int, int, float*, float*) { return CPG_ECUDA; }
}  // namespace cpg
#endif
