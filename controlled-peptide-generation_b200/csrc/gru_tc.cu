// GRU recurrences (forward with activation stash, and BPTT) on the 5th-gen tensor cores, with
// fp32-grade accuracy.  Same math as the fp32 SIMT kernels of gru.cu (models/encoder.py:25-30,42 and
// models/decoder.py:40,77; gate order r,z,n; h' = (1-z) n + z h) and the same row-major h / dg layouts;
// the gate stash (r, z, n, hn) is private to this file's forward/backward pair and tile-major.
//
// Formulation.  The batch is the N side of the MMA and the weights are the M side ("transposed"):
//     forward   G^T[3H x nb] = W_hh[3H x H]   . h^T[H x nb]          (M tiles of 128 gate rows)
//     backward  dh^T[H x nb] = W_hh^T[H x 3H] . (dr,dz,dhn)^T[3H x nb]
// N = nb = 16 batch rows per MMA ("chain"), so B = 4096 gives 512 / 256 independent chains on 128 CTAs
// whatever the hidden size (with the batch on M a tile is 128 rows: 32..64 CTAs on 148 SMs).
// tcgen05.mma kind::f16 with bf16 operands and fp32 accumulation in TMEM; every fp32 operand x is split
// x = x1 + x2 (two bf16 terms, 16 mantissa bits) and the three products x1 w1 + x1 w2 + x2 w1 are
// accumulated: relative error 2^-16 per term, inside the 1e-4 parity bar (a single bf16/tf32 pass is not).
// W_hh (both terms) lives in TENSOR MEMORY as the A operand for all L steps (tcgen05.st once per CTA):
// from shared memory every N = 16..32 MMA re-read 4 KB of weights and ran at 43 clk instead of 16.
//
// One CTA = NCH chains of one direction (encoder 4, decoder 2), all L steps.  Per chain:
//   one MMA warp    : waits for the chain's operand tile of the previous step (mbarrier bar_x), an elected
//                     lane issues the step's MMAs and commits to bar_d; in the backward kernel it also streams the
//                     next step's gate planes in with cp.async.bulk (forward, decoder: streams the finished ones out)
//   NWG warps       : phase 1  tcgen05.ld (lane = gate row, nb batch columns) -> P[batch][gate] in shared memory
//                     phase 2  thread = (batch row, 4 consecutive hidden units): gate math, fp32 stash with
//                              row-contiguous float4 stores, next operand tile (bf16 split, K-major core matrices,
//                              16-byte pad on the K stride: bank-conflict-free 8-byte stores), arrive on bar_x
//   The chains of a CTA are independent: the MMA -> TMEM -> shared memory -> gate math latency chain of one runs
//   under the others' (named barrier 1 + chain for the phase 1 -> 2 hand-over, no CTA-wide barrier in the loop).
#include "ctx.h"
#ifndef CPG_EMU
#include "tc_gru.cuh"

#ifndef CPG_CHAIN_SKEW_NS
#define CPG_CHAIN_SKEW_NS 900
#endif
#ifndef CPG_DEC_NWG
#define CPG_DEC_NWG 7              // decoder BPTT (13 measured 4 us slower there)
#endif
#ifndef CPG_DEC_FWD_NWG
#define CPG_DEC_FWD_NWG 13         // decoder forward (measured 73 -> 70 us against 7)
#endif
#ifndef CPG_ENC_FWD_NWG
#define CPG_ENC_FWD_NWG 10
#endif
#ifndef CPG_FWD_STREAM_STORES
#define CPG_FWD_STREAM_STORES 2      // stash stores with the evict-first policy: 1 = the encoder's, 2 = and the decoder's gate planes (TMA)
#endif
#ifndef CPG_ENC_BWD_NWG
#define CPG_ENC_BWD_NWG 10
#endif

#ifdef CPG_GRU_TIMELINE
// developer-only: clock64() stamps of one CTA / one step (tools/gru_timeline.py); never in the product build
__device__ long long g_gru_tl[256];
#define CPG_TL(slot) do { if (blockIdx.x == 0 && blockIdx.y == 0 && s == 12) g_gru_tl[C::KID * 64 + (slot)] = clock64(); } while (0)
#define CPG_TL0(slot) do { if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) g_gru_tl[C::KID * 64 + (slot)] = clock64(); } while (0)
extern "C" int cpg_debug_gru_timeline(long long* out) { return (int)cudaMemcpyFromSymbol(out, g_gru_tl, sizeof(g_gru_tl)); }
#else
#define CPG_TL(slot) do { } while (0)
#define CPG_TL0(slot) do { } while (0)
#endif

namespace cpg {
int check_launch(const char* where);

namespace {
// ------------------------------------------------------------------------------------- forward
template <int HP_, int KP_, int NB_, int NCH_, int NWG_, bool DEC_>
struct FwdCfg {
    static constexpr int HP = HP_, KP = KP_, NB = NB_, NCH = NCH_, NWG = NWG_;
    static constexpr bool DEC = DEC_;
    static constexpr int KID = DEC ? 1 : 0;                 // timeline-probe slot
    static constexpr int G3 = 3 * HP;                       // gate rows (M)
    static constexpr int MT = (G3 + 127) / 128;             // M tiles
    static constexpr int NQ = HP / 4;                       // unit quads per row
    static constexpr int NITEMS = NB * NQ;                  // (row, quad) items per chain and step
    static constexpr int NT_G = NWG * 32;                   // epilogue threads of one chain
    static constexpr int ITEMS = (NITEMS + NT_G - 1) / NT_G;
    static constexpr int NW_EPI = NCH * NWG;
    static constexpr int NTHREADS = (NW_EPI + NCH) * 32;    // + one MMA warp per chain
    static constexpr int KC = KP / 8, KSTEPS = KP / 16;
    static constexpr int X_LBO = (NB / 8) * 128 + 16;       // K-adjacent core matrices (padded: conflict-free stores)
    static constexpr int X_SPLIT = KC * X_LBO;
    static constexpr int P_FLOATS = NB * G3;
    // tensor memory: accumulators [NCH][MT] x NB columns, then W_hh as the A operand: [MT][2 terms][KP/2] columns
    static constexpr int WCOL0 = NCH * MT * NB;
    static constexpr int WCOLS = KP / 2;
    static constexpr uint32_t TMEM_COLS = 512;
    static_assert(WCOL0 + MT * 2 * WCOLS <= 512, "TMEM columns");
    static_assert(NWG >= 4, "every lane quadrant needs a warp in each group");
    static_assert(HP % 8 == 0 && KP % 16 == 0 && G3 % 8 == 0 && (NB == 16 || NB == 32), "geometry");
    static size_t smem_bytes(int V, int L) {
        return (size_t)NCH * 2 * X_SPLIT + (size_t)NCH * P_FLOATS * 4 + (DEC ? 0 : (size_t)V * G3 * 4) +
               (((size_t)NCH * NB * L + 15) & ~(size_t)15) + (STAGE_G ? (size_t)NCH * 4 * G_PLANE : 0) + 128;
    }
    static constexpr int G_PLANE = NB * HP * 4;             // bytes of one gate plane of one chain and step (staging)
    // gate planes through a shared-memory staging tile + TMA bulk stores (measured: decoder 80 -> 72 us; the encoder's
    // 40 KB per chain-step read-out would sit on the MMA completion signal, 110 -> 118 us, so it stores directly)
    static constexpr bool STAGE_G = DEC;
};

struct FwdArgs {
    const uint8_t* tok;        // [B][L]
    const float* table[2];     // [V][3*HP] per direction
    const float* whh[2];       // [3*HP][HP] natural (zero padded for the decoder)
    const float* bhn[2];       // [HP]
    const float* rowbias;      // decoder: [B][3*HP]
    const float* h0;           // decoder: [B][HP]
    float* hs[2];              // [B][L][HP] by step (nullable)
    float* gates[2];           // [ceil(B/32)][L][4][32][HP] tiled gate stash (nullable)
    float* hfin;               // encoder: [B][2*HP]
    int B, L, V;
    int nprod;                 // 3: split-bf16 products x1 w1 + x1 w2 + x2 w1 (fp32-grade); 1: the leading bf16 product only
};

template <class C>
__global__ void __launch_bounds__(C::NTHREADS, 1)
k_gru_fwd_tc(FwdArgs a) {
    constexpr int HP = C::HP, G3 = C::G3, MT = C::MT, NQ = C::NQ, NCH = C::NCH, KC = C::KC, NB = C::NB;
    extern __shared__ __align__(1024) unsigned char smem[];   // used directly: keeps every access an LDS/STS
    unsigned char* Xb = smem;                                            // [NCH][2 terms][KC][X_LBO]
    float* Pb = reinterpret_cast<float*>(Xb + NCH * 2 * C::X_SPLIT);     // [NCH][NB][G3]
    float* tab = Pb + NCH * C::P_FLOATS;                                 // encoder: [V][G3]
    uint8_t* toks = reinterpret_cast<uint8_t*>(tab + (C::DEC ? 0 : a.V * G3));   // [NCH*NB][L]
    float* Gs = reinterpret_cast<float*>(toks + ((NCH * NB * a.L + 15) & ~15));     // [NCH][4 planes][NB][HP]: gate stash staging
    __shared__ __align__(8) uint64_t bar_x[NCH], bar_d[NCH];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dir = blockIdx.y;
    const int row0 = blockIdx.x * (NCH * NB);
    const int B = a.B, L = a.L;
    CPG_TL0(50);

    // ---- one-time setup
    for (int idx = tid; idx < NCH * NB * KC; idx += C::NTHREADS) {
        const int bb = idx / KC, kc = idx % KC, ch = bb / NB, b = bb % NB;
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = 0.f;
        if (C::DEC && kc * 8 + 8 <= HP) {
            const int row = min(row0 + bb, B - 1);
            const float4 f0 = ldg4(a.h0 + (size_t)row * HP + kc * 8), f1 = ldg4(a.h0 + (size_t)row * HP + kc * 8 + 4);
            x[0] = f0.x; x[1] = f0.y; x[2] = f0.z; x[3] = f0.w; x[4] = f1.x; x[5] = f1.y; x[6] = f1.z; x[7] = f1.w;
        }
        uint4 hi, lo;
        split8(x, hi, lo);
        const int off = kc * C::X_LBO + (b >> 3) * X_SBO + (b & 7) * 16;
        *reinterpret_cast<uint4*>(Xb + (ch * 2 + 0) * C::X_SPLIT + off) = hi;
        *reinterpret_cast<uint4*>(Xb + (ch * 2 + 1) * C::X_SPLIT + off) = lo;
    }
    if (!C::DEC) {
        const float* tsrc = (dir ? a.table[1] : a.table[0]);
        for (int i = tid; i < a.V * G3; i += C::NTHREADS) tab[i] = __ldg(tsrc + i);
    }
    for (int i = tid; i < NCH * NB * L; i += C::NTHREADS) {
        const int row = min(row0 + i / L, B - 1);
        toks[i] = a.tok[(size_t)row * L + i % L];
    }
    if (warp == C::NW_EPI) {
        if (lane == 0) {
            for (int i = 0; i < NCH; ++i) {
                tc::mbar_init(&bar_x[i], C::NT_G);
                tc::mbar_init(&bar_d[i], 1);
            }
            tc::fence_barrier_init();
        }
        __syncwarp();
        tc::tmem_alloc<C::TMEM_COLS>(&tmem_slot);
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = tmem_slot;

    // W_hh -> tensor memory (A operand of every MMA of this CTA).  Task = (M tile, 16-wide K slice) of the lane
    // quadrant warp % 4; every warp of the CTA takes its share.  lane = gate row; leading and remainder bf16
    // terms side by side.
    {
        const float* whh = (dir ? a.whh[1] : a.whh[0]);
        constexpr int NWARPS = C::NTHREADS / 32;
        const int q = warp & 3;
        const int cnt = (NWARPS - q + 3) >> 2;
        for (int task = warp >> 2; task < MT * C::KSTEPS; task += cnt) {
            const int t = task / C::KSTEPS, ks = task % C::KSTEPS;
            const int m = t * 128 + q * 32 + lane;
            float x[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) x[e] = 0.f;
            if (m < G3) {
#pragma unroll
                for (int e4 = 0; e4 < 4; ++e4) {
                    const int k = ks * 16 + e4 * 4;
                    if (k + 4 <= HP) {
                        const float4 f = ldg4(whh + (size_t)m * HP + k);
                        x[e4 * 4 + 0] = f.x; x[e4 * 4 + 1] = f.y; x[e4 * 4 + 2] = f.z; x[e4 * 4 + 3] = f.w;
                    }
                }
            }
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) split2(x[2 * e], x[2 * e + 1], hi[e], lo[e]);
            const uint32_t lane_addr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(C::WCOL0 + t * 2 * C::WCOLS + ks * 8);
            tmem_st_32x8(lane_addr, hi);
            tmem_st_32x8(lane_addr + (uint32_t)C::WCOLS, lo);
        }
        tmem_st_wait();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    CPG_TL0(51);

    if (warp >= C::NW_EPI) {
        // ---------------- MMA issuer of chain ch (whole warp converged; one elected lane issues)
        const int ch = warp - C::NW_EPI;
        constexpr uint32_t idesc = make_idesc_bf16(128, NB);
        const uint32_t x0 = tc::smem_u32(Xb) + (uint32_t)(ch * 2 * C::X_SPLIT);
        const uint32_t d0 = tmem_d + (uint32_t)(ch * MT * NB);
        // start the chains half a step apart: the SFU-bound gate math of one then runs under the MMA / TMEM
        // read-out of the other instead of both phases coinciding
        if (ch > 0) __nanosleep(CPG_CHAIN_SKEW_NS * 2 * ch / NCH);
        // gate planes of a finished step leave through the TMA engine: the epilogue only writes them to the staging
        // tile in shared memory (its arrival on bar_x says they are there), this lane streams the four 10 KB planes out
        float* gates_dst = (dir ? a.gates[1] : a.gates[0]);
        const float* Gch = Gs + ch * 4 * NB * HP;
        auto store_gates = [&](int s_done) {
            if (!C::STAGE_G || gates_dst == nullptr || row0 + ch * NB >= B) return;   // no stash wanted / chain entirely past the batch
#pragma unroll
            for (int pl = 0; pl < 4; ++pl)
                if (CPG_FWD_STREAM_STORES > 1)
                    bulk_store_stream(gates_dst + gate_stash_offset<HP>(row0 + ch * NB, s_done, L) + (size_t)pl * 32 * HP,
                                      Gch + pl * NB * HP, C::G_PLANE);
                else
                    bulk_store(gates_dst + gate_stash_offset<HP>(row0 + ch * NB, s_done, L) + (size_t)pl * 32 * HP,
                               Gch + pl * NB * HP, C::G_PLANE);
            bulk_commit();
        };
        for (int s = 0; s < L; ++s) {
            if (s > 0) {
                tc::mbar_wait(&bar_x[ch], (s - 1) & 1);
                tc::tc_fence_after();
            }
            if (elect_one()) {
                CPG_TL(0 + ch);
                if (s > 0) store_gates(s - 1);
#pragma unroll
                for (int t = 0; t < MT; ++t) {
                    uint32_t acc = 0;
#pragma unroll
                    for (int p = 0; p < 3; ++p) {
                        if (p >= a.nprod) break;
#pragma unroll
                        for (int ks = 0; ks < C::KSTEPS; ++ks) {
                            const uint32_t ta = tmem_d + (uint32_t)(C::WCOL0 + (t * 2 + WS[p]) * C::WCOLS + ks * 8);
                            const uint64_t db = tc::make_smem_desc(x0 + XS[p] * C::X_SPLIT + ks * 2 * C::X_LBO, C::X_LBO, X_SBO, 0);
                            umma_bf16_ts(d0 + (uint32_t)(t * NB), ta, db, idesc, acc);
                            acc = 1;
                        }
                    }
                }
                // the staging tile is rewritten by the gate math that follows these MMAs: its read-out must be over
                // before their completion is signalled (it runs under the MMAs, so this does not add latency)
                if (C::STAGE_G && s > 0) bulk_wait_read();
                tc::umma_commit(&bar_d[ch]);
            }
            __syncwarp();
        }
        tc::mbar_wait(&bar_x[ch], (L - 1) & 1);
        if (elect_one()) {
            store_gates(L - 1);
            bulk_wait_all();
        }
        __syncwarp();
    } else {
        // ---------------- epilogue of chain ch
        const int ch = warp / C::NWG, wl = warp % C::NWG, tl = tid - ch * C::NT_G;
        unsigned char* X0 = Xb + (ch * 2 + 0) * C::X_SPLIT;
        unsigned char* X1 = Xb + (ch * 2 + 1) * C::X_SPLIT;
        float* P = Pb + ch * C::P_FLOATS;
        float* Gch = Gs + ch * 4 * NB * HP;
        const uint8_t* tks = toks + ch * NB * L;
        const uint32_t d0 = tmem_d + (uint32_t)(ch * MT * NB);
        const int rowc0 = row0 + ch * NB;
        const float* bhn_g = (dir ? a.bhn[1] : a.bhn[0]);
        const float* tabg = C::DEC ? a.table[0] : tab;
        float* hs_g = (dir ? a.hs[1] : a.hs[0]);
        float* gates_g = (dir ? a.gates[1] : a.gates[0]);

        int ib[C::ITEMS], ij[C::ITEMS];                      // item -> (row in tile, first unit); row < 0: no item
        float bhn[C::ITEMS][4];
        float hprev[C::ITEMS][4];
        float rb[C::DEC ? C::ITEMS : 1][3][4];
#pragma unroll
        for (int it = 0; it < C::ITEMS; ++it) {
            const int idx = tl + it * C::NT_G;
            const bool valid = idx < C::NITEMS;
            ib[it] = valid ? idx / NQ : -1;
            ij[it] = valid ? (idx % NQ) * 4 : 0;
            const int row = min(rowc0 + max(ib[it], 0), B - 1);
            const float4 b4 = ldg4(bhn_g + ij[it]);
            bhn[it][0] = b4.x; bhn[it][1] = b4.y; bhn[it][2] = b4.z; bhn[it][3] = b4.w;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (C::DEC) {
                v = ldg4(a.h0 + (size_t)row * HP + ij[it]);
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    const float4 r4 = ldg4(a.rowbias + (size_t)row * G3 + g * HP + ij[it]);
                    rb[C::DEC ? it : 0][g][0] = r4.x; rb[C::DEC ? it : 0][g][1] = r4.y;
                    rb[C::DEC ? it : 0][g][2] = r4.z; rb[C::DEC ? it : 0][g][3] = r4.w;
                }
            }
            hprev[it][0] = v.x; hprev[it][1] = v.y; hprev[it][2] = v.z; hprev[it][3] = v.w;
        }

        for (int s = 0; s < L; ++s) {
            const int t = dir ? (L - 1 - s) : s;
            // input-side pre-activations do not depend on the MMA: fetch them before waiting on it
            float4 tin[C::ITEMS][3];
#pragma unroll
            for (int it = 0; it < C::ITEMS; ++it) {
                const int tk = tks[max(ib[it], 0) * L + t];
                const float* trow = tabg + tk * G3 + ij[it];
                if (C::DEC) { tin[it][0] = ldg4(trow); tin[it][1] = ldg4(trow + HP); tin[it][2] = ldg4(trow + 2 * HP); }
                else { tin[it][0] = ld4(trow); tin[it][1] = ld4(trow + HP); tin[it][2] = ld4(trow + 2 * HP); }
            }
            if (lane == 0 && wl == 0) CPG_TL(8 + 16 * ch);
            tc::mbar_wait(&bar_d[ch], s & 1);
            tc::tc_fence_after();
            if (lane == 0 && wl == 0) CPG_TL(9 + 16 * ch);
            // phase 1: accumulator (lane = gate row, NB batch columns) -> P[batch][gate]
#pragma unroll
            for (int tq = 0; tq < MT; ++tq) {
                if (quadrant_task_is_mine(wl, C::NWG, tq)) {
                    const int q = warp & 3;
                    const int m = tq * 128 + q * 32 + lane;
                    if (tq * 128 + q * 32 < G3) {            // warp-uniform: skip quadrants that hold no gate row
                        float v[NB];
                        tmem_ld_cols<NB>(d0 + ((uint32_t)(q * 32) << 16) + (uint32_t)(tq * NB), v);
                        if (m < G3) {
#pragma unroll
                            for (int c = 0; c < NB; ++c) P[c * G3 + m] = v[c];
                        }
                    }
                }
            }
            tc::tc_fence_before();
            if (lane == 0 && wl == 0) CPG_TL(10 + 16 * ch);
            group_bar_sync(1 + ch, C::NT_G);
            if (lane == 0 && wl == 0) CPG_TL(11 + 16 * ch);
            // phase 2: gates for (row, 4 units)
#pragma unroll
            for (int it = 0; it < C::ITEMS; ++it) {
                const int b = ib[it], j0 = ij[it];
                if (b < 0) continue;
                const int row = rowc0 + b;
                const float4 pr = ld4(P + b * G3 + j0), pz = ld4(P + b * G3 + HP + j0), pn = ld4(P + b * G3 + 2 * HP + j0);
                const float4 tr = tin[it][0], tz = tin[it][1], tn = tin[it][2];
                float gr[4] = {tr.x + pr.x, tr.y + pr.y, tr.z + pr.z, tr.w + pr.w};
                float gz[4] = {tz.x + pz.x, tz.y + pz.y, tz.z + pz.z, tz.w + pz.w};
                float gn[4] = {tn.x, tn.y, tn.z, tn.w};
                const float pnv[4] = {pn.x, pn.y, pn.z, pn.w};
                float rr[4], zz[4], nn[4], hh[4], hn[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (C::DEC) {
                        gr[e] += rb[C::DEC ? it : 0][0][e];
                        gz[e] += rb[C::DEC ? it : 0][1][e];
                        gn[e] += rb[C::DEC ? it : 0][2][e];
                    }
                    sigmoid_pair(gr[e], gz[e], rr[e], zz[e]);
                    hh[e] = pnv[e] + bhn[it][e];
                    nn[e] = tanh_fast(gn[e] + rr[e] * hh[e]);
                    hn[e] = (1.0f - zz[e]) * nn[e] + zz[e] * hprev[it][e];
                    hprev[it][e] = hn[e];
                }
                uint2 hi, lo;
                split4(hn, hi, lo);
                const int off = (j0 >> 3) * C::X_LBO + (b >> 3) * X_SBO + (b & 7) * 16 + (j0 & 7) * 2;
                *reinterpret_cast<uint2*>(X0 + off) = hi;
                *reinterpret_cast<uint2*>(X1 + off) = lo;
                if (C::STAGE_G && gates_g != nullptr) {          // staging tile [plane][row in chain][HP] (linear: conflict-free)
                    float* g = Gch + b * HP + j0;
                    st4(g, make_float4(rr[0], rr[1], rr[2], rr[3]));
                    st4(g + NB * HP, make_float4(zz[0], zz[1], zz[2], zz[3]));
                    st4(g + 2 * NB * HP, make_float4(nn[0], nn[1], nn[2], nn[3]));
                    st4(g + 3 * NB * HP, make_float4(hh[0], hh[1], hh[2], hh[3]));
                }
                if (row < B) {
                    const size_t bs = (size_t)row * L + s;
                    if (hs_g != nullptr) {
                        // the encoder's h stash is read again by its BPTT only (~0.4 ms later); the decoder's feeds the output layer next
                        if (CPG_FWD_STREAM_STORES && !C::DEC) st_stream4(hs_g + bs * HP + j0, make_float4(hn[0], hn[1], hn[2], hn[3]));
                        else st_once4(hs_g + bs * HP + j0, make_float4(hn[0], hn[1], hn[2], hn[3]));
                    }
                    if (!C::STAGE_G && gates_g != nullptr) {
                        float* g = gates_g + gate_stash_offset<HP>(row, s, L) + j0;
                        if (CPG_FWD_STREAM_STORES) {
                            st_stream4(g, make_float4(rr[0], rr[1], rr[2], rr[3]));
                            st_stream4(g + 32 * HP, make_float4(zz[0], zz[1], zz[2], zz[3]));
                            st_stream4(g + 2 * 32 * HP, make_float4(nn[0], nn[1], nn[2], nn[3]));
                            st_stream4(g + 3 * 32 * HP, make_float4(hh[0], hh[1], hh[2], hh[3]));
                        } else {
                            st4(g, make_float4(rr[0], rr[1], rr[2], rr[3]));
                            st4(g + 32 * HP, make_float4(zz[0], zz[1], zz[2], zz[3]));
                            st4(g + 2 * 32 * HP, make_float4(nn[0], nn[1], nn[2], nn[3]));
                            st4(g + 3 * 32 * HP, make_float4(hh[0], hh[1], hh[2], hh[3]));
                        }
                    }
                    if (!C::DEC && s == L - 1)
                        st4(a.hfin + (size_t)row * (2 * HP) + dir * HP + j0, make_float4(hn[0], hn[1], hn[2], hn[3]));
                }
            }
            if (lane == 0 && wl == 0) CPG_TL(12 + 16 * ch);
            tc::fence_proxy_async();                         // operand-tile stores -> visible to the tensor core
            tc::mbar_arrive(&bar_x[ch]);
            if (lane == 0 && wl == 0) CPG_TL(13 + 16 * ch);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    CPG_TL0(52);
    if (warp == C::NW_EPI) tc::tmem_dealloc<C::TMEM_COLS>(tmem_d);
}

// ------------------------------------------------------------------------------------ backward
// Per step (reverse step order), as in gru.cu:
//   dht = dh + dh_out[s];  dn = dht (1-z); dz = dht (h_prev - n); carry = dht z
//   dn_pre = dn (1-n^2); dr = dn_pre hn; dhn = dn_pre r; dr_pre = dr r (1-r); dz_pre = dz z (1-z)
//   dh = carry + W_hh^T [dr_pre, dz_pre, dhn]
template <int HP_, int NB_, int NCH_, int NWG_, bool DEC_>
struct BwdCfg {
    static constexpr int HP = HP_, NB = NB_, NCH = NCH_, NWG = NWG_;
    static constexpr bool DEC = DEC_;
    static constexpr int KID = DEC ? 2 : 3;
    static constexpr int K3 = 3 * HP;
    static constexpr int KPAD = (K3 + 15) / 16 * 16;        // 240 | 320
    static constexpr int KC = KPAD / 8, KSTEPS = KPAD / 16;
    static constexpr int NQ = HP / 4;
    static constexpr int NITEMS = NB * NQ;
    static constexpr int NT_G = NWG * 32;
    static constexpr int ITEMS = (NITEMS + NT_G - 1) / NT_G;
    static constexpr int NW_EPI = NCH * NWG;
    static constexpr int NTHREADS = (NW_EPI + NCH) * 32;
    static constexpr int X_LBO = (NB / 8) * 128 + 16;
    static constexpr int X_SPLIT = KC * X_LBO;
    static constexpr int P_FLOATS = NB * HP;
    // tensor memory: accumulators [NCH] x NB columns, then W_hh^T as the A operand: [2 terms][KPAD/2] columns
    static constexpr int WCOL0 = NCH * NB < 32 ? 32 : NCH * NB;
    static constexpr int WCOLS = KPAD / 2;
    static constexpr uint32_t TMEM_COLS = 512;
    static_assert(WCOL0 + 2 * WCOLS <= 512, "TMEM columns");
    static_assert(HP <= 128 && NWG >= 4 && (NB == 16 || NB == 32), "geometry");
    static constexpr int G_PLANE = NB * HP * 4;              // bytes of one gate plane of one chain and step
    static size_t smem_bytes() { return (size_t)NCH * 2 * X_SPLIT + (size_t)NCH * P_FLOATS * 4 + (size_t)NCH * 4 * G_PLANE + 128; }
};

struct BwdArgs {
    const float* whh[2];       // [3*HP][HP] natural
    const float* hs[2];        // [B][L][HP]
    const float* gates[2];     // [ceil(B/32)][L][4][32][HP] tiled gate stash
    const float* h0;           // decoder: [B][HP] (null = zeros)
    const float* dh_out;       // decoder: [B][L][HP]
    const float* dh_fin;       // encoder: [B][2*HP]
    float* dg[2];              // [B][L][4][HP]
    float* dh0;                // decoder: [B][HP]
    float* drow;               // decoder: [B][3*HP]
    int B, L;
    int round_dg;              // 1: store dg rounded to tf32 (its consumer is k_wgrad_tc)
};

template <class C>
__global__ void __launch_bounds__(C::NTHREADS, 1)
k_gru_bwd_tc(BwdArgs a) {
    constexpr int HP = C::HP, NQ = C::NQ, NCH = C::NCH, K3 = C::K3, NB = C::NB;
    extern __shared__ __align__(1024) unsigned char smem[];   // used directly: keeps every access an LDS/STS
    unsigned char* Xb = smem;                                            // [NCH][2 terms][KC][X_LBO]
    float* Pb = reinterpret_cast<float*>(Xb + NCH * 2 * C::X_SPLIT);     // [NCH][NB][HP]
    float* Gb = Pb + NCH * C::P_FLOATS;                                  // [NCH][4 planes][NB][HP]: gate stash of the current step
    __shared__ __align__(8) uint64_t bar_x[NCH], bar_d[NCH], bar_g[NCH];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dir = blockIdx.y;
    const int row0 = blockIdx.x * (NCH * NB);
    const int B = a.B, L = a.L;
    CPG_TL0(50);

    // ---- one-time setup: operand tiles zeroed (K padding stays 0)
    for (int i = tid; i < NCH * 2 * C::X_SPLIT / 16; i += C::NTHREADS) reinterpret_cast<uint4*>(Xb)[i] = make_uint4(0, 0, 0, 0);
    if (warp == C::NW_EPI) {
        if (lane == 0) {
            for (int i = 0; i < NCH; ++i) {
                tc::mbar_init(&bar_x[i], C::NT_G);
                tc::mbar_init(&bar_d[i], 1);
                tc::mbar_init(&bar_g[i], 1);
            }
            tc::fence_barrier_init();
        }
        __syncwarp();
        tc::tmem_alloc<C::TMEM_COLS>(&tmem_slot);
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = tmem_slot;

    // W_hh^T -> tensor memory: lane = hidden unit j, K index = gate row k, A[j][k] = W_hh[k][j].
    // Task = 16-wide K slice of the lane quadrant warp % 4, dealt over all warps of the CTA.
    {
        const float* whh = (dir ? a.whh[1] : a.whh[0]);
        constexpr int NWARPS = C::NTHREADS / 32;
        const int q = warp & 3;
        const int cnt = (NWARPS - q + 3) >> 2;
        const int j = q * 32 + lane;
        if (q * 32 < HP) {
            for (int ks = warp >> 2; ks < C::KSTEPS; ks += cnt) {
                float x[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const int k = ks * 16 + e;
                    x[e] = (j < HP && k < K3) ? __ldg(whh + (size_t)k * HP + j) : 0.f;
                }
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) split2(x[2 * e], x[2 * e + 1], hi[e], lo[e]);
                const uint32_t lane_addr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(C::WCOL0 + ks * 8);
                tmem_st_32x8(lane_addr, hi);
                tmem_st_32x8(lane_addr + (uint32_t)C::WCOLS, lo);
            }
            tmem_st_wait();
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    CPG_TL0(51);

    if (warp >= C::NW_EPI) {
        // ---------------- MMA issuer of chain ch: iteration i handles step s = L-1-i
        const int ch = warp - C::NW_EPI;
        constexpr uint32_t idesc = make_idesc_bf16(128, NB);
        const uint32_t x0 = tc::smem_u32(Xb) + (uint32_t)(ch * 2 * C::X_SPLIT);
        const uint32_t d0 = tmem_d + (uint32_t)(ch * NB);
        // gate stash of (this chain, step s) -> shared memory: one bulk copy per plane through the TMA engine
        const float* gates_src = (dir ? a.gates[1] : a.gates[0]);
        float* G = Gb + ch * 4 * NB * HP;
        // a chain entirely past the batch (ragged last CTA) has no stash tile of its own: it reads tile 0 instead
        // (its results are never stored), so that no copy leaves the stash allocation of ceil(B/32) tiles
        const int first_row = row0 + ch * NB;
        const int src_row = first_row < ((B + 31) & ~31) ? first_row : (first_row & 31);
        auto load_gates = [&](int s) {
            tc::mbar_expect_tx(&bar_g[ch], 4 * C::G_PLANE);
#pragma unroll
            for (int pl = 0; pl < 4; ++pl)
                bulk_load(G + pl * NB * HP, gates_src + gate_stash_offset<HP>(src_row, s, L) + (size_t)pl * 32 * HP,
                          C::G_PLANE, &bar_g[ch]);
        };
        if (elect_one()) load_gates(L - 1);
        __syncwarp();
        for (int i = 0; i < L; ++i) {
            tc::mbar_wait(&bar_x[ch], i & 1);          // phase 2 of iteration i is over: X tile ready, G buffer free
            tc::tc_fence_after();
            if (elect_one()) {
                { const int s = L - 1 - i; CPG_TL(0 + ch); }
                if (i + 1 < L) load_gates(L - 2 - i);
                uint32_t acc = 0;
#pragma unroll
                for (int p = 0; p < 3; ++p) {
#pragma unroll
                    for (int ks = 0; ks < C::KSTEPS; ++ks) {
                        const uint32_t ta = tmem_d + (uint32_t)(C::WCOL0 + WS[p] * C::WCOLS + ks * 8);
                        const uint64_t db = tc::make_smem_desc(x0 + XS[p] * C::X_SPLIT + ks * 2 * C::X_LBO, C::X_LBO, X_SBO, 0);
                        umma_bf16_ts(d0, ta, db, idesc, acc);
                        acc = 1;
                    }
                }
                tc::umma_commit(&bar_d[ch]);
            }
            __syncwarp();
        }
    } else {
        // ---------------- epilogue of chain ch
        const int ch = warp / C::NWG, wl = warp % C::NWG, tl = tid - ch * C::NT_G;
        unsigned char* X0 = Xb + (ch * 2 + 0) * C::X_SPLIT;
        unsigned char* X1 = Xb + (ch * 2 + 1) * C::X_SPLIT;
        float* P = Pb + ch * C::P_FLOATS;
        const uint32_t d0 = tmem_d + (uint32_t)(ch * NB);
        const int rowc0 = row0 + ch * NB;
        const float* hs_g = (dir ? a.hs[1] : a.hs[0]);
        float* dg_g = (dir ? a.dg[1] : a.dg[0]);

        int ib[C::ITEMS], ij[C::ITEMS];
        float carry[C::ITEMS][4];
        float rs[C::DEC ? C::ITEMS : 1][3][4];
#pragma unroll
        for (int it = 0; it < C::ITEMS; ++it) {
            const int idx = tl + it * C::NT_G;
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
        // prefetch registers for the step about to be processed: gates (r,z,n,hn), h_prev, dh_out
        float4 ph[C::ITEMS], pd[C::ITEMS];
        const float* G = Gb + ch * 4 * NB * HP;
        auto prefetch = [&](int s) {
#pragma unroll
            for (int it = 0; it < C::ITEMS; ++it) {
                if (ib[it] < 0) continue;
                const int j0 = ij[it];
                const int row = min(rowc0 + ib[it], B - 1);
                const size_t bs = (size_t)row * L + s;
                const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
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
                if (lane == 0 && wl == 0) CPG_TL(8 + 16 * ch);
                tc::mbar_wait(&bar_d[ch], (i - 1) & 1);
                tc::tc_fence_after();
                if (lane == 0 && wl == 0) CPG_TL(9 + 16 * ch);
                const int q = warp & 3;
                if ((wl >> 2) == 0 && q * 32 < HP) {             // first warp of each quadrant; lane = hidden unit j
                    float v[NB];
                    tmem_ld_cols<NB>(d0 + ((uint32_t)(q * 32) << 16), v);
                    const int j = q * 32 + lane;
                    if (j < HP) {
#pragma unroll
                        for (int c = 0; c < NB; ++c) P[c * HP + j] = v[c];
                    }
                }
                tc::tc_fence_before();
                if (lane == 0 && wl == 0) CPG_TL(10 + 16 * ch);
                group_bar_sync(1 + ch, C::NT_G);
                if (lane == 0 && wl == 0) CPG_TL(11 + 16 * ch);
            }
            if (i < L) tc::mbar_wait(&bar_g[ch], i & 1);         // this step's gate planes have landed in G
#pragma unroll
            for (int it = 0; it < C::ITEMS; ++it) {
                const int b = ib[it], j0 = ij[it];
                if (b < 0) continue;
                const int row = rowc0 + b;
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
                const float4 g_r = ld4(G + b * HP + j0), g_z = ld4(G + (NB + b) * HP + j0);
                const float4 g_n = ld4(G + (2 * NB + b) * HP + j0), g_hn = ld4(G + (3 * NB + b) * HP + j0);
                const float r4[4] = {g_r.x, g_r.y, g_r.z, g_r.w};
                const float z4[4] = {g_z.x, g_z.y, g_z.z, g_z.w};
                const float n4[4] = {g_n.x, g_n.y, g_n.z, g_n.w};
                const float hn4[4] = {g_hn.x, g_hn.y, g_hn.z, g_hn.w};
                const float hp4[4] = {ph[it].x, ph[it].y, ph[it].z, ph[it].w};
                const float do4[4] = {pd[it].x, pd[it].y, pd[it].z, pd[it].w};
                float o_r[4], o_z[4], o_n[4], o_hn[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float dht = dh[e] + do4[e];
                    const float dn = dht * (1.0f - z4[e]);
                    const float dz = dht * (hp4[e] - n4[e]);
                    carry[it][e] = dht * z4[e];
                    const float dn_pre = dn * (1.0f - n4[e] * n4[e]);
                    const float dr = dn_pre * hn4[e];
                    o_hn[e] = dn_pre * r4[e];
                    o_r[e] = dr * r4[e] * (1.0f - r4[e]);
                    o_z[e] = dz * z4[e] * (1.0f - z4[e]);
                    o_n[e] = dn_pre;
                    if (C::DEC) {
                        rs[C::DEC ? it : 0][0][e] += o_r[e];
                        rs[C::DEC ? it : 0][1][e] += o_z[e];
                        rs[C::DEC ? it : 0][2][e] += o_n[e];
                    }
                }
                // operand tile: K index = g*HP + j for (dr_pre, dz_pre, dhn)
                const int boff = (b >> 3) * X_SBO + (b & 7) * 16;
                uint2 hi, lo;
                split4(o_r, hi, lo);
                int k = j0;
                *reinterpret_cast<uint2*>(X0 + (k >> 3) * C::X_LBO + boff + (k & 7) * 2) = hi;
                *reinterpret_cast<uint2*>(X1 + (k >> 3) * C::X_LBO + boff + (k & 7) * 2) = lo;
                split4(o_z, hi, lo);
                k = HP + j0;
                *reinterpret_cast<uint2*>(X0 + (k >> 3) * C::X_LBO + boff + (k & 7) * 2) = hi;
                *reinterpret_cast<uint2*>(X1 + (k >> 3) * C::X_LBO + boff + (k & 7) * 2) = lo;
                split4(o_hn, hi, lo);
                k = 2 * HP + j0;
                *reinterpret_cast<uint2*>(X0 + (k >> 3) * C::X_LBO + boff + (k & 7) * 2) = hi;
                *reinterpret_cast<uint2*>(X1 + (k >> 3) * C::X_LBO + boff + (k & 7) * 2) = lo;
                if (row < B) {
                    float* gp = dg_g + ((size_t)row * L + s) * 4 * HP + j0;
                    // dg's only reader on this path is the tf32 weight-gradient kernel: when it asks for it, the
                    // planes leave already rounded to tf32 (add half an ulp; the tensor core drops the low bits),
                    // which removes four of its five operand-conversion passes
                    const uint32_t rb = a.round_dg ? 0x1000u : 0u;
                    auto r4 = [rb](const float (&x)[4]) {
                        return make_float4(__uint_as_float(__float_as_uint(x[0]) + rb), __uint_as_float(__float_as_uint(x[1]) + rb),
                                           __uint_as_float(__float_as_uint(x[2]) + rb), __uint_as_float(__float_as_uint(x[3]) + rb));
                    };
                    st4(gp, r4(o_r));
                    st4(gp + HP, r4(o_z));
                    st4(gp + 2 * HP, r4(o_n));
                    st4(gp + 3 * HP, r4(o_hn));
                }
            }
            if (i < L) {
                if (lane == 0 && wl == 0) CPG_TL(12 + 16 * ch);
                tc::fence_proxy_async();
                tc::mbar_arrive(&bar_x[ch]);
                if (lane == 0 && wl == 0) CPG_TL(13 + 16 * ch);
                if (s > 0) prefetch(s - 1);                      // in flight under the MMAs / the other chain
                if (lane == 0 && wl == 0) CPG_TL(14 + 16 * ch);
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    CPG_TL0(52);
    if (warp == C::NW_EPI) tc::tmem_dealloc<C::TMEM_COLS>(tmem_d);
}

// chains per CTA x batch rows per chain: 4 x 16 = 64 rows per CTA (encoder, 2 directions -> 128 CTAs at B = 4096),
// 2 x 16 = 32 rows per CTA (decoder -> 128 CTAs).  More, thinner chains per SM hide each other's MMA -> TMEM ->
// shared-memory -> gate-math latency chain better than fewer, fatter ones.
#ifndef CPG_ENC_2CHAINS
using EncFwd = FwdCfg<ENC_H, ENC_H, 16, 4, 5, false>;      // four 16-row chains per CTA: 107 -> 97 us vs two 32-row chains
#else
using EncFwd = FwdCfg<ENC_H, ENC_H, 32, 2, CPG_ENC_FWD_NWG, false>;
#endif     // KID 0 | 1 (forward), 2 | 3 (backward)
using DecFwd = FwdCfg<DEC_HP, 112, 16, 2, CPG_DEC_FWD_NWG, true>;   // 13 warps per chain: one (row, quad) item per thread
#ifndef CPG_ENC_2CHAINS
using EncBwd = BwdCfg<ENC_H, 16, 4, 5, false>;             // 127 -> 119 us
#else
using EncBwd = BwdCfg<ENC_H, 32, 2, CPG_ENC_BWD_NWG, false>;
#endif
using DecBwd = BwdCfg<DEC_HP, 16, 2, CPG_DEC_NWG, true>;

template <class K>
int set_smem(K kfn, size_t bytes, size_t& set_for) {
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

int launch_gru_fwd_enc_tc(cudaStream_t s, const GruSeq* two, int B, int L, int V) {
    FwdArgs a;
    memset(&a, 0, sizeof(a));
    a.tok = two[0].tok;
    for (int d = 0; d < 2; ++d) {
        a.table[d] = two[d].table; a.whh[d] = two[d].whh; a.bhn[d] = two[d].bhn;
        a.hs[d] = two[d].hs; a.gates[d] = two[d].gates;
    }
    a.hfin = two[0].hfin;
    a.B = B; a.L = L; a.V = V; a.nprod = g_opt_matmul_terms == 1 ? 1 : 3;
    const size_t smem = EncFwd::smem_bytes(V, L);
    static size_t set_for = 0;
    if (set_smem(k_gru_fwd_tc<EncFwd>, smem, set_for)) return CPG_ECUDA;
    CPG_LAUNCH_NAMED("k_gru_fwd_enc_tc", k_gru_fwd_tc<EncFwd>, dim3(ceil_div(B, EncFwd::NCH * EncFwd::NB), 2), EncFwd::NTHREADS, smem, s, a);
    return CPG_OK;
}

int launch_gru_fwd_dec_tc(cudaStream_t s, const GruSeq& q, int B, int L, int V) {
    FwdArgs a;
    memset(&a, 0, sizeof(a));
    a.tok = q.tok; a.table[0] = q.table; a.whh[0] = q.whh; a.bhn[0] = q.bhn;
    a.rowbias = q.rowbias; a.h0 = q.h0; a.hs[0] = q.hs; a.gates[0] = q.gates;
    a.B = B; a.L = L; a.V = V; a.nprod = g_opt_matmul_terms == 1 ? 1 : 3;
    const size_t smem = DecFwd::smem_bytes(V, L);
    static size_t set_for = 0;
    if (set_smem(k_gru_fwd_tc<DecFwd>, smem, set_for)) return CPG_ECUDA;
    CPG_LAUNCH_NAMED("k_gru_fwd_dec_tc", k_gru_fwd_tc<DecFwd>, dim3(ceil_div(B, DecFwd::NCH * DecFwd::NB), 1), DecFwd::NTHREADS, smem, s, a);
    return CPG_OK;
}

int launch_gru_bwd_enc_tc(cudaStream_t s, const GruSeq* two, int B, int L, int round_dg) {
    BwdArgs a;
    memset(&a, 0, sizeof(a));
    for (int d = 0; d < 2; ++d) {
        a.whh[d] = two[d].whh; a.hs[d] = two[d].hs; a.gates[d] = two[d].gates; a.dg[d] = two[d].dg;
    }
    a.dh_fin = two[0].dh_fin;
    a.B = B; a.L = L; a.round_dg = round_dg;
    const size_t smem = EncBwd::smem_bytes();
    static size_t set_for = 0;
    if (set_smem(k_gru_bwd_tc<EncBwd>, smem, set_for)) return CPG_ECUDA;
    CPG_LAUNCH_NAMED("k_gru_bwd_enc_tc", k_gru_bwd_tc<EncBwd>, dim3(ceil_div(B, EncBwd::NCH * EncBwd::NB), 2), EncBwd::NTHREADS, smem, s, a);
    return CPG_OK;
}

int launch_gru_bwd_dec_tc(cudaStream_t s, const GruSeq& q, int B, int L, int round_dg) {
    BwdArgs a;
    memset(&a, 0, sizeof(a));
    a.whh[0] = q.whh; a.hs[0] = q.hs; a.gates[0] = q.gates; a.dg[0] = q.dg;
    a.h0 = q.h0; a.dh_out = q.dh_out; a.dh0 = q.dh0; a.drow = q.drow;
    a.B = B; a.L = L; a.round_dg = round_dg;
    const size_t smem = DecBwd::smem_bytes();
    static size_t set_for = 0;
    if (set_smem(k_gru_bwd_tc<DecBwd>, smem, set_for)) return CPG_ECUDA;
    CPG_LAUNCH_NAMED("k_gru_bwd_dec_tc", k_gru_bwd_tc<DecBwd>, dim3(ceil_div(B, DecBwd::NCH * DecBwd::NB), 1), DecBwd::NTHREADS, smem, s, a);
    return CPG_OK;
}

}  // namespace cpg
#else   // CPG_EMU: tcgen05 kernels do not exist in the CPU emulation; the selection logic never picks them there
namespace cpg {
int launch_gru_fwd_enc_tc(cudaStream_t, const GruSeq*, int, int, int) { return CPG_ECUDA; }
int launch_gru_fwd_dec_tc(cudaStream_t, const GruSeq&, int, int, int) { return CPG_ECUDA; }
int launch_gru_bwd_enc_tc(cudaStream_t, const GruSeq*, int, int, int) { return CPG_ECUDA; }
int launch_gru_bwd_dec_tc(cudaStream_t, const GruSeq&, int, int, int) { return CPG_ECUDA; }
}  // namespace cpg
#endif  // CPG_EMU
