// GRU recurrences (forward with activation stash, and BPTT) for the bi-GRU encoder
// and the teacher-forced GRU decoder.
//
// Replaces the cuDNN/ATen nn.GRU calls at models/encoder.py:25-30,42 and
// models/decoder.py:40,77 (gate order r,z,n; h' = (1-z) n + z h).  The input-side
// projection arrives as a token->gate table (prep.cu) plus, for the decoder, a
// per-row constant [z;c] W_ih[:,150:]^T, so one step is a [R x H] x [H x 3H]
// contraction against W_hh^T held in shared memory for all L steps.  Samples are
// independent: the batch is tiled over CTAs (R = 32 rows), no grid-wide sync.
//
// Thread tile: 4 rows x 4 hidden units x 3 gates (48 fp32 accumulators); h lives in
// shared memory k-major with an XOR swizzle on 4-row chunks so that both the
// broadcast reads and the 128-bit writes are bank-conflict free.
#include "kernels.h"

namespace cpg {

struct GruSeqPair { GruSeq s[2]; };

// Gate non-linearities on the SFU (ex2.approx + approximate reciprocal): absolute error <= ~6e-7,
// i.e. fp32-rounding level, at a third of the instruction count of expf()/tanhf() -- the gate math
// is ~25 % of the forward recurrence's instructions.  (The decode kernels keep the exact versions.)
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }

template <int HP, int R>
struct GruCfg {
    static constexpr int NU = HP / 4;
    static constexpr int NRG = R / 4;
    static constexpr int NT = NU * NRG;
    static constexpr int G = 3 * HP;
    static constexpr size_t SMEM_FWD = (size_t)(HP * G + 2 * HP * R) * sizeof(float);
    static constexpr size_t SMEM_BWD = (size_t)(G * HP + 2 * G * R) * sizeof(float);
};

// offset of the 4-row chunk `ty` inside row-vector k of a [K][R] swizzled tile
template <int R>
__device__ __forceinline__ int swz(int k, int ty) { return k * R + ((ty ^ ((k >> 2) & (R / 4 - 1))) << 2); }

template <int HP, int R>
__global__ void __launch_bounds__(GruCfg<HP, R>::NT)
k_gru_fwd(GruSeqPair pr, int B, int L) {
    using C = GruCfg<HP, R>;
    constexpr int G = C::G;
    const GruSeq& q = pr.s[blockIdx.y];
    CPG_DYN_SMEM(float, smem);
    float* Wt = smem;                 // [HP][G]
    float* hb = smem + HP * G;        // [2][HP][R] swizzled
    const int tid = threadIdx.x;
    const int tx = tid % C::NU, ty = tid / C::NU;
    const int j0 = 4 * tx, r0 = 4 * ty;
    const int row0 = blockIdx.x * R;

    for (int i = tid * 4; i < HP * G; i += C::NT * 4) st4(Wt + i, ld4(q.whh_t + i));

    float hprev[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int row = row0 + r0 + i;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q.h0 != nullptr && row < B) v = ld4(q.h0 + (size_t)row * HP + j0);
        hprev[i][0] = v.x; hprev[i][1] = v.y; hprev[i][2] = v.z; hprev[i][3] = v.w;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
        st4(hb + swz<R>(j0 + u, ty), make_float4(hprev[0][u], hprev[1][u], hprev[2][u], hprev[3][u]));
    const float4 bhn4 = ld4(q.bhn + j0);
    const float bhn[4] = {bhn4.x, bhn4.y, bhn4.z, bhn4.w};
    __syncthreads();

    for (int s = 0; s < L; ++s) {
        const int t = q.reverse ? (L - 1 - s) : s;
        const float* hcur = hb + (s & 1) * HP * R;
        float* hnxt = hb + ((s & 1) ^ 1) * HP * R;

        // input-side pre-activations for this thread's 4 rows x 4 units x 3 gates
        float gi[4][3][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int row = min(row0 + r0 + i, B - 1);
            int tk = q.tok[(size_t)row * L + t];
            const float* base = q.table + (size_t)tk * G + j0;
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                float4 v = __ldg(reinterpret_cast<const float4*>(base + g * HP));
                if (q.rowbias != nullptr) {
                    float4 b = __ldg(reinterpret_cast<const float4*>(q.rowbias + (size_t)row * G + g * HP + j0));
                    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
                }
                gi[i][g][0] = v.x; gi[i][g][1] = v.y; gi[i][g][2] = v.z; gi[i][g][3] = v.w;
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
            const float4 hv = ld4(hcur + swz<R>(k, ty));
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
        for (int i = 0; i < 4; ++i) {
            const int row = row0 + r0 + i;
            float rr[4], zz[4], nn[4], hh[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                rr[u] = sigmoid_fast(gi[i][0][u] + acc[i][0][u]);
                zz[u] = sigmoid_fast(gi[i][1][u] + acc[i][1][u]);
                hh[u] = acc[i][2][u] + bhn[u];
                nn[u] = tanh_fast(gi[i][2][u] + rr[u] * hh[u]);
                hnew[i][u] = (1.0f - zz[u]) * nn[u] + zz[u] * hprev[i][u];
                hprev[i][u] = hnew[i][u];
            }
            if (row < B) {
                const size_t bs = (size_t)row * L + s;
                if (q.hs != nullptr)
                    st4(q.hs + bs * HP + j0, make_float4(hnew[i][0], hnew[i][1], hnew[i][2], hnew[i][3]));
                if (q.gates != nullptr) {
                    float* gp = q.gates + bs * 4 * HP + j0;
                    st4(gp, make_float4(rr[0], rr[1], rr[2], rr[3]));
                    st4(gp + HP, make_float4(zz[0], zz[1], zz[2], zz[3]));
                    st4(gp + 2 * HP, make_float4(nn[0], nn[1], nn[2], nn[3]));
                    st4(gp + 3 * HP, make_float4(hh[0], hh[1], hh[2], hh[3]));
                }
                if (s == L - 1 && q.hfin != nullptr) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) q.hfin[(size_t)row * q.hfin_stride + j0 + u] = hnew[i][u];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            st4(hnxt + swz<R>(j0 + u, ty), make_float4(hnew[0][u], hnew[1][u], hnew[2][u], hnew[3][u]));
        __syncthreads();
    }
}

// BPTT.  Per step (reverse step order):
//   dht = dh + dh_out[s];  dn = dht (1-z); dz = dht (h_prev - n); dh_prev = dht z
//   dn_pre = dn (1-n^2); dr = dn_pre hn; dhn = dn_pre r; dr_pre = dr r (1-r); dz_pre = dz z (1-z)
//   dh_prev += [dr_pre, dz_pre, dhn] W_hh
// and the four planes (dr_pre, dz_pre, dn_pre, dhn) are written out for the weight-gradient
// contractions (wgrad.cu).  W_hh (natural [3H][H]) stays in shared memory.
template <int HP, int R, bool ROWSUM>
__global__ void __launch_bounds__(GruCfg<HP, R>::NT)
k_gru_bwd(GruSeqPair pr, int B, int L) {
    using C = GruCfg<HP, R>;
    constexpr int G = C::G;
    const GruSeq& q = pr.s[blockIdx.y];
    CPG_DYN_SMEM(float, smem);
    float* W = smem;                  // [G][HP]
    float* db = smem + G * HP;        // [2][G][R] swizzled
    const int tid = threadIdx.x;
    const int tx = tid % C::NU, ty = tid / C::NU;
    const int j0 = 4 * tx, r0 = 4 * ty;
    const int row0 = blockIdx.x * R;

    for (int i = tid * 4; i < G * HP; i += C::NT * 4) st4(W + i, ld4(q.whh + i));

    float dh[4][4];
    float rs[ROWSUM ? 4 : 1][3][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int u = 0; u < 4; ++u) dh[i][u] = 0.f;
    if (ROWSUM) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int g = 0; g < 3; ++g)
#pragma unroll
                for (int u = 0; u < 4; ++u) rs[ROWSUM ? i : 0][g][u] = 0.f;
    }

    // prefetch registers: gates (r,z,n,hn), h_prev, dh_out for the step about to be processed
    float4 pg[4][4], ph[4], pd[4];
    auto prefetch = [&](int s) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = row0 + r0 + i;
            const bool ok = row < B;
            const size_t bs = (size_t)(ok ? row : 0) * L + s;
            const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int pl = 0; pl < 4; ++pl)
                pg[i][pl] = ok ? __ldg(reinterpret_cast<const float4*>(q.gates + (bs * 4 + pl) * HP + j0)) : zero;
            if (s > 0) ph[i] = ok ? __ldg(reinterpret_cast<const float4*>(q.hs + (bs - 1) * HP + j0)) : zero;
            else ph[i] = (ok && q.h0 != nullptr) ? __ldg(reinterpret_cast<const float4*>(q.h0 + (size_t)row * HP + j0)) : zero;
            pd[i] = (ok && q.dh_out != nullptr) ? __ldg(reinterpret_cast<const float4*>(q.dh_out + bs * HP + j0)) : zero;
            if (s == L - 1 && ok && q.dh_fin != nullptr) {
                const float* f = q.dh_fin + (size_t)row * q.dh_fin_stride + j0;
                pd[i].x += f[0]; pd[i].y += f[1]; pd[i].z += f[2]; pd[i].w += f[3];
            }
        }
    };
    prefetch(L - 1);
    __syncthreads();

    for (int s = L - 1; s >= 0; --s) {
        float* dcur = db + (s & 1) * G * R;
        float acc[4][4];
        float dgr[4][4], dgz[4][4], dgn[4][4];       // [u][i] : transposed for the float4 row stores
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = row0 + r0 + i;
            const float r4[4] = {pg[i][0].x, pg[i][0].y, pg[i][0].z, pg[i][0].w};
            const float z4[4] = {pg[i][1].x, pg[i][1].y, pg[i][1].z, pg[i][1].w};
            const float n4[4] = {pg[i][2].x, pg[i][2].y, pg[i][2].z, pg[i][2].w};
            const float hn4[4] = {pg[i][3].x, pg[i][3].y, pg[i][3].z, pg[i][3].w};
            const float hp4[4] = {ph[i].x, ph[i].y, ph[i].z, ph[i].w};
            const float do4[4] = {pd[i].x, pd[i].y, pd[i].z, pd[i].w};
            float o_r[4], o_z[4], o_n[4], o_hn[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float dht = dh[i][u] + do4[u];
                const float dn = dht * (1.0f - z4[u]);
                const float dz = dht * (hp4[u] - n4[u]);
                acc[i][u] = dht * z4[u];
                const float dn_pre = dn * (1.0f - n4[u] * n4[u]);
                const float dr = dn_pre * hn4[u];
                o_hn[u] = dn_pre * r4[u];
                o_r[u] = dr * r4[u] * (1.0f - r4[u]);
                o_z[u] = dz * z4[u] * (1.0f - z4[u]);
                o_n[u] = dn_pre;
                dgr[u][i] = o_r[u]; dgz[u][i] = o_z[u]; dgn[u][i] = o_hn[u];
                if (ROWSUM) {
                    rs[ROWSUM ? i : 0][0][u] += o_r[u];
                    rs[ROWSUM ? i : 0][1][u] += o_z[u];
                    rs[ROWSUM ? i : 0][2][u] += o_n[u];
                }
            }
            if (row < B) {
                float* gp = q.dg + ((size_t)row * L + s) * 4 * HP + j0;
                st4(gp, make_float4(o_r[0], o_r[1], o_r[2], o_r[3]));
                st4(gp + HP, make_float4(o_z[0], o_z[1], o_z[2], o_z[3]));
                st4(gp + 2 * HP, make_float4(o_n[0], o_n[1], o_n[2], o_n[3]));
                st4(gp + 3 * HP, make_float4(o_hn[0], o_hn[1], o_hn[2], o_hn[3]));
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            st4(dcur + swz<R>(j0 + u, ty), make_float4(dgr[u][0], dgr[u][1], dgr[u][2], dgr[u][3]));
            st4(dcur + swz<R>(HP + j0 + u, ty), make_float4(dgz[u][0], dgz[u][1], dgz[u][2], dgz[u][3]));
            st4(dcur + swz<R>(2 * HP + j0 + u, ty), make_float4(dgn[u][0], dgn[u][1], dgn[u][2], dgn[u][3]));
        }
        __syncthreads();
        if (s > 0) prefetch(s - 1);          // in flight during the contraction below

#pragma unroll 4
        for (int g = 0; g < G; ++g) {
            const float4 dv = ld4(dcur + swz<R>(g, ty));
            const float4 wv = ld4(W + g * HP + j0);
            const float d4[4] = {dv.x, dv.y, dv.z, dv.w};
            const float w4[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int u = 0; u < 4; ++u) acc[i][u] = fmaf(d4[i], w4[u], acc[i][u]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int u = 0; u < 4; ++u) dh[i][u] = acc[i][u];
        // double-buffered dg tile: the next step writes the other buffer, and the barrier of
        // that step orders this step's reads before the step after next overwrites this one.
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = row0 + r0 + i;
        if (row >= B) continue;
        if (q.dh0 != nullptr) st4(q.dh0 + (size_t)row * HP + j0, make_float4(dh[i][0], dh[i][1], dh[i][2], dh[i][3]));
        if (ROWSUM && q.drow != nullptr) {
#pragma unroll
            for (int g = 0; g < 3; ++g)
                st4(q.drow + (size_t)row * G + g * HP + j0,
                    make_float4(rs[ROWSUM ? i : 0][g][0], rs[ROWSUM ? i : 0][g][1], rs[ROWSUM ? i : 0][g][2],
                                rs[ROWSUM ? i : 0][g][3]));
        }
    }
}

constexpr int GRU_R = 32;

void launch_gru_fwd_enc(cudaStream_t s, const GruSeq* two, int B, int L) {
    using C = GruCfg<ENC_H, GRU_R>;
    GruSeqPair pr; pr.s[0] = two[0]; pr.s[1] = two[1];
    auto kfn = k_gru_fwd<ENC_H, GRU_R>;
    CPG_SET_MAX_SMEM(kfn, C::SMEM_FWD);
    CPG_LAUNCH_NAMED("k_gru_fwd_enc", kfn, dim3(ceil_div(B, GRU_R), 2), C::NT, C::SMEM_FWD, s, pr, B, L);
}
void launch_gru_fwd_dec(cudaStream_t s, const GruSeq& seq, int B, int L) {
    using C = GruCfg<DEC_HP, GRU_R>;
    GruSeqPair pr; pr.s[0] = seq; pr.s[1] = seq;
    auto kfn = k_gru_fwd<DEC_HP, GRU_R>;
    CPG_SET_MAX_SMEM(kfn, C::SMEM_FWD);
    CPG_LAUNCH_NAMED("k_gru_fwd_dec", kfn, dim3(ceil_div(B, GRU_R), 1), C::NT, C::SMEM_FWD, s, pr, B, L);
}
void launch_gru_bwd_enc(cudaStream_t s, const GruSeq* two, int B, int L) {
    using C = GruCfg<ENC_H, GRU_R>;
    GruSeqPair pr; pr.s[0] = two[0]; pr.s[1] = two[1];
    auto kfn = k_gru_bwd<ENC_H, GRU_R, false>;
    CPG_SET_MAX_SMEM(kfn, C::SMEM_BWD);
    CPG_LAUNCH_NAMED("k_gru_bwd_enc", kfn, dim3(ceil_div(B, GRU_R), 2), C::NT, C::SMEM_BWD, s, pr, B, L);
}
void launch_gru_bwd_dec(cudaStream_t s, const GruSeq& seq, int B, int L) {
    using C = GruCfg<DEC_HP, GRU_R>;
    GruSeqPair pr; pr.s[0] = seq; pr.s[1] = seq;
    auto kfn = k_gru_bwd<DEC_HP, GRU_R, true>;
    CPG_SET_MAX_SMEM(kfn, C::SMEM_BWD);
    CPG_LAUNCH_NAMED("k_gru_bwd_dec", kfn, dim3(ceil_div(B, GRU_R), 1), C::NT, C::SMEM_BWD, s, pr, B, L);
}

}  // namespace cpg
