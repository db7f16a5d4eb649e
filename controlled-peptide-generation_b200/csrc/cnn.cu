// Kim-CNN sequence classifier forward (eval mode), used by RNN_VAE.forward(q_c='classifier')
// during encoding extraction (models/classifier.py:39-60 via models/model.py:135-144,186-188).
//
// conv(1 -> 100, (w, 150)) over the embedded sequence for w in {3,4,5}, relu, max over time,
// concat (300), Linear(300 -> 2).  With V <= 32 the convolution collapses to table look-ups:
//   out[f][t] = b_f + sum_{i<w} Tc_w[i][tok[t+i]][f],   Tc_w[i][v][f] = sum_e W_f[i][e] E[v][e]
// (12 V 100 table entries rebuilt from the weights), i.e. w adds per output instead of w*150 FMAs.
#include "ctx.h"

namespace cpg {
int check_launch(const char* where);

constexpr int CNN_F = 100;             // num_filters (cfg.py:293)
constexpr int CNN_WMIN = 3, CNN_NW = 3;
constexpr int CNN_ROWS = 3 + 4 + 5;    // table rows (width, tap) pairs

struct CnnWeights {
    const float* emb;                  // [V][150]
    const float* conv_w[CNN_NW];       // [100][1][w][150]
    const float* conv_b[CNN_NW];       // [100]
    const float* fc_w;                 // [2][300]
    const float* fc_b;                 // [2]
};

// table[(row)][v][f], row = tap index within the concatenated (w=3: 0..2, w=4: 3..6, w=5: 7..11)
__global__ void k_cnn_tables(CnnWeights w, int V, float* __restrict__ table) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= CNN_ROWS * V * CNN_F) return;
    int f = i % CNN_F, v = (i / CNN_F) % V, row = i / (CNN_F * V);
    int wi = row < 3 ? 0 : (row < 7 ? 1 : 2);
    int tap = row - (wi == 0 ? 0 : (wi == 1 ? 3 : 7));
    int width = CNN_WMIN + wi;
    const float* wf = w.conv_w[wi] + ((size_t)f * width + tap) * EMB;
    const float* e = w.emb + (size_t)v * EMB;
    float s = 0.f;
    for (int k = 0; k < EMB; ++k) s = fmaf(wf[k], e[k], s);
    table[i] = s;
}

// one CTA per sample; thread = (width, filter) feature
__global__ void __launch_bounds__(320)
k_cnn_fwd(const int64_t* __restrict__ tokens, int B, int L, int V, const float* __restrict__ table, CnnWeights w,
          float* __restrict__ logits) {
    __shared__ int tok[LMAX];
    __shared__ float feat[CNN_NW * CNN_F];
    __shared__ float red[2][10];
    const int b = blockIdx.x, tid = threadIdx.x;
    if (tid < L) {
        int64_t t = tokens[(size_t)b * L + tid];
        tok[tid] = (t < 0 || t >= V) ? UNK : (int)t;
    }
    __syncthreads();
    if (tid < CNN_NW * CNN_F) {
        const int wi = tid / CNN_F, f = tid % CNN_F, width = CNN_WMIN + wi;
        const int row0 = wi == 0 ? 0 : (wi == 1 ? 3 : 7);
        const float bias = w.conv_b[wi][f];
        float best = 0.f;                                   // relu output is >= 0, and T - w + 1 >= 1 windows exist
        for (int t = 0; t + width <= L; ++t) {
            float s = 0.f;
            for (int i = 0; i < width; ++i) s += __ldg(table + ((size_t)(row0 + i) * V + tok[t + i]) * CNN_F + f);
            s += bias;
            best = fmaxf(best, s);                          // max_t relu(s) = max(0, max_t s)
        }
        feat[tid] = best;
    }
    __syncthreads();
    // Linear(300 -> 2): two block reductions
    float p0 = 0.f, p1 = 0.f;
    if (tid < CNN_NW * CNN_F) {
        p0 = feat[tid] * w.fc_w[tid];
        p1 = feat[tid] * w.fc_w[CNN_NW * CNN_F + tid];
    }
    p0 = warp_sum(p0); p1 = warp_sum(p1);
    if ((tid & 31) == 0) { red[0][tid >> 5] = p0; red[1][tid >> 5] = p1; }
    __syncthreads();
    if (tid < 2) {
        float s = 0.f;
        for (int i = 0; i < 10; ++i) s += red[tid][i];
        logits[(size_t)b * 2 + tid] = s + w.fc_b[tid];
    }
}

}  // namespace cpg

using namespace cpg;

extern "C" {

int cpg_cnn_classifier_fwd(cpg_ctx* ctx, cpg_stream stream, const float* emb, const float* conv_w3, const float* conv_b3,
                           const float* conv_w4, const float* conv_b4, const float* conv_w5, const float* conv_b5,
                           const float* fc_w, const float* fc_b, int V, int B, int L, const int64_t* tokens,
                           float* table_ws, float* logits) {
    if (!ctx || !emb || !conv_w3 || !conv_w4 || !conv_w5 || !fc_w || !fc_b || !tokens || !table_ws || !logits) {
        set_error("cpg_cnn_classifier_fwd: null argument");
        return CPG_EINVAL;
    }
    if (V < 4 || V > VMAX || B < 1 || L < 5 || L > LMAX) { set_error("cpg_cnn_classifier_fwd: bad shape (needs 5 <= L <= 32)"); return CPG_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    CnnWeights w;
    w.emb = emb;
    w.conv_w[0] = conv_w3; w.conv_b[0] = conv_b3; w.conv_w[1] = conv_w4; w.conv_b[1] = conv_b4;
    w.conv_w[2] = conv_w5; w.conv_b[2] = conv_b5;
    w.fc_w = fc_w; w.fc_b = fc_b;
    int nt = CNN_ROWS * V * CNN_F;
    CPG_LAUNCH(k_cnn_tables, ceil_div(nt, 256), 256, 0, s, w, V, table_ws);
    CPG_LAUNCH(k_cnn_fwd, B, 320, 0, s, tokens, B, L, V, table_ws, w, logits);
    return check_launch("cpg_cnn_classifier_fwd");
}

}  // extern "C"
