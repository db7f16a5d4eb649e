// The stages of a CLaSS sampling round that follow the accept test, kept on the device so that only the unique
// accepted peptides ever cross PCIe:
//
//   cpg_compact_accepted     indices of the accepted draws, ascending (stable stream compaction of the accept mask;
//                            reference: `samples[accept]`-style boolean indexing on the host, sample_pipeline.py:195-207)
//   cpg_gather_rows          rows of a [n][D] fp32 matrix at those indices
//   cpg_dedup_rows           first occurrence of every distinct token row (pandas `drop_duplicates()` on the decoded
//                            peptide strings, sample_pipeline.py:312-313) -- open-addressing hash table of row indices,
//                            exact (rows are compared, the hash only picks the probe start), deterministic (the lowest
//                            index of a group wins whatever the thread order)
//   cpg_peptide_descriptors  H, uH, charge of every peptide (modlamp GlobalAnalysis.calc_H / calc_uH / calc_charge as
//                            called by compute_modlamp, sample_pipeline.py:210-218): mean hydrophobicity, hydrophobic
//                            moment at 100 degrees over the whole sequence, net charge from per-residue partial charges
//
// Byte / index work: HBM-bound streaming, one pass over the mask / the token rows.
#include "ctx.h"

namespace cpg {
int check_launch(const char* where);

// ------------------------------------------------------------------------------------- compaction
constexpr int CP_BLOCK = 256, CP_PER = 8, CP_TILE = CP_BLOCK * CP_PER;       // 2048 flags per CTA

__global__ void __launch_bounds__(CP_BLOCK)
k_compact_count(const uint8_t* __restrict__ flag, int64_t n, int* __restrict__ tile_count) {
    __shared__ int wsum[CP_BLOCK / 32];
    const int64_t base = (int64_t)blockIdx.x * CP_TILE;
    int c = 0;
#pragma unroll
    for (int k = 0; k < CP_PER; ++k) {
        const int64_t i = base + k * CP_BLOCK + threadIdx.x;
        c += (i < n && flag[i]) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < CP_BLOCK / 32; ++w) s += wsum[w];
        tile_count[blockIdx.x] = s;
    }
}

// exclusive scan of the tile counts (one CTA; ntiles <= a few 10^4), total -> *count
__global__ void __launch_bounds__(1024)
k_compact_scan(const int* __restrict__ tile_count, int ntiles, int64_t* __restrict__ tile_off, unsigned long long* __restrict__ count) {
    __shared__ long long part[1024];
    const int t = threadIdx.x;
    const int per = (ntiles + 1023) / 1024;
    const int lo = t * per, hi = min(ntiles, lo + per);
    long long s = 0;
    for (int i = lo; i < hi; ++i) s += tile_count[i];
    part[t] = s;
    __syncthreads();
    if (t == 0) {
        long long run = 0;
        for (int i = 0; i < 1024; ++i) { long long v = part[i]; part[i] = run; run += v; }
        *count = (unsigned long long)run;
    }
    __syncthreads();
    long long run = part[t];
    for (int i = lo; i < hi; ++i) { tile_off[i] = run; run += tile_count[i]; }
}

__global__ void __launch_bounds__(CP_BLOCK)
k_compact_scatter(const uint8_t* __restrict__ flag, int64_t n, const int64_t* __restrict__ tile_off, int64_t first_index,
                  int64_t cap, int64_t* __restrict__ idx_out) {
    __shared__ int wbase[CP_BLOCK / 32];
    __shared__ int run_s;
    const int64_t base = (int64_t)blockIdx.x * CP_TILE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) run_s = 0;
    __syncthreads();
    for (int k = 0; k < CP_PER; ++k) {                      // index order: k-major, then thread -> ascending i
        const int64_t i = base + k * CP_BLOCK + threadIdx.x;
        const int f = (i < n && flag[i]) ? 1 : 0;
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (lane == 0) wbase[warp] = __popc(m);
        __syncthreads();
        int before = run_s;
        for (int w = 0; w < warp; ++w) before += wbase[w];
        if (f) {
            const int64_t pos = tile_off[blockIdx.x] + before + __popc(m & ((1u << lane) - 1u));
            if (pos < cap) idx_out[pos] = first_index + i;
        }
        __syncthreads();
        if (threadIdx.x == 0) { int s = 0; for (int w = 0; w < CP_BLOCK / 32; ++w) s += wbase[w]; run_s += s; }
        __syncthreads();
    }
}

__global__ void k_gather_rows(const float* __restrict__ src, const int64_t* __restrict__ idx, int64_t index_base, int64_t m,
                              int D, float* __restrict__ dst) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t row = t / D;
    if (row >= m) return;
    const int d = (int)(t - row * D);
    dst[row * D + d] = src[(idx[row] - index_base) * D + d];
}

// ------------------------------------------------------------------------------------- dedup
__device__ __forceinline__ uint64_t mix64(uint64_t x) {          // splitmix64 finaliser
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
    return x;
}
__device__ __forceinline__ uint64_t row_hash(const int* __restrict__ r, int W) {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (int j = 0; j < W; ++j) h = mix64(h ^ (uint64_t)(uint32_t)r[j]);
    return h;
}
__device__ __forceinline__ bool rows_equal(const int* __restrict__ a, const int* __restrict__ b, int W) {
    for (int j = 0; j < W; ++j) if (a[j] != b[j]) return false;
    return true;
}
// table[slot] = lowest row index seen with that slot's row content (-1 = empty).  A slot's representative can only
// be replaced by a row with the SAME content, so comparing against whichever representative is there is exact.
__global__ void k_dedup_insert(const int* __restrict__ rows, int64_t n, int W, int* __restrict__ table, uint32_t mask) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int* r = rows + i * W;
    uint32_t s = (uint32_t)row_hash(r, W) & mask;
    for (;;) {
        int cur = table[s];
        if (cur < 0) {
            const int prev = atomicCAS(&table[s], -1, (int)i);
            if (prev < 0) return;
            cur = prev;
        }
        if (cur == (int)i) return;
        if (rows_equal(rows + (int64_t)cur * W, r, W)) { atomicMin(&table[s], (int)i); return; }
        s = (s + 1) & mask;
    }
}
__global__ void k_dedup_lookup(const int* __restrict__ rows, int64_t n, int W, const int* __restrict__ table, uint32_t mask,
                               int* __restrict__ first_idx, uint8_t* __restrict__ is_first) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int* r = rows + i * W;
    uint32_t s = (uint32_t)row_hash(r, W) & mask;
    for (;;) {
        const int cur = table[s];
        if (cur == (int)i || rows_equal(rows + (int64_t)cur * W, r, W)) {
            if (first_idx != nullptr) first_idx[i] = cur;
            if (is_first != nullptr) is_first[i] = cur == (int)i ? 1 : 0;
            return;
        }
        s = (s + 1) & mask;
    }
}

// ------------------------------------------------------------------------------------- descriptors
struct DescTables {
    float hyd[32];          // hydrophobicity per residue code (0..19 used)
    double charge[32];      // partial charge of the side chain at the chosen pH
    double charge_ends;     // N-terminus + C-terminus
    float cs[LMAX + 2], sn[LMAX + 2];     // cos / sin of position * angle
    int8_t aa_of_token[64]; // token id -> residue code, -1 = not a residue (skipped: <start>, <eos>, <pad>, <unk>)
};
// one thread per peptide: tokens [n][W] (entries < 0 or non-residue tokens are skipped)
__global__ void k_peptide_descriptors(const int* __restrict__ tok, int64_t n, int W, DescTables T, float* __restrict__ H,
                                      float* __restrict__ uH, float* __restrict__ charge, int* __restrict__ length) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int* r = tok + i * W;
    float sh = 0.f, sc = 0.f, ss = 0.f;
    double q = T.charge_ends;
    int len = 0;
    for (int j = 0; j < W; ++j) {
        const int t = r[j];
        if (t < 0 || t >= 64) continue;
        const int a = T.aa_of_token[t];
        if (a < 0) continue;
        const float h = T.hyd[a];
        sh += h;
        if (len < LMAX + 2) { sc = fmaf(h, T.cs[len], sc); ss = fmaf(h, T.sn[len], ss); }
        q += T.charge[a];
        ++len;
    }
    const float inv = len > 0 ? 1.0f / (float)len : nanf("");
    H[i] = sh * inv;
    uH[i] = sqrtf(sc * sc + ss * ss) * inv;
    charge[i] = (float)(rint(q * 1000.0) / 1000.0);          // modlamp rounds the charge to 3 decimals
    if (length != nullptr) length[i] = len;
}

// ------------------------------------------------------------------------------------- data feed
// One batch of the weighted random iterator (data_processing/dataset.py:60-77: torch.multinomial(weights, B,
// replacement=True), then the padded token rows of the chosen examples): row i of the batch takes a Philox uniform,
// finds its example by binary search in the cumulative sampling weights and copies the example's token row.
__global__ void k_feed_batch(const uint8_t* __restrict__ tokens, const double* __restrict__ cdf, int64_t n_examples, int L,
                             uint64_t seed, uint64_t step, int B, int64_t* __restrict__ out, int64_t* __restrict__ out_index) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);          // one warp per batch row
    const int lane = threadIdx.x & 31;
    if (i >= B) return;
    int64_t ex = 0;
    if (lane == 0) {
        uint32_t r[4];
        Philox::gen(seed, step * (uint64_t)B + (uint64_t)i, 0x80000001u, r);
        const double u = u64_to_unit(r[0], r[1]) * cdf[n_examples - 1];
        int64_t lo = 0, hi = n_examples - 1;                                    // first example with cdf > u
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (cdf[mid] > u) hi = mid; else lo = mid + 1; }
        ex = lo;
        if (out_index != nullptr) out_index[i] = ex;
    }
    ex = __shfl_sync(0xffffffffu, ex, 0);
    for (int t = lane; t < L; t += 32) out[(size_t)i * L + t] = (int64_t)tokens[ex * L + t];
}

}  // namespace cpg

using namespace cpg;

extern "C" {

int cpg_compact_accepted(cpg_ctx* ctx, cpg_stream stream, const uint8_t* accept, int64_t n, int64_t first_index, int64_t cap,
                         int64_t* idx_out, unsigned long long* count) {
    if (!ctx || !accept || !idx_out || !count || n < 1 || cap < 0) { set_error("cpg_compact_accepted: bad argument"); return CPG_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t ntiles = (n + CP_TILE - 1) / CP_TILE;
    if (ntiles > (1 << 24)) { set_error("cpg_compact_accepted: n too large for one call (<= 2^35 flags)"); return CPG_EINVAL; }
    int rc = ensure_aux(ctx, (size_t)ntiles * (sizeof(int) + sizeof(int64_t)) + 256, s);
    if (rc) return rc;
    int64_t* tile_off = (int64_t*)ctx->aux;
    int* tile_count = (int*)(tile_off + ntiles);
    CPG_LAUNCH(k_compact_count, (unsigned)ntiles, CP_BLOCK, 0, s, accept, n, tile_count);
    CPG_LAUNCH(k_compact_scan, 1, 1024, 0, s, tile_count, (int)ntiles, tile_off, count);
    CPG_LAUNCH(k_compact_scatter, (unsigned)ntiles, CP_BLOCK, 0, s, accept, n, tile_off, first_index, cap, idx_out);
    return check_launch("cpg_compact_accepted");
}

int cpg_gather_rows(cpg_ctx* ctx, cpg_stream stream, const float* src, const int64_t* idx, int64_t index_base, int64_t m, int D,
                    float* dst) {
    if (!ctx || !src || !idx || !dst || m < 0 || D < 1) { set_error("cpg_gather_rows: bad argument"); return CPG_EINVAL; }
    if (m == 0) return CPG_OK;
    const int64_t total = m * D;
    CPG_LAUNCH(k_gather_rows, (unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream, src, idx, index_base, m, D, dst);
    return check_launch("cpg_gather_rows");
}

int cpg_dedup_rows(cpg_ctx* ctx, cpg_stream stream, const int* rows, int64_t n, int width, int* first_index, uint8_t* is_first) {
    if (!ctx || !rows || n < 1 || width < 1 || (!first_index && !is_first)) { set_error("cpg_dedup_rows: bad argument"); return CPG_EINVAL; }
    if (n > (1ll << 30)) { set_error("cpg_dedup_rows: at most 2^30 rows per call"); return CPG_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    uint64_t cap = 64;
    while (cap < (uint64_t)n * 2) cap <<= 1;
    int rc = ensure_aux(ctx, cap * sizeof(int), s);
    if (rc) return rc;
    int* table = (int*)ctx->aux;
#ifdef CPG_EMU
    memset(table, 0xff, cap * sizeof(int));
#else
    cudaMemsetAsync(table, 0xff, cap * sizeof(int), s);
#endif
    CPG_LAUNCH(k_dedup_insert, (unsigned)((n + 255) / 256), 256, 0, s, rows, n, width, table, (uint32_t)(cap - 1));
    CPG_LAUNCH(k_dedup_lookup, (unsigned)((n + 255) / 256), 256, 0, s, rows, n, width, table, (uint32_t)(cap - 1), first_index, is_first);
    return check_launch("cpg_dedup_rows");
}

int cpg_peptide_descriptors(cpg_ctx* ctx, cpg_stream stream, const int* tokens, int64_t n, int width, const int8_t* aa_of_token,
                            int n_tokens, const float* hydrophobicity20, const double* side_chain_charge20, double charge_ends,
                            float angle_deg, float* H, float* uH, float* charge, int* length) {
    if (!ctx || !tokens || !aa_of_token || !hydrophobicity20 || !side_chain_charge20 || !H || !uH || !charge || n < 1 || width < 1) {
        set_error("cpg_peptide_descriptors: bad argument"); return CPG_EINVAL;
    }
    if (n_tokens < 1 || n_tokens > 64) { set_error("cpg_peptide_descriptors: 1 <= n_tokens <= 64"); return CPG_EINVAL; }
    DescTables T;
    memset(&T, 0, sizeof(T));
    for (int a = 0; a < 20; ++a) { T.hyd[a] = hydrophobicity20[a]; T.charge[a] = side_chain_charge20[a]; }
    T.charge_ends = charge_ends;
    for (int i = 0; i < LMAX + 2; ++i) {
        const double rad = (double)i * (double)angle_deg * 3.141592653589793 / 180.0;
        T.cs[i] = (float)cos(rad);
        T.sn[i] = (float)sin(rad);
    }
    for (int t = 0; t < 64; ++t) T.aa_of_token[t] = t < n_tokens ? aa_of_token[t] : (int8_t)-1;
    CPG_LAUNCH(k_peptide_descriptors, (unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream, tokens, n, width, T, H, uH, charge, length);
    return check_launch("cpg_peptide_descriptors");
}

int cpg_feed_batch(cpg_ctx* ctx, cpg_stream stream, const uint8_t* tokens, const double* cdf, int64_t n_examples, int L,
                   uint64_t seed, uint64_t step, int B, int64_t* out_tokens, int64_t* out_index) {
    if (!ctx || !tokens || !cdf || !out_tokens || n_examples < 1 || L < 1 || B < 1) { set_error("cpg_feed_batch: bad argument"); return CPG_EINVAL; }
    CPG_LAUNCH(k_feed_batch, (unsigned)((B + 7) / 8), 256, 0, (cudaStream_t)stream, tokens, cdf, n_examples, L, seed, step, B, out_tokens, out_index);
    return check_launch("cpg_feed_batch");
}

}  // extern "C"
