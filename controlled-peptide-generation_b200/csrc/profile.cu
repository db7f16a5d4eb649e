// Optional per-kernel timing with CUDA events on the launching stream (used by bench.py for the
// roofline block and by developers as a poor man's launch list).  Off by default.
#include <map>
#include <string>
#include <vector>
#include <string.h>
#include <stdio.h>
#include "ctx.h"

namespace cpg {

bool g_profile_on = false;

// "name[MxNxK]" / "name[MxNxKxn]" (n products in one launch) with static storage: the per-kernel timing keeps dense
// products of different shapes apart
const char* shape_label(const char* name, int M, int N, int K, int nprod) {
    static std::map<std::string, std::string> pool;
    char buf[96];
    if (nprod > 1) snprintf(buf, sizeof(buf), "%s[%dx%dx%dx%d]", name, M, N, K, nprod);
    else snprintf(buf, sizeof(buf), "%s[%dx%dx%d]", name, M, N, K);
    auto it = pool.find(buf);
    if (it == pool.end()) it = pool.emplace(buf, buf).first;
    return it->second.c_str();
}

#ifndef CPG_EMU
struct ProfRec { const char* label; cudaEvent_t a, b; cudaStream_t s; };
static std::vector<ProfRec> g_recs;
static std::vector<cudaEvent_t> g_pool;
static cudaEvent_t g_pending_a = nullptr;
static const char* g_pending_label = nullptr;

struct TlRec { const char* label; float start_ms, dur_ms; int stream; };
static std::vector<TlRec> g_timeline;

static cudaEvent_t take_event() {
    if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
void prof_begin(const char* label, cudaStream_t s) {
    g_pending_a = take_event();
    g_pending_label = label;
    cudaEventRecord(g_pending_a, s);
}
void prof_end(cudaStream_t s) {
    cudaEvent_t b = take_event();
    cudaEventRecord(b, s);
    g_recs.push_back(ProfRec{g_pending_label, g_pending_a, b, s});
}
#endif

}  // namespace cpg

using namespace cpg;

extern "C" {

int cpg_profile_enable(int on) {
    g_profile_on = on != 0;
    return CPG_OK;
}

// Synchronises, then writes up to `cap` records "label\0" packed into `names` (name_stride bytes each),
// total milliseconds and launch counts per label; clears the recorded events.  Returns the number of labels.
int cpg_profile_read(char* names, int name_stride, float* total_ms, int* counts, int cap) {
#ifdef CPG_EMU
    (void)names; (void)name_stride; (void)total_ms; (void)counts; (void)cap;
    return 0;
#else
    cudaDeviceSynchronize();
    std::map<std::string, std::pair<double, int>> acc;
    std::vector<std::string> order;
    for (auto& r : g_recs) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        auto it = acc.find(r.label);
        if (it == acc.end()) { acc[r.label] = {ms, 1}; order.push_back(r.label); }
        else { it->second.first += ms; it->second.second += 1; }
        g_pool.push_back(r.a);
        g_pool.push_back(r.b);
    }
    // keep the raw records (start offset from the first one, duration) for cpg_profile_timeline
    g_timeline.clear();
    std::vector<cudaStream_t> streams;
    for (auto& r : g_recs) {
        int si = 0;
        while (si < (int)streams.size() && streams[si] != r.s) ++si;
        if (si == (int)streams.size()) streams.push_back(r.s);
        float t0 = 0.f, dt = 0.f;
        cudaEventElapsedTime(&t0, g_recs.front().a, r.a);
        cudaEventElapsedTime(&dt, r.a, r.b);
        g_timeline.push_back(TlRec{r.label, t0, dt, si});
    }
    g_recs.clear();
    int n = 0;
    for (auto& name : order) {
        if (n >= cap) break;
        strncpy(names + (size_t)n * name_stride, name.c_str(), name_stride - 1);
        names[(size_t)n * name_stride + name_stride - 1] = 0;
        total_ms[n] = (float)acc[name].first;
        counts[n] = acc[name].second;
        ++n;
    }
    return n;
#endif
}

// The raw records of the launches seen by the last cpg_profile_read, in launch order: label, start offset from the first
// record and duration (ms), both from CUDA events on the launching streams.  Returns the number of records written.
int cpg_profile_timeline(char* names, int name_stride, float* start_ms, float* dur_ms, int cap) {
#ifdef CPG_EMU
    (void)names; (void)name_stride; (void)start_ms; (void)dur_ms; (void)cap;
    return 0;
#else
    int n = 0;
    for (auto& r : g_timeline) {
        if (n >= cap) break;
        snprintf(names + (size_t)n * name_stride, name_stride, "S%d %s", r.stream, r.label);   // S0 = first stream seen
        start_ms[n] = r.start_ms;
        dur_ms[n] = r.dur_ms;
        ++n;
    }
    return n;
#endif
}

}  // extern "C"
