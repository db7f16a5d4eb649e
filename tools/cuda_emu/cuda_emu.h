// Minimal single-process CUDA *source-level* emulator for the SIMT kernels in
// controlled-peptide-generation_b200/csrc.
//
// DEVELOPER TOOL ONLY -- never built, loaded or referenced by the product.  The
// build container has no GPU; this header lets the same .cu sources be compiled
// by g++ (-DCPG_EMU -x c++) so that indexing / barrier logic can be debugged
// locally (tools/cuda_emu/README.md).  Each CUDA thread of a block is a ucontext
// fiber; blocks run one after another on the calling thread.  __syncthreads and
// warp shuffles are cooperative yield points.  It models functional behaviour
// only (no memory model, no timing); tcgen05/TMA kernels are compiled out.
#pragma once
#include <ucontext.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <stdio.h>
#include <math.h>
#include <algorithm>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct __attribute__((aligned(8))) float2 { float x, y; };
struct __attribute__((aligned(16))) double2 { double x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct __attribute__((aligned(8))) uint2 { unsigned x, y; };
struct __attribute__((aligned(8))) int2 { int x, y; };
struct uchar4 { unsigned char x, y, z, w; };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline double2 make_double2(double a, double b) { return double2{a, b}; }
static inline int4 make_int4(int a, int b, int c, int d) { return int4{a, b, c, d}; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };

namespace emu {

struct Fiber {
    ucontext_t ctx;
    uint3 tid;
    bool done = false;
    int barrier_gen = 0;
};

struct Block {
    std::vector<Fiber> fibers;
    ucontext_t main_ctx;
    int cur = 0;
    int nthreads = 0;
    // block barrier
    int bar_arrived = 0, bar_gen = 0;
    // per-warp barrier + exchange
    std::vector<int> warp_arrived, warp_gen;
    std::vector<uint64_t> xchg;
    uint3 bid;
    dim3 bdim, gdim;
    unsigned char* dyn_smem = nullptr;
    std::function<void()>* body = nullptr;
};

inline Block*& cur_block() { static thread_local Block* b = nullptr; return b; }
inline Fiber& cur_fiber() { Block* b = cur_block(); return b->fibers[b->cur]; }

inline void yield() {
    Block* b = cur_block();
    Fiber& f = b->fibers[b->cur];
    swapcontext(&f.ctx, &b->main_ctx);
}

inline int alive_threads(Block* b) {
    int n = 0;
    for (auto& f : b->fibers) n += f.done ? 0 : 1;
    return n;
}

inline void syncthreads() {
    Block* b = cur_block();
    int my_gen = b->bar_gen;
    b->bar_arrived++;
    if (b->bar_arrived >= alive_threads(b)) { b->bar_arrived = 0; b->bar_gen++; return; }
    while (b->bar_gen == my_gen) yield();
}

inline int warp_lanes(Block* b, int warp) {
    int lo = warp * 32, hi = std::min(b->nthreads, lo + 32), n = 0;
    for (int i = lo; i < hi; ++i) n += b->fibers[i].done ? 0 : 1;
    return n;
}

inline void syncwarp() {
    Block* b = cur_block();
    int w = b->cur / 32;
    int my_gen = b->warp_gen[w];
    b->warp_arrived[w]++;
    if (b->warp_arrived[w] >= warp_lanes(b, w)) { b->warp_arrived[w] = 0; b->warp_gen[w]++; return; }
    while (b->warp_gen[w] == my_gen) yield();
}

template <typename T>
inline T shfl_idx(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    Block* b = cur_block();
    int w = b->cur / 32, lane = b->cur % 32;
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    b->xchg[w * 32 + lane] = bits;
    syncwarp();
    int lo = w * 32;
    int src = lo + (src_lane & 31);
    if (src >= b->nthreads) src = b->cur;
    uint64_t got = b->xchg[src];
    syncwarp();
    T out;
    memcpy(&out, &got, sizeof(T));
    return out;
}

inline void fiber_entry() {
    Block* b = cur_block();
    (*b->body)();
    b->fibers[b->cur].done = true;
    // a finished thread may complete a pending barrier
    if (b->bar_arrived > 0 && b->bar_arrived >= alive_threads(b)) { b->bar_arrived = 0; b->bar_gen++; }
    int w = b->cur / 32;
    if (b->warp_arrived[w] > 0 && b->warp_arrived[w] >= warp_lanes(b, w)) { b->warp_arrived[w] = 0; b->warp_gen[w]++; }
    swapcontext(&b->fibers[b->cur].ctx, &b->main_ctx);
}

static const size_t kStack = 256 * 1024;

inline void launch(dim3 grid, dim3 block, size_t smem, std::function<void()> body) {
    int nthreads = block.x * block.y * block.z;
    static thread_local std::vector<unsigned char> stacks;
    if (stacks.size() < (size_t)nthreads * kStack) stacks.resize((size_t)nthreads * kStack);
    std::vector<unsigned char> dyn(smem + 64);
    Block b;
    b.nthreads = nthreads;
    b.bdim = block;
    b.gdim = grid;
    b.body = &body;
    b.dyn_smem = (unsigned char*)(((uintptr_t)dyn.data() + 63) & ~(uintptr_t)63);
    int nwarps = (nthreads + 31) / 32;
    Block* saved = cur_block();
    cur_block() = &b;
    for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) {
        b.bid = uint3{bx, by, bz};
        b.fibers.assign(nthreads, Fiber());
        b.bar_arrived = 0; b.bar_gen = 0;
        b.warp_arrived.assign(nwarps, 0);
        b.warp_gen.assign(nwarps, 0);
        b.xchg.assign(nwarps * 32, 0);
        for (int t = 0; t < nthreads; ++t) {
            Fiber& f = b.fibers[t];
            f.tid = uint3{(unsigned)(t % block.x), (unsigned)((t / block.x) % block.y),
                          (unsigned)(t / (block.x * block.y))};
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = stacks.data() + (size_t)t * kStack;
            f.ctx.uc_stack.ss_size = kStack;
            f.ctx.uc_link = &b.main_ctx;
            makecontext(&f.ctx, (void (*)())fiber_entry, 0);
        }
        int remaining = nthreads;
        while (remaining > 0) {
            remaining = 0;
            for (int t = 0; t < nthreads; ++t) {
                if (b.fibers[t].done) continue;
                b.cur = t;
                swapcontext(&b.main_ctx, &b.fibers[t].ctx);
                if (!b.fibers[t].done) remaining++;
            }
        }
    }
    cur_block() = saved;
}

}  // namespace emu

#define threadIdx (emu::cur_fiber().tid)
#define blockIdx (emu::cur_block()->bid)
#define blockDim (emu::cur_block()->bdim)
#define gridDim (emu::cur_block()->gdim)
#define warpSize 32

static inline void __syncthreads() { emu::syncthreads(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::syncwarp(); }
template <typename T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return emu::shfl_idx(v, src); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) {
    return emu::shfl_idx(v, (emu::cur_block()->cur % 32) ^ m);
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
    int lane = emu::cur_block()->cur % 32;
    return emu::shfl_idx(v, lane + (int)d < 32 ? lane + (int)d : lane);
}
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
    int lane = emu::cur_block()->cur % 32;
    return emu::shfl_idx(v, lane - (int)d >= 0 ? lane - (int)d : lane);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) r |= (emu::shfl_idx(pred ? 1 : 0, l) ? 1u : 0u) << l;
    return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) {
    emu::Block* b = emu::cur_block();
    int w = b->cur / 32;
    int lanes = std::min(32, b->nthreads - w * 32);
    unsigned full = lanes == 32 ? 0xffffffffu : ((1u << lanes) - 1);
    return (__ballot_sync(m, pred) & full) == full;
}

template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { auto o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicMax(T* p, T v) { T o = *p; *p = std::max(o, v); return o; }
template <typename T> static inline T atomicMin(T* p, T v) { T o = *p; *p = std::min(o, v); return o; }
template <typename T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <typename T> static inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

// glibc already declares __expf & co as extern symbols: map the CUDA fast intrinsics by macro
#define __expf(x) expf(x)
#define __logf(x) logf(x)
#define __sinf(x) sinf(x)
#define __cosf(x) cosf(x)
#define __sincosf(x, s, c) sincosf((x), (s), (c))
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __frcp_rn(float x) { return 1.0f / x; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline float __saturatef(float x) { return x < 0 ? 0 : (x > 1 ? 1 : x); }
using std::max;
using std::min;

static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
