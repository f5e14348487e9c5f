// tools/emu/cuda_runtime.h — TEST INFRASTRUCTURE, never part of libzrt.
//
// A stand-in for <cuda_runtime.h> that lets g++ compile zraytrace_b200/csrc/*.cu (with -DZRT_EMU -D__CUDACC__) into
// tools/emu/libzrt_emu.so, where every kernel runs on the CPU one thread block at a time, each CUDA thread as a fiber
// (ucontext) and every *_sync warp intrinsic as a rendezvous of the 32 fibers of a warp.  Purpose: debug the warp-level
// CONTROL FLOW of the persistent kernels (item queue, slot pools, rings, the warp schedulers of k_trace_ws / k_trace_pool /
// k_trace_bpool) - deadlocks, lost slots, mismatched collectives - without a GPU, and check the result against the oracle.
// It is ~10^4 times slower than a B200 and is not reachable from the product: zraytrace_b200 loads libzrt.so, which is
// built by nvcc for sm_100a only and has no CPU path.  Arithmetic note: MUFU.RSQ / MUFU.RCP are replaced by correctly rounded
// values; the sequences built on them (unit(), div_exact) produce correctly rounded results either way.
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __grid_constant__
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static const

// ---- vector types ----
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
struct dim3 { unsigned x = 1, y = 1, z = 1; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

// ---- the fiber machine ----
namespace zrt_emu {
enum Tag : int { T_NONE, T_BALLOT, T_ANY, T_SHFL, T_MATCH, T_REDADD, T_SYNC };
struct Warp {
    uint32_t arrived = 0, gen = 0, exited = 0;
    unsigned long long seen[64] = {};
    int tag = T_NONE;
    uint32_t vals[32], aux[32], res[32];
    uint32_t res_scalar = 0;
};
struct Fiber {
    ucontext_t ctx;
    std::unique_ptr<char[]> stack;
    dim3 tid, bid;
    uint32_t lane = 0;
    Warp *warp = nullptr;
    bool done = false;
    uint32_t wait_gen = 0;
    bool waiting = false;
    unsigned long long tick = 0; // ZRT_PROF_TICK: the lane's iteration of the kernel's main loop (loops are warp-uniform)
};
struct Machine {
    ucontext_t main_ctx;
    Fiber *cur = nullptr;
    dim3 bdim, gdim;
    const std::function<void()> *body = nullptr;
    unsigned long long switches = 0, budget = 0;
};
inline Machine &M() { static thread_local Machine m; return m; }
inline void yield() { Machine &m = M(); m.switches++; swapcontext(&m.cur->ctx, &m.main_ctx); }

[[noreturn]] inline void die(const char *what) {
    std::fprintf(stderr, "[zrt_emu] %s (block %u, thread %u)\n", what, M().cur ? M().cur->bid.x : 0u, M().cur ? M().cur->tid.x : 0u);
    std::abort();
}

// every *_sync intrinsic: publish (v, a), wait for the warp's live lanes, the last arriver computes the results
template <class Finish>
inline void rendezvous(int tag, uint32_t v, uint32_t a, Finish finish) {
    Fiber &f = *M().cur;
    Warp &w = *f.warp;
    if (w.arrived == 0) w.tag = tag;
    else if (w.tag != tag) die("lanes of one warp are in different *_sync intrinsics (divergent collective)");
    w.vals[f.lane] = v;
    w.aux[f.lane] = a;
    if (++w.arrived == 32u - w.exited) {
        finish(w);
        w.arrived = 0;
        w.gen++;
    } else {
        f.wait_gen = w.gen;
        f.waiting = true;
        while (w.gen == f.wait_gen) yield();
        f.waiting = false;
    }
}
inline void trampoline() {
    Machine &m = M();
    (*m.body)();
    Fiber &f = *m.cur;
    f.done = true;
    Warp &w = *f.warp;
    w.exited++;
    if (w.arrived && w.arrived == 32u - w.exited) die("a lane left the kernel while the rest of its warp waits in a *_sync intrinsic");
    swapcontext(&f.ctx, &m.main_ctx);
}
// one thread block at a time, warps one after the other (no __syncthreads in these kernels), lanes round-robin
inline void launch(unsigned grid, unsigned block, const std::function<void()> &body) {
    Machine &m = M();
    if (m.cur) die("nested launch");
    const unsigned long long budget = std::getenv("ZRT_EMU_BUDGET") ? std::strtoull(std::getenv("ZRT_EMU_BUDGET"), nullptr, 10) : 4000000000ull;
    m.body = &body;
    m.bdim.x = block;
    m.gdim.x = grid;
    m.switches = 0;
    constexpr size_t STACK = 256 << 10;
    const unsigned warps = (block + 31u) / 32u;
    for (unsigned b = 0; b < grid; b++) {
        std::vector<Warp> ws(warps);
        std::vector<Fiber> fs(block);
        for (unsigned t = 0; t < block; t++) {
            Fiber &f = fs[t];
            f.stack.reset(new char[STACK]);
            f.tid.x = t;
            f.bid.x = b;
            f.lane = t & 31u;
            f.warp = &ws[t >> 5];
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack.get();
            f.ctx.uc_stack.ss_size = STACK;
            f.ctx.uc_link = nullptr;
            makecontext(&f.ctx, (void (*)())trampoline, 0);
        }
        for (unsigned w = 0; w < warps; w++) {
            const unsigned t0 = w * 32u, t1 = std::min(block, t0 + 32u);
            ws[w].exited = 32u - (t1 - t0); // a partial last warp: the missing lanes never existed
            for (;;) {
                bool any = false;
                for (unsigned t = t0; t < t1; t++) {
                    Fiber &f = fs[t];
                    if (f.done || (f.waiting && f.warp->gen == f.wait_gen)) continue;
                    any = true;
                    m.cur = &f;
                    swapcontext(&m.main_ctx, &f.ctx);
                    m.cur = nullptr;
                    if (m.switches > budget) { std::fprintf(stderr, "[zrt_emu] fiber-switch budget exhausted in block %u warp %u: livelock?\n", b, w); std::abort(); }
                }
                bool all_done = true;
                for (unsigned t = t0; t < t1; t++) all_done = all_done && fs[t].done;
                if (all_done) break;
                if (!any) { std::fprintf(stderr, "[zrt_emu] deadlock in block %u warp %u\n", b, w); std::abort(); }
            }
        }
    }
    m.body = nullptr;
}
} // namespace zrt_emu

// ---- section profile: how often a warp ran a code region and with how many lanes (ZRT_PROF in the kernels) ----
namespace zrt_emu {
struct Prof { unsigned long long calls[64], lanes[64]; };
inline Prof g_prof{};
inline void prof(int sec, bool on) {
    if (!on) return;
    Fiber &f = *M().cur;
    Warp &w = *f.warp;
    if (w.seen[sec] != f.tick) { w.seen[sec] = f.tick; g_prof.calls[sec]++; }
    g_prof.lanes[sec]++;
}
inline void prof_tick() { M().cur->tick++; }
} // namespace zrt_emu
#define ZRT_PROF(sec, on) zrt_emu::prof((sec), (on))
#define ZRT_PROF_TICK() zrt_emu::prof_tick()

#define threadIdx (zrt_emu::M().cur->tid)
#define blockIdx (zrt_emu::M().cur->bid)
#define blockDim (zrt_emu::M().bdim)
#define gridDim (zrt_emu::M().gdim)
#define ZRT_LAUNCH(kernel, grid, block, stream, ...) zrt_emu::launch((unsigned)(grid), (unsigned)(block), [=]() { kernel(__VA_ARGS__); })

// ---- warp intrinsics (full masks only: that is all the kernels use) ----
inline unsigned __ballot_sync(unsigned, bool p) {
    zrt_emu::rendezvous(zrt_emu::T_BALLOT, p ? 1u : 0u, 0, [](zrt_emu::Warp &w) {
        uint32_t r = 0;
        for (int i = 0; i < 32; i++) r |= (w.vals[i] & 1u) << i;
        w.res_scalar = r;
    });
    return zrt_emu::M().cur->warp->res_scalar;
}
inline bool __any_sync(unsigned, bool p) {
    zrt_emu::rendezvous(zrt_emu::T_ANY, p ? 1u : 0u, 0, [](zrt_emu::Warp &w) {
        uint32_t r = 0;
        for (int i = 0; i < 32; i++) r |= w.vals[i];
        w.res_scalar = r;
    });
    return zrt_emu::M().cur->warp->res_scalar != 0;
}
inline unsigned __shfl_sync(unsigned, unsigned v, int src) {
    zrt_emu::rendezvous(zrt_emu::T_SHFL, v, (uint32_t)src & 31u, [](zrt_emu::Warp &w) {
        for (int i = 0; i < 32; i++) w.res[i] = w.vals[w.aux[i]];
    });
    return zrt_emu::M().cur->warp->res[zrt_emu::M().cur->lane];
}
inline unsigned __match_any_sync(unsigned, unsigned v) {
    zrt_emu::rendezvous(zrt_emu::T_MATCH, v, 0, [](zrt_emu::Warp &w) {
        for (int i = 0; i < 32; i++) {
            uint32_t r = 0;
            for (int j = 0; j < 32; j++) r |= (w.vals[j] == w.vals[i] ? 1u : 0u) << j;
            w.res[i] = r;
        }
    });
    return zrt_emu::M().cur->warp->res[zrt_emu::M().cur->lane];
}
inline unsigned __reduce_add_sync(unsigned, unsigned v) {
    zrt_emu::rendezvous(zrt_emu::T_REDADD, v, 0, [](zrt_emu::Warp &w) {
        uint32_t r = 0;
        for (int i = 0; i < 32; i++) r += w.vals[i];
        w.res_scalar = r;
    });
    return zrt_emu::M().cur->warp->res_scalar;
}
inline void __syncwarp(unsigned = 0xffffffffu) { zrt_emu::rendezvous(zrt_emu::T_SYNC, 0, 0, [](zrt_emu::Warp &) {}); }
// NOTE: lanes that have exited publish stale vals[]; the kernels never mix exited and live lanes in one collective.

// ---- scalar intrinsics ----
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float2 __fadd2_rn(float2 a, float2 b) { return float2{a.x + b.x, a.y + b.y}; }
inline float2 __fmul2_rn(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)}; }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
inline unsigned __float2uint_rz(float f) { // cvt.rzi.u32.f32: saturates, NaN -> 0
    if (!(f > 0.0f)) return 0u;
    if (f >= 4294967296.0f) return 0xFFFFFFFFu;
    return (unsigned)f;
}
inline int __float2int_rz(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return 0x7FFFFFFF;
    if (f <= -2147483648.0f) return (int)0x80000000;
    return (int)f;
}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline unsigned __brev(unsigned v) {
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
inline int __ffs(int v) { return __builtin_ffs(v); }
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) { // PRMT, default mode
    const unsigned long long xy = ((unsigned long long)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned sel = (s >> (4 * i)) & 0xFu;
        unsigned b = (unsigned)(xy >> (8 * (sel & 7u))) & 0xFFu;
        if (sel & 8u) b = (b & 0x80u) ? 0xFFu : 0x00u;
        r |= b << (8 * i);
    }
    return r;
}
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
inline float rsqrt_approx(float x) { return (float)(1.0 / std::sqrt((double)x)); }
inline float rcp_approx(float x) { return (float)(1.0 / (double)x); }
template <class T> inline T __ldg(const T *p) { return *p; }
template <class T> inline T __ldcg(const T *p) { return *p; }
inline unsigned atomicAdd(unsigned *p, unsigned v) { const unsigned o = *p; *p = o + v; return o; }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
inline unsigned min(unsigned a, unsigned b) { return b < a ? b : a; }
inline unsigned max(unsigned a, unsigned b) { return a < b ? b : a; }
inline int min(int a, int b) { return b < a ? b : a; }
inline int max(int a, int b) { return a < b ? b : a; }
inline unsigned min(unsigned a, int b) { return min(a, (unsigned)b); }
inline unsigned min(int a, unsigned b) { return min((unsigned)a, b); }
inline unsigned max(unsigned a, int b) { return max(a, (unsigned)b); }
inline unsigned max(int a, unsigned b) { return max((unsigned)a, b); }

// ---- the runtime API, as far as libzrt uses it: one emulated device, host memory, synchronous "streams" ----
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2 };
typedef struct zrt_emu_stream *cudaStream_t;
struct zrt_emu_event { std::chrono::steady_clock::time_point t; };
typedef zrt_emu_event *cudaEvent_t;
typedef int cudaMemPool_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocPortable = 1 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum cudaFuncAttribute { cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
enum { cudaSharedmemCarveoutMaxShared = 100 };
enum cudaMemPoolAttr { cudaMemPoolAttrReleaseThreshold = 4 };
struct cudaDeviceProp { int multiProcessorCount; int clockRate; char name[256]; };
constexpr int ZRT_EMU_SMS = 2; // two "SMs", one resident block each: a second block meets an already busy item queue

inline const char *cudaGetErrorString(cudaError_t) { return "zrt_emu"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int) { *v = ZRT_EMU_SMS; return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { p->multiProcessorCount = ZRT_EMU_SMS; p->clockRate = 1000000; std::strcpy(p->name, "zrt_emu"); return cudaSuccess; }
inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, const void *, int, size_t) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaFuncSetAttribute(const void *, cudaFuncAttribute, int) { return cudaSuccess; }
inline cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t *p, int) { *p = 0; return cudaSuccess; }
inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, cudaMemPoolAttr, void *) { return cudaSuccess; }
inline cudaError_t cudaMalloc(void **p, size_t n) { *p = std::aligned_alloc(256, (n + 255) & ~(size_t)255); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc(reinterpret_cast<void **>(p), n); }
inline cudaError_t cudaMallocAsync(void **p, size_t n, cudaStream_t) { return cudaMalloc(p, n); }
template <class T> inline cudaError_t cudaMallocAsync(T **p, size_t n, cudaStream_t s) { return cudaMallocAsync(reinterpret_cast<void **>(p), n, s); }
inline cudaError_t cudaFree(void *p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaFreeAsync(void *p, cudaStream_t) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeHost(void *p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void *d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new zrt_emu_event(); return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess; }
