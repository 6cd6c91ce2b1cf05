// tests/emu/syg_emu.h -- TEST INFRASTRUCTURE ONLY.
//
// A small CUDA-on-CPU emulator: one ucontext fiber per CUDA thread, blocks distributed over a few OS
// threads.  It exists so that the kernels in sygnals_b200/csrc (compiled a second time with g++ -DSYG_EMU)
// can be executed in the GPU-less build container: index arithmetic, barrier placement, warp collectives,
// atomics and all the host-side plan logic are exercised before any GPU time is spent.  It is NOT a product
// path: sygnals_b200 only ever loads libsygb200.so (nvcc, sm_100a) and raises if that is missing.
//
// Fiber scheduling is run-to-barrier in ascending (or, with SYG_EMU_REVERSE=1, descending) thread order, so a
// missing __syncthreads()/__syncwarp() shows up as a wrong result in at least one of the two orders.
// Optional shared-memory bank-conflict accounting: SYG_EMU_BANKS=1 (accesses made through sygdev::sld/sst).
#pragma once

#include <ucontext.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <thread>
#include <tuple>
#include <vector>

// ------------------------------------------------------------------ CUDA vocabulary
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __align__(n) alignas(n)
#define __constant__ static

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }

typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
#define cudaStreamLegacy ((cudaStream_t)0x1)
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2,
                      cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrMaxSharedMemoryPerBlockOptin = 97 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0, cudaHostRegisterDefault = 0 };

namespace sygemu {

constexpr int kWarp = 32;
constexpr size_t kStack = 256 * 1024;

struct Fiber {
    ucontext_t ctx;
    char* stack = nullptr;
    bool done = false;
    std::map<int, int> occ;  // bank accounting: per-site occurrence counter
};

struct WarpState {
    uint64_t slot[kWarp];
    int count = 0;
    unsigned gen = 0;
};

struct BankRec { int warp, site, occ, lane; uintptr_t addr; int bytes; };

struct BlockCtx {
    std::vector<Fiber> fibers;
    std::vector<WarpState> warps;
    ucontext_t sched;
    int nthreads = 0, alive = 0;
    int bar_count = 0;
    unsigned bar_gen = 0;
    int cur = 0;
    std::function<void()>* body = nullptr;
    unsigned char* dyn = nullptr;
    std::vector<BankRec> bank;
};

extern thread_local BlockCtx* g_blk;
extern bool g_banks;
extern std::mutex g_bank_mu;
extern std::map<int, std::pair<long, long>> g_bank_stats;  // site -> (warp-instructions, wavefronts)

void launch_impl(dim3 grid, dim3 block, size_t smem, std::function<void()> body);
void yield();
void bank_report(FILE* f);
void bank_reset();

template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem, F&& f) {
    launch_impl(grid, block, smem, std::function<void()>(f));
}
inline void launch(int grid, int block, size_t smem, std::function<void()> f) {
    launch_impl(dim3(grid), dim3(block), smem, std::move(f));
}
inline unsigned char* dyn_smem() { return g_blk->dyn; }

inline int lane_id();
void warp_barrier(unsigned mask);
void block_barrier();
void smem_access(const void* p, int bytes, int site);

}  // namespace sygemu

extern thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
constexpr int warpSize = 32;

inline int sygemu::lane_id() { return (int)(threadIdx.x % 32); }

// ------------------------------------------------------------------ synchronisation + warp collectives
static inline void __syncthreads() { sygemu::block_barrier(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { sygemu::warp_barrier(mask); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

namespace sygemu {
template <class T>
inline T shfl_generic(unsigned mask, T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shfl payload");
    WarpState& w = g_blk->warps[threadIdx.x / 32];
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    w.slot[lane_id()] = raw;
    warp_barrier(mask);
    uint64_t got = w.slot[src_lane & 31];
    warp_barrier(mask);
    T out;
    std::memcpy(&out, &got, sizeof(T));
    return out;
}
}  // namespace sygemu

template <class T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    int lane = sygemu::lane_id();
    int base = lane & ~(width - 1);
    return sygemu::shfl_generic(mask, v, base + (src & (width - 1)));
}
template <class T>
static inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask, int width = 32) {
    int lane = sygemu::lane_id();
    int tgt = lane ^ lane_mask;
    if ((tgt & ~(width - 1)) != (lane & ~(width - 1))) tgt = lane;
    return sygemu::shfl_generic(mask, v, tgt);
}
template <class T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    int lane = sygemu::lane_id();
    int tgt = lane + (int)delta;
    if ((tgt & ~(width - 1)) != (lane & ~(width - 1))) tgt = lane;
    return sygemu::shfl_generic(mask, v, tgt);
}
template <class T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    int lane = sygemu::lane_id();
    int tgt = lane - (int)delta;
    if (tgt < (lane & ~(width - 1))) tgt = lane;
    return sygemu::shfl_generic(mask, v, tgt);
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    sygemu::WarpState& w = sygemu::g_blk->warps[threadIdx.x / 32];
    w.slot[sygemu::lane_id()] = pred ? 1 : 0;
    sygemu::warp_barrier(mask);
    unsigned r = 0;
    int nth = sygemu::g_blk->nthreads, wbase = (threadIdx.x / 32) * 32;
    for (int l = 0; l < 32; ++l)
        if (((mask >> l) & 1u) && wbase + l < nth && w.slot[l]) r |= 1u << l;
    sygemu::warp_barrier(mask);
    return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, !pred) == 0; }
static inline unsigned __activemask() { return 0xffffffffu; }
namespace sygemu {
template <class T, class Op>
inline T warp_reduce_generic(unsigned mask, T v, Op op) {
    WarpState& w = g_blk->warps[threadIdx.x / 32];
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    w.slot[lane_id()] = raw;
    warp_barrier(mask);
    bool first = true;
    T acc{};
    int nth = g_blk->nthreads, wbase = (threadIdx.x / 32) * 32;
    for (int l = 0; l < 32; ++l) {
        if (!((mask >> l) & 1u) || wbase + l >= nth) continue;
        T x;
        std::memcpy(&x, &w.slot[l], sizeof(T));
        acc = first ? x : op(acc, x);
        first = false;
    }
    warp_barrier(mask);
    return acc;
}
}  // namespace sygemu
static inline int __reduce_add_sync(unsigned m, int v) { return sygemu::warp_reduce_generic(m, v, [](int a, int b) { return a + b; }); }
static inline unsigned __reduce_add_sync(unsigned m, unsigned v) { return sygemu::warp_reduce_generic(m, v, [](unsigned a, unsigned b) { return a + b; }); }
static inline int __reduce_max_sync(unsigned m, int v) { return sygemu::warp_reduce_generic(m, v, [](int a, int b) { return a > b ? a : b; }); }
static inline unsigned __reduce_max_sync(unsigned m, unsigned v) { return sygemu::warp_reduce_generic(m, v, [](unsigned a, unsigned b) { return a > b ? a : b; }); }
static inline int __reduce_min_sync(unsigned m, int v) { return sygemu::warp_reduce_generic(m, v, [](int a, int b) { return a < b ? a : b; }); }
static inline unsigned __reduce_min_sync(unsigned m, unsigned v) { return sygemu::warp_reduce_generic(m, v, [](unsigned a, unsigned b) { return a < b ? a : b; }); }
static inline unsigned __reduce_or_sync(unsigned m, unsigned v) { return sygemu::warp_reduce_generic(m, v, [](unsigned a, unsigned b) { return a | b; }); }

// ------------------------------------------------------------------ atomics (blocks may run on several OS threads)
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline float atomicAdd(float* p, float v) {
    uint32_t* ip = reinterpret_cast<uint32_t*>(p);
    uint32_t old = __atomic_load_n(ip, __ATOMIC_RELAXED), nw;
    float f;
    do {
        std::memcpy(&f, &old, 4);
        f += v;
        std::memcpy(&nw, &f, 4);
    } while (!__atomic_compare_exchange_n(ip, &old, nw, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    std::memcpy(&f, &old, 4);
    return f;
}
static inline unsigned atomicMax(unsigned* p, unsigned v) {
    unsigned old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
static inline int atomicMax(int* p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
static inline unsigned atomicMin(unsigned* p, unsigned v) {
    unsigned old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}

// ------------------------------------------------------------------ device math / intrinsics
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float rsqrtf(float a) { return 1.0f / std::sqrt(a); }
static inline float __fsqrt_rn(float a) { return std::sqrt(a); }
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
static inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline double __dsqrt_rn(double a) { return std::sqrt(a); }
template <class T> static inline T __ldg(const T* p) { return *p; }
using std::max;
using std::min;
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)}; }
static inline float2 __fadd2_rn(float2 a, float2 b) { return float2{a.x + b.x, a.y + b.y}; }
static inline float2 __fmul2_rn(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }

// ------------------------------------------------------------------ runtime API subset
static inline cudaError_t cudaMalloc(void** p, size_t n) {
    *p = n ? std::aligned_alloc(256, (n + 255) / 256 * 256) : nullptr;
    return (*p || !n) ? cudaSuccess : cudaErrorMemoryAllocation;
}
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), n); }
static inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaHostRegister(void*, size_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaHostUnregister(void*) { return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { if (n) std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { if (n) std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulator"; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int) {
    *v = (a == cudaDevAttrMultiProcessorCount) ? 3 : (a == cudaDevAttrMaxSharedMemoryPerBlockOptin ? 227 * 1024 : 0);
    return cudaSuccess;
}
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
