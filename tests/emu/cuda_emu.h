// TEST INFRASTRUCTURE ONLY — a minimal CPU emulation of the CUDA subset used by yaik_b200/csrc/*.cu so the
// kernels' LOGIC (indexing, stream order, ownership rules) can be exercised in this GPU-less container
// before spending GPU time.  It is compiled only into tests/emu/_build/libyaik_b200_emu.so by
// tests/emu/Makefile; the product library (nvcc, sm_100a) never sees this header and has no CPU path.
//
// Model: every launch runs its CTAs one after another; the threads of a CTA are real OS threads that meet
// at std::barrier objects for __syncthreads() and for warp collectives (which therefore require all 32
// lanes of a warp to participate — the kernels are written that way).  `__shared__` becomes a function
// static (one CTA is resident at a time).
#pragma once
#include <atomic>
#include <barrier>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>
#include <algorithm>

#define YK_EMULATE 1
#include <math.h>
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __shared__ static
#define __restrict__
#define __launch_bounds__(...)
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct float4 { float x, y, z, w; };
struct float2 { float x, y; };
static inline float2 make_float2(float a, float b) { return float2{ a, b }; }
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct uchar4 { unsigned char x, y, z, w; };
static inline int4 make_int4(int a, int b, int c, int d) { return int4{ a, b, c, d }; }
static inline int2 make_int2(int a, int b) { return int2{ a, b }; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{ a, b, c, d }; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{ a, b }; }

namespace yk_emu {
struct Warp { std::unique_ptr<std::barrier<>> bar; unsigned long long slot[32]; };
struct Cta {
    std::unique_ptr<std::barrier<>> bar;
    std::vector<Warp> warps;
    std::vector<unsigned char> dynSmem;
};
extern Cta* g_cta;
extern dim3 g_blockDim, g_gridDim;
extern thread_local uint3 t_threadIdx, t_blockIdx;
extern thread_local int t_lane, t_warp;
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
inline Warp& warp() { return g_cta->warps[t_warp]; }
template <class F> inline unsigned long long collect(unsigned long long mine, F reduce) {
    Warp& w = warp();
    w.slot[t_lane] = mine;
    w.bar->arrive_and_wait();
    unsigned long long r = reduce(w.slot);
    w.bar->arrive_and_wait();
    return r;
}
}  // namespace yk_emu

#define threadIdx (yk_emu::t_threadIdx)
#define blockIdx (yk_emu::t_blockIdx)
#define blockDim (yk_emu::g_blockDim)
#define gridDim (yk_emu::g_gridDim)
#define warpSize 32

static inline void __syncthreads() { yk_emu::g_cta->bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { yk_emu::warp().bar->arrive_and_wait(); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }

static inline void yk_emu_fullmask(unsigned m) { if (m != 0xffffffffu) { fprintf(stderr, "yk_emu: partial warp mask %08x\n", m); abort(); } }
static inline unsigned __ballot_sync(unsigned m, int p) {
    yk_emu_fullmask(m);
    return (unsigned)yk_emu::collect(p ? 1 : 0, [](unsigned long long* s) { unsigned r = 0; for (int i = 0; i < 32; i++) if (s[i]) r |= 1u << i; return (unsigned long long)r; });
}
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
static inline int __all_sync(unsigned m, int p) { return __ballot_sync(m, p) == 0xffffffffu; }
template <class T> static inline T __shfl_sync(unsigned m, T v, int src, int width = 32) {
    yk_emu_fullmask(m);
    unsigned long long bits = 0; memcpy(&bits, &v, sizeof(T));
    int lane = yk_emu::t_lane; int base = lane & ~(width - 1);
    int from = base + (src & (width - 1));
    unsigned long long r = yk_emu::collect(bits, [from](unsigned long long* s) { return s[from]; });
    T out; memcpy(&out, &r, sizeof(T)); return out;
}
template <class T> static inline T __shfl_xor_sync(unsigned m, T v, int x, int width = 32) { return __shfl_sync(m, v, (yk_emu::t_lane ^ x) & (width - 1), width); }
template <class T> static inline T __shfl_down_sync(unsigned m, T v, unsigned d, int width = 32) {
    int l = yk_emu::t_lane & (width - 1); int src = (l + (int)d < width) ? l + (int)d : l; return __shfl_sync(m, v, src, width);
}
template <class T> static inline T __shfl_up_sync(unsigned m, T v, unsigned d, int width = 32) {
    int l = yk_emu::t_lane & (width - 1); int src = (l - (int)d >= 0) ? l - (int)d : l; return __shfl_sync(m, v, src, width);
}
// reductions take a member mask: every lane of the warp reaches the call (the kernels only split a warp into its two halves,
// each half naming itself), and a lane's result covers the lanes of its own mask
template <class F> static inline long long yk_emu_reduce(unsigned m, long long v, F op) {
    return (long long)yk_emu::collect((unsigned long long)v, [m, op](unsigned long long* s) {
        bool first = true; long long r = 0;
        for (int i = 0; i < 32; i++) if ((m >> i) & 1u) { r = first ? (long long)s[i] : op(r, (long long)s[i]); first = false; }
        return (unsigned long long)r; });
}
static inline int __reduce_min_sync(unsigned m, int v) { return (int)yk_emu_reduce(m, v, [](long long a, long long b) { return std::min(a, b); }); }
static inline int __reduce_max_sync(unsigned m, int v) { return (int)yk_emu_reduce(m, v, [](long long a, long long b) { return std::max(a, b); }); }
static inline unsigned __reduce_max_sync(unsigned m, unsigned v) { return (unsigned)yk_emu_reduce(m, (long long)v, [](long long a, long long b) { return std::max(a, b); }); }
static inline unsigned __reduce_min_sync(unsigned m, unsigned v) { return (unsigned)yk_emu_reduce(m, (long long)v, [](long long a, long long b) { return std::min(a, b); }); }
static inline int __reduce_add_sync(unsigned m, int v) { return (int)yk_emu_reduce(m, v, [](long long a, long long b) { return a + b; }); }
static inline unsigned __reduce_add_sync(unsigned m, unsigned v) { return (unsigned)yk_emu_reduce(m, (long long)v, [](long long a, long long b) { return a + b; }); }
static inline unsigned __reduce_or_sync(unsigned m, unsigned v) { return (unsigned)yk_emu_reduce(m, (long long)v, [](long long a, long long b) { return a | b; }); }

static inline unsigned __match_any_sync(unsigned m, int v) {
    yk_emu_fullmask(m);
    return (unsigned)yk_emu::collect((unsigned long long)(unsigned)v, [v](unsigned long long* s) { unsigned r = 0; for (int i = 0; i < 32; i++) if ((unsigned)s[i] == (unsigned)v) r |= 1u << i; return (unsigned long long)r; });
}
static inline void __nanosleep(unsigned) { std::this_thread::yield(); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; i++) if (v & (1u << i)) r |= 1u << (31 - i); return r; }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
    unsigned long long v = ((unsigned long long)b << 32) | a; unsigned r = 0;
    for (int i = 0; i < 4; i++) { unsigned sel = (s >> (4 * i)) & 7; r |= (unsigned)((v >> (8 * sel)) & 255) << (8 * i); }
    return r;
}
static inline int __vimax3_s32(int a, int b, int c) { return std::max(a, std::max(b, c)); }
static inline int __vimin3_s32(int a, int b, int c) { return std::min(a, std::min(b, c)); }
static inline unsigned __vmaxu2(unsigned a, unsigned b) { return std::max(a & 0xFFFFu, b & 0xFFFFu) | (std::max(a >> 16, b >> 16) << 16); }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __int2float_rn(int a) { return (float)a; }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) { sh &= 31u; return (unsigned)(((((unsigned long long)hi << 32) | lo) << sh) >> 32); }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) { sh &= 31u; return (unsigned)((((unsigned long long)hi << 32) | lo) >> sh); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
using std::max; using std::min;
template <class T> static inline T __ldg(const T* p) { return *p; }

template <class T> static inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicSub(T* p, T v) { return __atomic_fetch_sub(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicOr(T* p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicAnd(T* p, T v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicExch(T* p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicMin(T* p, T v) { T o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (v < o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
template <class T> static inline T atomicMax(T* p, T v) { T o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (v > o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
template <class T> static inline T atomicCAS(T* p, T c, T v) { __atomic_compare_exchange_n(p, &c, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return c; }

// ---- runtime API subset ------------------------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef struct yk_emu_event { double t; }* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) & ~(size_t)255); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = (void*)1; return 0; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (void*)1; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new yk_emu_event{0}; return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
#define cudaEventDisableTiming 2
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new yk_emu_event{0}; return 0; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0; return 0; }
#define cudaStreamNonBlocking 1
#define cudaFuncAttributeMaxDynamicSharedMemorySize 8
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }

#define YK_LAUNCH(kernel, grid, block, smem, stream, ...) \
    yk_emu::launch((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); })
