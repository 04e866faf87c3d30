// tests/emu/cuda_emu.h -- TEST INFRASTRUCTURE ONLY.  Never shipped, never loaded by the
// daliti_b200 package.
//
// A small single-threaded emulator of the CUDA execution model, just large enough to run
// the library's kernels (daliti_b200/csrc/*.cu compiled as C++ with -DDLT_EMU) on a CPU so
// that kernel *logic* (hash insert, ring search, warp top-k selection, block reductions)
// can be debugged in a container without a GPU before a gpurun call is spent.  Every CUDA
// thread is a stackful coroutine; warp collectives and __syncthreads are rendezvous points.
// Blocks run one after another, so `__shared__` maps to a function-local static.
//
// It is not a CPU fallback of the product: the product library is the nvcc build and its
// loader refuses to run without a CUDA device.  Results from this emulator are used only
// by `-m "not gpu"` tests to check kernel logic against the oracle.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __shared__ static
#define __restrict__ __restrict
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)

struct uint3 {
    unsigned x, y, z;
};
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct __attribute__((aligned(16))) float4 {
    float x, y, z, w;
};
struct __attribute__((aligned(8))) float2 {
    float x, y;
};
struct __attribute__((aligned(16))) double2 {
    double x, y;
};
struct __attribute__((aligned(16))) int4 {
    int x, y, z, w;
};
struct __attribute__((aligned(16))) uint4 {
    unsigned x, y, z, w;
};
struct __attribute__((aligned(16))) ulonglong2 {
    unsigned long long x, y;
};
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }

typedef void *cudaStream_t;

namespace emu {
extern uint3 g_threadIdx, g_blockIdx;
extern dim3 g_blockDim, g_gridDim;
void launch(dim3 grid, dim3 block, const std::function<void()> &body);
void block_sync();
// generic warp collective: every live lane contributes (payload, aux); returns after all arrived.
// kind: 0 shfl idx, 1 shfl xor, 2 shfl down, 3 shfl up, 4 ballot, 5 syncwarp
unsigned long long warp_collective(int kind, unsigned mask, unsigned long long payload, int aux, int width);
}  // namespace emu

#define threadIdx emu::g_threadIdx
#define blockIdx emu::g_blockIdx
#define blockDim emu::g_blockDim
#define gridDim emu::g_gridDim
#define warpSize 32

static inline void __syncthreads() { emu::block_sync(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::warp_collective(5, mask, 0, 0, 32); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __threadfence_system() { __sync_synchronize(); }  // (peer mailboxes are shared memory between emulator processes)

namespace emu {
template <typename T>
static inline unsigned long long to_bits(T v) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    unsigned long long b = 0;
    std::memcpy(&b, &v, sizeof(T));
    return b;
}
template <typename T>
static inline T from_bits(unsigned long long b) {
    T v;
    std::memcpy(&v, &b, sizeof(T));
    return v;
}
}  // namespace emu

template <typename T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    return emu::from_bits<T>(emu::warp_collective(0, mask, emu::to_bits(v), src, width));
}
template <typename T>
static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
    return emu::from_bits<T>(emu::warp_collective(1, mask, emu::to_bits(v), lanemask, width));
}
template <typename T>
static inline T __shfl_down_sync(unsigned mask, T v, int delta, int width = 32) {
    return emu::from_bits<T>(emu::warp_collective(2, mask, emu::to_bits(v), delta, width));
}
template <typename T>
static inline T __shfl_up_sync(unsigned mask, T v, int delta, int width = 32) {
    return emu::from_bits<T>(emu::warp_collective(3, mask, emu::to_bits(v), delta, width));
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    return (unsigned)emu::warp_collective(4, mask, pred ? 1ull : 0ull, 0, 32);
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return emu::warp_collective(4, mask, pred ? 0ull : 1ull, 0, 32) == 0; }

static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline unsigned __float_as_uint(float f) { return emu::from_bits<unsigned>(emu::to_bits(f)); }
static inline int __float_as_int(float f) { return emu::from_bits<int>(emu::to_bits(f)); }
static inline float __uint_as_float(unsigned u) { return emu::from_bits<float>(emu::to_bits(u)); }
static inline float __int_as_float(int u) { return emu::from_bits<float>(emu::to_bits(u)); }
static inline long long __double_as_longlong(double d) { return emu::from_bits<long long>(emu::to_bits(d)); }
static inline double __longlong_as_double(long long l) { return emu::from_bits<double>(emu::to_bits(l)); }
static inline long long __double2ll_rn(double d) { return llrint(d); }
static inline double __ll2double_rn(long long v) { return (double)v; }
template <typename T>
static inline T __ldg(const T *p) {
    return *p;
}
template <typename T>
static inline T __ldcg(const T *p) {
    return *p;
}

// atomics: the emulator is single-threaded, so these are plain read-modify-writes
template <typename T>
static inline T atomicAdd(T *a, T v) {
    T o = *a;
    *a = o + v;
    return o;
}
template <typename T>
static inline T atomicSub(T *a, T v) {
    T o = *a;
    *a = o - v;
    return o;
}
template <typename T>
static inline T atomicExch(T *a, T v) {
    T o = *a;
    *a = v;
    return o;
}
template <typename T>
static inline T atomicCAS(T *a, T cmp, T v) {
    T o = *a;
    if (o == cmp) *a = v;
    return o;
}
template <typename T>
static inline T atomicMin(T *a, T v) {
    T o = *a;
    if (v < o) *a = v;
    return o;
}
template <typename T>
static inline T atomicMax(T *a, T v) {
    T o = *a;
    if (v > o) *a = v;
    return o;
}
template <typename T>
static inline T atomicOr(T *a, T v) {
    T o = *a;
    *a = o | v;
    return o;
}
template <typename T>
static inline T atomicAnd(T *a, T v) {
    T o = *a;
    *a = o & v;
    return o;
}

using std::max;
using std::min;

#include <atomic>
namespace dlt { namespace rt { extern std::atomic<unsigned long long> g_launches; } }
#define DLT_LAUNCH(kernel, grid, block, stream, ...)                              \
    do {                                                                          \
        emu::launch(dim3(grid), dim3(block), [=]() { kernel(__VA_ARGS__); });     \
        ++dlt::rt::g_launches;                                                    \
    } while (0)
