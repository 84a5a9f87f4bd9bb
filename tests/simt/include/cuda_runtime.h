// cuda_runtime.h -- SIMT emulation shim (TEST INFRASTRUCTURE ONLY, never part of the product).
//
// tests/simt/ compiles the unmodified kernel sources of cdnet_b200/csrc/*.cu with g++ against this
// header so that the -m "not gpu" tier can execute the kernels' logic on the host: the CUDA threads of
// a block are coroutines on one OS thread, blocks are spread over a few OS threads, __syncthreads and
// warp collectives are real rendezvous (tests/simt/simt_runtime.cpp).  It exists to catch logic
// regressions where no GPU is available.  It is NOT a CPU fallback: cdnet_b200/ never loads the
// emulated library (tests/test_cabi_symbols.py asserts that), and nothing measured or shipped goes
// through it.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <functional>
#include <type_traits>

#define CDNET_SIMT 1

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __constant__
#define __shared__ static thread_local

// ---- vector types ---------------------------------------------------------------------------------
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(8) int2 { int x, y; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct alignas(16) longlong2 { long long x, y; };
struct alignas(16) ulonglong2 { unsigned long long x, y; };
struct alignas(4) uchar4 { unsigned char x, y, z, w; };
struct alignas(8) ushort4 { unsigned short x, y, z, w; };
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline longlong2 make_longlong2(long long x, long long y) { return longlong2{x, y}; }
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }
static inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) {
    return uchar4{x, y, z, w};
}
static inline ushort4 make_ushort4(unsigned short x, unsigned short y, unsigned short z, unsigned short w) {
    return ushort4{x, y, z, w};
}

extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;
static const int warpSize = 32;

// ---- host runtime stubs ---------------------------------------------------------------------------
typedef int cudaError_t;
static const cudaError_t cudaSuccess = 0;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { int major, minor, multiProcessorCount; char name[256]; };

static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t = 0) { memset(p, v, n); return 0; }
static inline cudaError_t cudaMemset(void* p, int v, size_t n) { memset(p, v, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) {
    memmove(d, s, n);
    return 0;
}
template <typename T>
static inline cudaError_t cudaMemcpyToSymbolAsync(T& sym, const void* s, size_t n, size_t off, cudaMemcpyKind,
                                                  cudaStream_t = 0) {
    memcpy((char*)&sym + off, s, n);
    return 0;
}
template <typename T>
static inline cudaError_t cudaMemcpyToSymbol(T& sym, const void* s, size_t n, size_t off = 0) {
    memcpy((char*)&sym + off, s, n);
    return 0;
}
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 3; return 0; }  // "3 SMs"
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    memset(p, 0, sizeof *p);
    p->major = 10;
    p->multiProcessorCount = 3;
    strcpy(p->name, "SIMT emulation");
    return 0;
}
template <typename F>
static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }

// ---- kernel launch --------------------------------------------------------------------------------
namespace simt {
void launch(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()>& body, const char* name);
void* dyn_smem();
void sync_block();
int sync_block_or(int pred);
// every lane named in `mask` deposits `v`; returns after all (live) lanes of the mask did; out[32] = their values
void warp_exchange(unsigned mask, unsigned long long v, unsigned long long* out);
int lane_id();
}  // namespace simt

// common.cuh routes every launch through CDNET_LAUNCH; the emulated build substitutes it here
#define CDNET_LAUNCH(kernel, grid, block, smem, stream, ...)                                          \
    do {                                                                                              \
        ::simt::launch(dim3(grid), dim3(block), (size_t)(smem), [&]() { kernel(__VA_ARGS__); }, #kernel); \
        ++::cdnet::g_launches;                                                                        \
    } while (0)
// programmatic dependent launch only changes WHEN a kernel may start; the emulator runs launches back to back anyway
#define CDNET_LAUNCH_PDL CDNET_LAUNCH
#ifdef CDNET_SIMT_ASAN  // tests/simt/build.py --asan: poisoned gaps between the workspace slices (common.cuh Arena)
extern "C" void __asan_poison_memory_region(void const volatile*, size_t);
extern "C" void __asan_unpoison_memory_region(void const volatile*, size_t);
#define CDNET_ARENA_GAP 256
#define CDNET_ARENA_RESET(p, n) do { if (p) __asan_unpoison_memory_region((p), (n)); } while (0)
#define CDNET_ARENA_POISON(p, n) __asan_poison_memory_region((p), (n))
#endif
#define CDNET_DYN_SHARED(type, name) type* name = (type*)::simt::dyn_smem()
#define CDNET_KEEP_IN_REG64(x) ((void)0)

// ---- synchronisation and warp collectives ---------------------------------------------------------
static inline void __syncthreads() { simt::sync_block(); }
static inline int __syncthreads_or(int p) { return simt::sync_block_or(p); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) {
    unsigned long long o[32];
    simt::warp_exchange(mask, 0, o);
}
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

namespace simt {
template <typename T>
static inline unsigned long long to_bits(T v) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    unsigned long long b = 0;
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <typename T>
static inline T from_bits(unsigned long long b) {
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}
}  // namespace simt

template <typename T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    unsigned long long o[32];
    simt::warp_exchange(mask, simt::to_bits(v), o);
    const int lane = simt::lane_id();
    const int base = lane & ~(width - 1);
    return simt::from_bits<T>(o[base + (src & (width - 1))]);
}
template <typename T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    unsigned long long o[32];
    simt::warp_exchange(mask, simt::to_bits(v), o);
    const int lane = simt::lane_id();
    const int base = lane & ~(width - 1);
    const int src = lane - (int)delta;
    return src < base ? v : simt::from_bits<T>(o[src]);
}
template <typename T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    unsigned long long o[32];
    simt::warp_exchange(mask, simt::to_bits(v), o);
    const int lane = simt::lane_id();
    const int base = lane & ~(width - 1);
    const int src = lane + (int)delta;
    return src >= base + width ? v : simt::from_bits<T>(o[src]);
}
template <typename T>
static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
    unsigned long long o[32];
    simt::warp_exchange(mask, simt::to_bits(v), o);
    const int lane = simt::lane_id();
    const int src = lane ^ lanemask;
    (void)width;
    return simt::from_bits<T>(o[src & 31]);
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    unsigned long long o[32];
    simt::warp_exchange(mask, pred ? 1ull : 0ull, o);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i)
        if (((mask >> i) & 1u) && o[i]) r |= 1u << i;
    return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, !pred) == 0; }
static inline unsigned __activemask() { return 0xffffffffu; }
template <typename T>
static inline unsigned __match_any_sync(unsigned mask, T v) {
    unsigned long long o[32];
    const unsigned long long mine = simt::to_bits(v);
    simt::warp_exchange(mask, mine, o);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i)
        if (((mask >> i) & 1u) && o[i] == mine) r |= 1u << i;
    return r;
}
#define SIMT_REDUCE(NAME, T, INIT, OP)                               \
    static inline T NAME(unsigned mask, T v) {                      \
        unsigned long long o[32];                                    \
        simt::warp_exchange(mask, simt::to_bits(v), o);              \
        T r = INIT;                                                  \
        for (int i = 0; i < 32; ++i)                                 \
            if ((mask >> i) & 1u) { T x = simt::from_bits<T>(o[i]); r = OP; } \
        return r;                                                    \
    }
SIMT_REDUCE(__reduce_add_sync, unsigned, 0u, r + x)
SIMT_REDUCE(__reduce_add_sync, int, 0, r + x)
SIMT_REDUCE(__reduce_max_sync, unsigned, 0u, (x > r ? x : r))
SIMT_REDUCE(__reduce_max_sync, int, INT32_MIN, (x > r ? x : r))
SIMT_REDUCE(__reduce_min_sync, unsigned, 0xffffffffu, (x < r ? x : r))
SIMT_REDUCE(__reduce_min_sync, int, INT32_MAX, (x < r ? x : r))
SIMT_REDUCE(__reduce_or_sync, unsigned, 0u, r | x)
SIMT_REDUCE(__reduce_and_sync, unsigned, 0xffffffffu, r & x)
#undef SIMT_REDUCE

// ---- atomics (relaxed is enough: kernels only rely on atomicity, ordering comes from barriers) ----
template <typename T>
static inline T simt_atomic_rmw(T* p, T v, T (*op)(T, T)) {
    T old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (!__atomic_compare_exchange_n(p, &old, op(old, v), true, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED)) {}
    return old;
}
#define SIMT_INT_ATOMICS(T)                                                                            \
    static inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }        \
    static inline T atomicSub(T* p, T v) { return __atomic_fetch_sub(p, v, __ATOMIC_SEQ_CST); }        \
    static inline T atomicOr(T* p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }          \
    static inline T atomicAnd(T* p, T v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }        \
    static inline T atomicXor(T* p, T v) { return __atomic_fetch_xor(p, v, __ATOMIC_SEQ_CST); }        \
    static inline T atomicExch(T* p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }      \
    static inline T atomicCAS(T* p, T cmp, T v) {                                                      \
        __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);            \
        return cmp;                                                                                    \
    }                                                                                                  \
    static inline T atomicMin(T* p, T v) { return simt_atomic_rmw<T>(p, v, [](T a, T b) { return a < b ? a : b; }); } \
    static inline T atomicMax(T* p, T v) { return simt_atomic_rmw<T>(p, v, [](T a, T b) { return a > b ? a : b; }); }
SIMT_INT_ATOMICS(int)
SIMT_INT_ATOMICS(unsigned)
SIMT_INT_ATOMICS(long long)
SIMT_INT_ATOMICS(unsigned long long)
#undef SIMT_INT_ATOMICS
static inline unsigned short atomicCAS(unsigned short* p, unsigned short cmp, unsigned short v) {
    __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}
static inline float atomicAdd(float* p, float v) {
    unsigned* u = (unsigned*)p;
    unsigned old = __atomic_load_n(u, __ATOMIC_RELAXED), nw;
    float f;
    do {
        memcpy(&f, &old, 4);
        f += v;
        memcpy(&nw, &f, 4);
    } while (!__atomic_compare_exchange_n(u, &old, nw, true, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED));
    memcpy(&f, &old, 4);
    return f;
}
static inline double atomicAdd(double* p, double v) {
    unsigned long long* u = (unsigned long long*)p;
    unsigned long long old = __atomic_load_n(u, __ATOMIC_RELAXED), nw;
    double f;
    do {
        memcpy(&f, &old, 8);
        f += v;
        memcpy(&nw, &f, 8);
    } while (!__atomic_compare_exchange_n(u, &old, nw, true, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED));
    memcpy(&f, &old, 8);
    return f;
}

// ---- loads ------------------------------------------------------------------------------------------
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcg(const T* p) { return *(const volatile T*)p; }
template <typename T> static inline T __ldcs(const T* p) { return *p; }
template <typename T> static inline void __stcg(T* p, T v) { *p = v; }
template <typename T> static inline void __stcs(T* p, T v) { *p = v; }

// ---- integer intrinsics -----------------------------------------------------------------------------
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline int __clzll(long long v) { return v == 0 ? 64 : __builtin_clzll((unsigned long long)v); }
static inline unsigned __brev(unsigned v) {
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
    const unsigned long long src = ((unsigned long long)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) {
        const unsigned sel = (s >> (4 * i)) & 0xfu;
        unsigned byte = (unsigned)(src >> (8 * (sel & 7u))) & 0xffu;
        if (sel & 8u) byte = (byte & 0x80u) ? 0xffu : 0x00u;  // msb replication mode
        r |= byte << (8 * i);
    }
    return r;
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) {
    const unsigned long long v = ((unsigned long long)hi << 32) | lo;
    return (unsigned)((v << (sh & 31u)) >> 32);
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
    const unsigned long long v = ((unsigned long long)hi << 32) | lo;
    return (unsigned)(v >> (sh & 31u));
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline int __mulhi(int a, int b) { return (int)(((long long)a * b) >> 32); }

// ---- floating point intrinsics (the emulated build uses -ffp-contract=off: IEEE add/mul like -fmad=false)
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return sqrt(a); }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __double2float_rn(double a) { return (float)a; }
static inline int __double2int_rn(double a) { return (int)nearbyint(a); }
static inline int __float2int_rn(float a) { return (int)nearbyintf(a); }
static inline int __double2loint(double a) { return (int)(simt::to_bits(a) & 0xffffffffull); }
static inline int __double2hiint(double a) { return (int)(simt::to_bits(a) >> 32); }
static inline long long __double_as_longlong(double a) { return simt::from_bits<long long>(simt::to_bits(a)); }
static inline double __longlong_as_double(long long a) { return simt::from_bits<double>((unsigned long long)a); }
static inline unsigned __float_as_uint(float a) { return (unsigned)simt::to_bits(a); }
static inline int __float_as_int(float a) { return (int)simt::to_bits(a); }
static inline float __uint_as_float(unsigned a) { return simt::from_bits<float>(a); }
static inline float __int_as_float(int a) { return simt::from_bits<float>((unsigned)a); }
static inline double __int2double_rn(int a) { return (double)a; }
static inline float __int2float_rn(int a) { return (float)a; }

// CUDA's global min/max overloads
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline unsigned min(int a, unsigned b) { return min((unsigned)a, b); }
static inline unsigned min(unsigned a, int b) { return min(a, (unsigned)b); }
static inline unsigned max(int a, unsigned b) { return max((unsigned)a, b); }
static inline unsigned max(unsigned a, int b) { return max(a, (unsigned)b); }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
static inline size_t min(size_t a, size_t b) { return a < b ? a : b; }
static inline size_t max(size_t a, size_t b) { return a > b ? a : b; }
static inline float min(float a, float b) { return fminf(a, b); }
static inline float max(float a, float b) { return fmaxf(a, b); }
static inline double min(double a, double b) { return fmin(a, b); }
static inline double max(double a, double b) { return fmax(a, b); }
