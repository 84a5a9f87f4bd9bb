// common.cuh -- shared device/host helpers of libcdnet_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/cdnet_b200.h"
#if defined(__CUDACC__) && !defined(CDNET_NO_NVTX)
#include <nvtx3/nvToolsExt.h>
#endif

namespace cdnet {

extern unsigned long long g_launches;  // api.cu
extern int g_prof_on;                  // api.cu: per-launch CUDA-event timing (cdnet_profile_*)
void prof_begin(const char* name, cudaStream_t st);
void prof_end(cudaStream_t st);

// CDNET_LAUNCH / CDNET_DYN_SHARED / CDNET_KEEP_IN_REG64 are the only three places where the sources use
// syntax a host compiler cannot parse; tests/simt/ (a SIMT emulator used by the CPU test tier to execute the
// kernels' logic, never by the product) pre-defines them.
#ifndef CDNET_LAUNCH
#define CDNET_LAUNCH(kernel, grid, block, smem, stream, ...)                      \
    do {                                                                          \
        if (::cdnet::g_prof_on) ::cdnet::prof_begin(#kernel, (stream));           \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);               \
        if (::cdnet::g_prof_on) ::cdnet::prof_end((stream));                      \
        ++::cdnet::g_launches;                                                    \
    } while (0)
// The same, as a programmatic dependent launch (PDL): the grid may be scheduled while the previous kernel of the stream
// drains, and its blocks wait in pdl_wait() until that kernel has completed and its writes are visible.  Saves the
// launch gap and the tail of every boundary of a chain of short kernels.  Every kernel launched this way calls
// pdl_wait() before its first global access (and before any early return); the kernel in front of it calls
// pdl_trigger().  Off under the per-launch profiler and with CDNET_NO_PDL=1.
#define CDNET_LAUNCH_PDL(kernel, grid, block, smem, stream, ...)                                        \
    do {                                                                                                \
        if (::cdnet::g_prof_on) ::cdnet::prof_begin(#kernel, (stream));                                 \
        ::cdnet::launch_pdl(kernel, dim3(grid), dim3(block), (size_t)(smem), (stream), __VA_ARGS__);    \
        if (::cdnet::g_prof_on) ::cdnet::prof_end((stream));                                            \
        ++::cdnet::g_launches;                                                                          \
    } while (0)
// dynamic shared memory of the running block, viewed as `type name[]`
#define CDNET_DYN_SHARED(type, name) extern __shared__ type name[]
// pins a 64-bit value in one register pair (stops ptxas from rematerialising it per use)
#define CDNET_KEEP_IN_REG64(x) asm volatile("" : "+l"(x))
#endif

// NVTX ranges around the pipeline phases (visible in nsys / ncu timelines; header-only NVTX v3, a no-op unless a
// profiler injects its library).  CDNET_RANGE("name") covers the rest of the enclosing scope.
#if defined(__CUDACC__) && !defined(CDNET_NO_NVTX)
struct NvtxScope {
    explicit NvtxScope(const char* name) { nvtxRangePushA(name); }
    ~NvtxScope() { nvtxRangePop(); }
};
#define CDNET_RANGE_CAT2(a, b) a##b
#define CDNET_RANGE_CAT(a, b) CDNET_RANGE_CAT2(a, b)
#define CDNET_RANGE(name) ::cdnet::NvtxScope CDNET_RANGE_CAT(nvtx_scope_, __LINE__)(name)
static inline void nvtx_mark(const char* name) { nvtxMarkA(name); }
#else
#define CDNET_RANGE(name) ((void)0)
static inline void nvtx_mark(const char*) {}
#endif

#if defined(__CUDACC__)
bool pdl_enabled();  // api.cu
template <typename... KArgs, typename... Args>
static inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl_enabled() && !g_prof_on) ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);  // errors surface through last_error()
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#else
static inline void pdl_wait() {}
static inline void pdl_trigger() {}
#endif

#define CDNET_CUDA_OK(expr)                                 \
    do {                                                    \
        cudaError_t e__ = (expr);                           \
        if (e__ != cudaSuccess) return -(int)e__;           \
    } while (0)

static inline int last_error() {
    cudaError_t e = cudaGetLastError();  // returns AND clears: an error is attributed to the call that caused it
    return e == cudaSuccess ? 0 : -(int)e;
}

// AddressSanitizer builds of the emulated test library (tests/simt/build.py --asan) put a poisoned gap after
// every workspace slice, so a kernel that runs over the end of its slice is caught; no-ops everywhere else.
#ifndef CDNET_ARENA_GAP
#define CDNET_ARENA_GAP 0
#define CDNET_ARENA_RESET(p, n) ((void)0)
#define CDNET_ARENA_POISON(p, n) ((void)0)
#endif

// bump allocator over the caller's workspace (256-byte aligned slices)
struct Arena {
    char* base;
    size_t size, off;
    bool ok;
    Arena(void* p, size_t n) : base((char*)p), size(n), off(0), ok(true) { CDNET_ARENA_RESET(p, n); }
    template <typename T>
    T* take(size_t count) {
        size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
        if (base == nullptr || off + bytes > size) { ok = false; off += bytes; return nullptr; }
        T* r = (T*)(base + off);
        const size_t gap = (CDNET_ARENA_GAP && off + bytes + CDNET_ARENA_GAP <= size) ? CDNET_ARENA_GAP : 0;
        CDNET_ARENA_POISON(base + off + count * sizeof(T), bytes - count * sizeof(T) + gap);
        off += bytes + gap;
        return r;
    }
};
static inline size_t pad256(size_t n) { return (n + 255) & ~size_t(255); }

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- union-find on an int32 parent array; links always point to a smaller index, so the root
// ---- of a finished component is its first pixel in raster order -------------------------------
// Plain (L1-cached) loads: a stale parent is still a node of the same tree whose chain ends at the
// current root, and uf_union re-validates the root with the value atomicMin returns, so coherence is
// not needed here -- and most hops hit L1 instead of paying an L2 round trip.
__device__ __forceinline__ int uf_find(const int* L, int p) {
    int q = L[p];
    while (q != p) {
        p = q;
        q = L[p];
    }
    return p;
}

__device__ __forceinline__ void uf_union(int* L, int a, int b) {
    const int a0 = a, b0 = b;
    for (;;) {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a == b) break;
        if (a < b) { int t = a; a = b; b = t; }
        int old = atomicMin(L + a, b);  // a > b: hang a under b
        if (old == a) { a = b; break; }
        a = old;
    }
    // keep the trees shallow: both starting points now point (almost) at the common root.  Results unused ->
    // fire-and-forget reductions, nothing waits on them.
    const int r = a < b ? a : b;
    if (a0 != r) atomicMin(L + a0, r);
    if (b0 != r) atomicMin(L + b0, r);
}

// bitwise select (m ? x : y) in ONE LOP3 -- ptxas does not always fuse the two masked halves itself
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t bitsel(uint32_t m, uint32_t x, uint32_t y) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(r) : "r"(m), "r"(x), "r"(y));
    return r;
}
#else
static inline uint32_t bitsel(uint32_t m, uint32_t x, uint32_t y) { return (x & m) | (y & ~m); }
#endif

// find with path halving: every visited node is re-pointed at its grandparent with a PLAIN store.  Parent pointers
// only ever point at smaller indices of the same tree, so a stale store can at worst undo a little compression --
// it cannot break the forest -- and there is no atomic traffic on the hot upper levels of a very large tree (the
// background of a tile).
__device__ __forceinline__ int uf_find_c(int* L, int p) {
    int q = L[p];
    while (q != p) {
        const int g = L[q];
        if (g != q) L[p] = g;
        p = q;
        q = g;
    }
    return p;
}

// both finds of a union at once: the two chains of dependent loads overlap instead of following each other
__device__ __forceinline__ void uf_find2_c(int* L, int& a, int& b) {
    int qa = L[a], qb = L[b];
    while (qa != a || qb != b) {
        const int ga = L[qa], gb = L[qb];
        if (qa != a) {
            if (ga != qa) L[a] = ga;
            a = qa;
            qa = ga;
        }
        if (qb != b) {
            if (gb != qb) L[b] = gb;
            b = qb;
            qb = gb;
        }
    }
}

__device__ __forceinline__ void uf_union_c(int* L, int a, int b) {
    for (;;) {
        uf_find2_c(L, a, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }
        const int old = atomicMin(L + a, b);  // a > b: hang a under b
        if (old == a) return;
        a = old;
    }
}

// float <-> order-preserving uint32 (for atomicMax on floats of any sign)
__device__ __forceinline__ unsigned int f32_to_ordered(float f) {
    unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_f32(unsigned int u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

}  // namespace cdnet
