// ccl.cu -- union-find connected components with canonical raster-order numbering (K6-K8).
//
// Replaces scipy.ndimage.label (4-conn; postproc_other.py:37,44), skimage.measure.label (8-conn;
// test_dam.py:561, test.py:292, my_transforms_direction.py:755,773), scipy binary_fill_holes
// (test_dam.py:546, test.py:277, postproc_other.py:42,51) and skimage remove_small_objects
// (test_dam.py:548, test.py:279, postproc_other.py:46,48,53).
//
// Design: one int32 parent plane L per tile.  (1) init: every pixel points at the first pixel of
// its horizontal run inside its 32-pixel warp segment (ballot + clz, no memory traffic besides the
// store); (2) merge: only the links that are not implied by a neighbour are united with an
// atomicMin union-find (first column of each vertical overlap, warp-segment seams, non-redundant
// diagonals); links always point to the smaller index so the root is the component's first raster
// pixel; (3) flatten.  Ids = rank of the root among roots in raster order: per-row root counts
// (ballot/popc), an exclusive scan over rows, and a per-row running rank.
//
// fill holes / remove small / 8-connected labelling of the inference post-processing run on ONE
// forest: equal-value 4-connected components of the mask (foreground and background alike) ->
// background components that do not touch the frame are holes -> holes are united with their
// foreground neighbours -> areas -> small components dropped -> surviving components that touch
// diagonally are united -> numbering.
#include <stdlib.h>

#include "internal.h"

namespace cdnet {

constexpr int kBX = 128, kBY = 4;
#define CCL_COORDS                                             \
    const int x = blockIdx.x * kBX + threadIdx.x;              \
    const int y = blockIdx.y * kBY + threadIdx.y;              \
    const int b = blockIdx.z;                                  \
    const bool inb = (x < W) && (y < H);                       \
    const size_t tile = (size_t)b * H * W;                     \
    const int p = y * W + x;                                   \
    const int lane = threadIdx.x & 31;                         \
    (void)lane; (void)p; (void)tile; (void)inb;

static inline dim3 ccl_grid(int B, int H, int W) { return dim3(ceil_div(W, kBX), ceil_div(H, kBY), B); }
static inline dim3 ccl_block() { return dim3(kBX, kBY); }

// EQ = false: foreground-only components (mask != 0).  EQ = true: equal-value components of a 0/1 mask.
// One block covers whole rows (row_warps warps per row, up to 1024 pixels per row chunk): every pixel
// points at the first pixel of its horizontal run inside the chunk -- link bits by ballot, the run start
// by clz on the warp word or, for runs spanning warps, a prefix-max over the chunk's 32 words.
__global__ void __launch_bounds__(1024) k_ccl_init_rows(const uint8_t* __restrict__ mask, int* __restrict__ L,
                                                        int* __restrict__ zero1, int* __restrict__ zero2, int H, int W,
                                                        int row_warps, int rows_per_block, int eq) {
    __shared__ unsigned s_link[32];
    __shared__ int s_lastzero[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int b = blockIdx.z;
    const size_t tile = (size_t)b * H * W;
    const int r_in_block = wid / row_warps, seg = wid % row_warps;
    const int y = blockIdx.y * rows_per_block + r_in_block;
    const int x = blockIdx.x * (row_warps * 32) + seg * 32 + lane;
    const bool active = r_in_block < rows_per_block;
    const bool inb = active && y < H && x < W;
    const int p = y * W + x;
    int v = -1;
    if (inb) v = mask[tile + p] != 0;
    // link to the left neighbour (same row chunk only; chunk seams are united by the merge kernel)
    int vl = __shfl_up_sync(0xffffffffu, v, 1);
    if (lane == 0) vl = (inb && seg > 0) ? (mask[tile + p - 1] != 0) : -2;
    const bool link = inb && (eq ? (v == vl) : (v == 1 && vl == 1));
    const unsigned m = __ballot_sync(0xffffffffu, link);
    s_link[wid] = m;
    __syncthreads();
    // highest run-start position (a zero link bit) at or before the end of each warp word, per row
    if (wid == 0) {
        const int my_row = lane / row_warps, my_seg = lane % row_warps;
        const unsigned z = ~s_link[lane];
        int lz = (31 - __clz(z)) + my_seg * 32;  // z != 0 for segment 0 (bit 0 is never linked)
        if (z == 0) lz = -1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, lz, d);
            const int orow = __shfl_up_sync(0xffffffffu, my_row, d);
            if (lane >= d && orow == my_row) lz = max(lz, o);
        }
        s_lastzero[lane] = lz;
    }
    __syncthreads();
    if (!inb) return;
    const unsigned zmine = ~m & (0xffffffffu >> (31 - lane));
    int start_in_chunk;
    if (zmine) start_in_chunk = seg * 32 + (31 - __clz(zmine));
    else start_in_chunk = s_lastzero[wid - 1];
    const int start = p - (seg * 32 + lane) + start_in_chunk;
    L[tile + p] = start;
    if (zero1) zero1[tile + p] = 0;
    if (zero2) zero2[tile + p] = 0;
}

struct InitGeom {
    dim3 grid, block;
    int row_warps, rows_per_block;
};
static inline InitGeom init_geom(int B, int H, int W) {
    InitGeom g;
    const int rw = ceil_div(W, 32) < 32 ? ceil_div(W, 32) : 32;
    g.row_warps = rw;
    g.rows_per_block = 32 / rw;
    g.block = dim3(1024);
    g.grid = dim3(ceil_div(W, rw * 32), ceil_div(H, g.rows_per_block), B);
    return g;
}
#define CCL_INIT(EQ, st, mask, L, z1, z2)                                                                       \
    do {                                                                                                        \
        InitGeom g__ = init_geom(B, H, W);                                                                      \
        CDNET_LAUNCH(k_ccl_init_rows, g__.grid, g__.block, 0, st, mask, L, z1, z2, H, W, g__.row_warps,         \
                     g__.rows_per_block, (EQ) ? 1 : 0);                                                         \
    } while (0)

// Vertical (and, for CONN 8, diagonal) links between row y and row y-1 for the rows y = (2k+1) << level.
// Rows are merged in binary-tree order (level 0 joins row pairs, level 1 joins pairs of pairs, ...): every
// level hangs the roots of the lower strip under roots of the upper strip, so the depth of the forest is
// bounded by the number of levels (<= log2 H) instead of growing with H when all rows link at once.
template <bool EQ, int CONN>
__global__ void __launch_bounds__(kBX* kBY) k_ccl_merge(const uint8_t* __restrict__ mask, int* __restrict__ L, int H, int W,
                                                        int level, int stride) {
    const int x = blockIdx.x * kBX + threadIdx.x;
    const int ri = blockIdx.y * kBY + threadIdx.y;
    const int y = stride > 0 ? (ri + 1) * stride : ((2 * ri + 1) << level);
    const int b = blockIdx.z;
    if (x >= W || y >= H) return;
    const size_t tile = (size_t)b * H * W;
    const int p = y * W + x;
    const uint8_t* M = mask + tile;
    int* Lt = L + tile;
    const int v = M[p] != 0;
    if (!EQ && !v) return;
    auto same = [&](int q) -> bool { const int u = M[q] != 0; return EQ ? (u == v) : (u != 0); };
    const bool hl = x > 0 && same(p - 1);
    const bool up = same(p - W);
    if (up) {
        // implied by the left neighbour's own vertical link when all four pixels agree
        const bool redundant = hl && same(p - W - 1);
        if (!redundant) uf_union(Lt, p, p - W);
    } else if (CONN == 8) {
        if (x > 0 && !hl && same(p - W - 1)) uf_union(Lt, p, p - W - 1);
        if (x + 1 < W && same(p - W + 1) && !same(p + 1)) uf_union(Lt, p, p - W + 1);
    }
}

// 4 pixels per thread (W % 4 == 0): two 32-bit mask loads + two edge bytes, the link tests on 4-bit vectors;
// only the (rare) pixels that really need a union touch the parent plane.
__device__ __forceinline__ uint32_t nz_bits4(uint32_t w) {  // bit i = byte i of w is non-zero
    const uint32_t n = (((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) & 0x80808080u;
    return ((n >> 7) & 1u) | ((n >> 14) & 2u) | ((n >> 21) & 4u) | ((n >> 28) & 8u);
}

template <bool EQ, int CONN>
__global__ void __launch_bounds__(256) k_ccl_merge4(const uint8_t* __restrict__ mask, int* __restrict__ L, int H, int W,
                                                    int level, int stride) {
    const int x4 = (blockIdx.x * 64 + threadIdx.x) * 4;
    const int ri = blockIdx.y * 4 + threadIdx.y;
    const int y = stride > 0 ? (ri + 1) * stride : ((2 * ri + 1) << level);
    const int b = blockIdx.z;
    if (x4 >= W || y >= H) return;
    const size_t tile = (size_t)b * H * W;
    const int p = y * W + x4;
    const uint8_t* M = mask + tile;
    const uint32_t V = nz_bits4(*(const uint32_t*)(M + p));
    if (!EQ && V == 0) return;
    const uint32_t U = nz_bits4(*(const uint32_t*)(M + p - W));
    uint32_t lv = 0, lu = 0, okl = 0xEu;
    if (x4 > 0) { lv = M[p - 1] != 0; lu = M[p - W - 1] != 0; okl = 0xFu; }
    const uint32_t VL = ((V << 1) | lv) & 0xFu, UL = ((U << 1) | lu) & 0xFu;
    uint32_t same_up, hl, same_ul;
    if (EQ) { same_up = ~(V ^ U); hl = ~(V ^ VL) & okl; same_ul = ~(V ^ UL) & okl; }
    else { same_up = V & U; hl = V & VL & okl; same_ul = V & UL & okl; }
    uint32_t need = same_up & ~(hl & same_ul) & 0xFu;
    int* Lt = L + tile;
    while (need) {
        const int i = __ffs(need) - 1;
        need &= need - 1;
        uf_union(Lt, p + i, p + i - W);
    }
    if (CONN == 8) {
        uint32_t rv = 0, ru = 0, okr = 0x7u;
        if (x4 + 4 < W) { rv = M[p + 4] != 0; ru = M[p - W + 4] != 0; okr = 0xFu; }
        const uint32_t VR = (V >> 1) | (rv << 3), UR = (U >> 1) | (ru << 3);
        uint32_t dl = V & ~U & ~VL & UL & okl;
        uint32_t dr = V & ~U & UR & ~VR & okr;
        while (dl) {
            const int i = __ffs(dl) - 1;
            dl &= dl - 1;
            uf_union(Lt, p + i, p + i - W - 1);
        }
        while (dr) {
            const int i = __ffs(dr) - 1;
            dr &= dr - 1;
            uf_union(Lt, p + i, p + i - W + 1);
        }
    }
}

// runs that continue across a 1024-pixel row-chunk seam of the init kernel (only when W > 1024)
template <bool EQ>
__global__ void k_ccl_merge_seams(const uint8_t* __restrict__ mask, int* __restrict__ L, int H, int W) {
    const int nseam = (W - 1) / 1024;
    const int b = blockIdx.y;
    const size_t tile = (size_t)b * H * W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nseam * H; i += gridDim.x * blockDim.x) {
        const int y = i / nseam, x = (i % nseam + 1) * 1024;
        const int p = y * W + x;
        const int v = mask[tile + p] != 0, u = mask[tile + p - 1] != 0;
        if (EQ ? (u == v) : (u && v)) uf_union(L + tile, p, p - 1);
    }
}

template <bool EQ, int CONN>
static void merge_all(const uint8_t* mask, int* L, int B, int H, int W, cudaStream_t st) {
    if (W > 1024) {
        const int n = ((W - 1) / 1024) * H;
        CDNET_LAUNCH(k_ccl_merge_seams<EQ>, dim3(ceil_div(n, 256), B), 256, 0, st, mask, L, H, W);
    }
    for (int level = 0; (1 << level) < H; ++level) {
        const int nrows = (H - (1 << level) + (2 << level) - 1) / (2 << level);  // rows y = (2k+1)<<level < H
        if (nrows <= 0) break;
        if (W % 4 == 0 && ((uintptr_t)mask & 3) == 0)
            CDNET_LAUNCH((k_ccl_merge4<EQ, CONN>), dim3(ceil_div(W, 256), ceil_div(nrows, 4), B), dim3(64, 4), 0, st, mask,
                         L, H, W, level, 0);
        else
            CDNET_LAUNCH((k_ccl_merge<EQ, CONN>), dim3(ceil_div(W, kBX), ceil_div(nrows, kBY), B), ccl_block(), 0, st, mask,
                         L, H, W, level, 0);
    }
}

// ---- strip-local labelling in shared memory -------------------------------------------------------------
// A block owns SR consecutive rows of one tile (SR * W <= 32768 pixels): mask bytes and an int32 parent per
// pixel live in shared memory, where a union-find hop costs ~30 cycles instead of an L2 round trip.  Runs
// are initialised per 32-pixel segment (ballot + clz), every link inside the strip is united in shared
// memory, the strip is flattened and written out with tile-global indices.  What is left for global memory
// are the links across strip seams: ONE sparse launch (chains are at most H / SR links long).
__device__ __forceinline__ int sfind(const int* P, int p) {
    int q = P[p];
    while (q != p) { p = q; q = P[p]; }
    return p;
}
__device__ __forceinline__ void sunion(int* P, int a, int b) {
    for (;;) {
        a = sfind(P, a);
        b = sfind(P, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }
        const int old = atomicMin(P + a, b);
        if (old == a) return;
        a = old;
    }
}

template <bool EQ, int CONN>
__global__ void __launch_bounds__(1024) k_ccl_strip(const uint8_t* __restrict__ mask, int* __restrict__ L,
                                                    int* __restrict__ zero1, int* __restrict__ zero2, int H, int W, int SR) {
    CDNET_DYN_SHARED(int, s_par);
    uint8_t* s_m = (uint8_t*)(s_par + SR * W);
    const int tid = threadIdx.x, lane = tid & 31, nt = blockDim.x;
    const int y0 = blockIdx.x * SR;
    const int rows = min(SR, H - y0);
    const int n = rows * W;
    const size_t base = (size_t)blockIdx.y * H * W + (size_t)y0 * W;
    const uint8_t* M = mask + base;
    const bool vec = (W % 4 == 0) && (((uintptr_t)M & 3) == 0) && (((uintptr_t)(L + base) & 15) == 0) &&
                     (!zero1 || ((uintptr_t)(zero1 + base) & 15) == 0) && (!zero2 || ((uintptr_t)(zero2 + base) & 15) == 0);
    if (vec) {
        const int4 z4 = make_int4(0, 0, 0, 0);
        for (int i = tid * 4; i < n; i += nt * 4) {
            const uint32_t w = *(const uint32_t*)(M + i);
            const uint32_t nz = ((((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) & 0x80808080u) >> 7;  // 0/1 per byte
            *(uint32_t*)(s_m + i) = nz;
            if (zero1) *(int4*)(zero1 + base + i) = z4;
            if (zero2) *(int4*)(zero2 + base + i) = z4;
        }
    } else {
        for (int i = tid; i < n; i += nt) {
            s_m[i] = M[i] != 0;
            if (zero1) zero1[base + i] = 0;
            if (zero2) zero2[base + i] = 0;
        }
    }
    __syncthreads();
    auto same = [&](int v, int u) -> bool { return EQ ? (v == u) : (v && u); };
    // rows are walked explicitly (no per-pixel division); a warp always covers x0 .. x0+31 of one row, so the
    // 32-pixel segments start at multiples of 32 in every row
    const int wbase = tid - lane;
    for (int r = 0; r < rows; ++r) {
        const int rb = r * W;
        for (int x0 = wbase; x0 < W; x0 += nt) {
            const int x = x0 + lane;
            const bool valid = x < W;
            const int v = valid ? s_m[rb + x] : 0;
            const bool link = valid && lane > 0 && same(v, s_m[rb + x - 1]);
            const unsigned m = __ballot_sync(0xffffffffu, link);
            if (valid) s_par[rb + x] = rb + x - __clz(~(m << (31 - lane)));
        }
    }
    __syncthreads();
    for (int r = 0; r < rows; ++r) {
        const int rb = r * W;
        for (int x = tid; x < W; x += nt) {
            const int i = rb + x;
            const int v = s_m[i];
            if (!EQ && !v) continue;
            const bool hl = x > 0 && same(v, s_m[i - 1]);
            if (hl && lane == 0) sunion(s_par, i, i - 1);
            if (r > 0) {
                if (same(v, s_m[i - W])) {
                    if (!(hl && same(v, s_m[i - W - 1]))) sunion(s_par, i, i - W);
                } else if (CONN == 8) {
                    if (x > 0 && !hl && same(v, s_m[i - W - 1])) sunion(s_par, i, i - W - 1);
                    if (x + 1 < W && same(v, s_m[i - W + 1]) && !same(v, s_m[i + 1])) sunion(s_par, i, i - W + 1);
                }
            }
        }
    }
    __syncthreads();
    const int goff = y0 * W;
    if (vec) {
        for (int i = tid * 4; i < n; i += nt * 4) {
            int r0 = sfind(s_par, i);
            int r1 = s_par[i + 1] == s_par[i] ? r0 : sfind(s_par, i + 1);
            int r2 = s_par[i + 2] == s_par[i + 1] ? r1 : sfind(s_par, i + 2);
            int r3 = s_par[i + 3] == s_par[i + 2] ? r2 : sfind(s_par, i + 3);
            *(int4*)(L + base + i) = make_int4(goff + r0, goff + r1, goff + r2, goff + r3);
        }
    } else {
        for (int i = tid; i < n; i += nt) L[base + i] = goff + sfind(s_par, i);
    }
}

static int strip_rows(int H, int W) {
    if (W > 16384) return 0;  // a strip needs at least 2 rows in 160 KB
    int sr = 32768 / W;
    static int cap = 0;
    if (!cap) { const char* e = getenv("CDNET_STRIP_ROWS"); cap = e ? atoi(e) : 4; if (cap < 2) cap = 4; }
    if (sr > cap) sr = cap;  // measured best on B200 for 1000-wide tiles: 4 rows x 512 threads (20 KB, many blocks per SM)
    return sr >= 2 ? sr : 0;
}

template <bool EQ, int CONN>
static int forest_build(const uint8_t* mask, int* L, int* zero1, int* zero2, int B, int H, int W, cudaStream_t st);

__global__ void __launch_bounds__(kBX* kBY) k_flatten(int* __restrict__ L, int H, int W) {
    CCL_COORDS
    if (!inb) return;
    int* Lt = L + tile;
    Lt[p] = uf_find(Lt, p);
}

template <bool EQ, int CONN>
static int forest_build(const uint8_t* mask, int* L, int* zero1, int* zero2, int B, int H, int W, cudaStream_t st) {
    const int SR = strip_rows(H, W);
    static int disable = -1;
    if (disable < 0) disable = getenv("CDNET_NO_STRIP") ? 1 : 0;
    if (SR == 0 || disable) {
        CCL_INIT(EQ, st, mask, L, zero1, zero2);
        merge_all<EQ, CONN>(mask, L, B, H, W, st);
        return last_error();
    }
    const size_t smem = (size_t)SR * W * 5;
    static bool attr_done[2][2] = {{false, false}, {false, false}};
    if (!attr_done[EQ ? 1 : 0][CONN == 8 ? 1 : 0]) {
        CDNET_CUDA_OK(cudaFuncSetAttribute(k_ccl_strip<EQ, CONN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 163840));
        attr_done[EQ ? 1 : 0][CONN == 8 ? 1 : 0] = true;
    }
    static int nthreads = 0;
    if (!nthreads) { const char* e = getenv("CDNET_STRIP_THREADS"); nthreads = e ? atoi(e) : 512; }
    CDNET_LAUNCH((k_ccl_strip<EQ, CONN>), dim3(ceil_div(H, SR), B), nthreads, smem, st, mask, L, zero1, zero2, H, W, SR);
    const int nseam = (H - 1) / SR;  // rows SR, 2 SR, ... < H
    if (nseam > 0) {
        if (W % 4 == 0 && ((uintptr_t)mask & 3) == 0)
            CDNET_LAUNCH((k_ccl_merge4<EQ, CONN>), dim3(ceil_div(W, 256), ceil_div(nseam, 4), B), dim3(64, 4), 0, st, mask,
                         L, H, W, 0, SR);
        else
            CDNET_LAUNCH((k_ccl_merge<EQ, CONN>), dim3(ceil_div(W, kBX), ceil_div(nseam, kBY), B), ccl_block(), 0, st, mask,
                         L, H, W, 0, SR);
    }
    return last_error();
}

// =====================================================================================================
// 4-pixels-per-thread variants of the per-pixel passes (W % 4 == 0): 128-bit parent loads/stores, 32-bit
// mask loads, four independent root chases in flight per thread.  Same results as the scalar kernels.
// =====================================================================================================
#define V4_COORDS                                                   \
    const int x4 = (blockIdx.x * 64 + threadIdx.x) * 4;             \
    const int y = blockIdx.y * 4 + threadIdx.y;                     \
    const int b = blockIdx.z;                                       \
    const bool inb = (x4 < W) && (y < H);                           \
    const size_t tile = (size_t)b * H * W;                          \
    const int p = y * W + x4;                                       \
    (void)p; (void)tile; (void)inb;
static inline dim3 v4_grid(int B, int H, int W) { return dim3(ceil_div(W, 256), ceil_div(H, 4), B); }
static inline dim3 v4_block() { return dim3(64, 4); }
static inline bool v4_ok(int W, const void* a, const void* b2 = nullptr, const void* c = nullptr) {
    return W % 4 == 0 && ((uintptr_t)a & 15) == 0 && ((uintptr_t)b2 & 15) == 0 && ((uintptr_t)c & 15) == 0;
}

__device__ __forceinline__ void find4(const int* __restrict__ Lt, int p, int4 l, int r[4]) {
    const int q[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (i > 0 && q[i] == q[i - 1]) { r[i] = r[i - 1]; continue; }  // same run: same root
        r[i] = (q[i] == p + i) ? q[i] : uf_find(Lt, q[i]);
    }
}

__global__ void __launch_bounds__(256) k_flatten_fill4(const uint8_t* __restrict__ mask, int* __restrict__ L,
                                                       const int* __restrict__ touch, uint8_t* __restrict__ state, int H,
                                                       int W) {
    V4_COORDS
    if (!inb) return;
    int* Lt = L + tile;
    int r[4];
    find4(Lt, p, *(const int4*)(Lt + p), r);
    *(int4*)(Lt + p) = make_int4(r[0], r[1], r[2], r[3]);
    const uint32_t m = *(const uint32_t*)(mask + tile + p);
    uint32_t st = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const bool fg = (m >> (8 * i)) & 0xffu;
        uint32_t s;
        if (fg) s = 1;
        else if (i > 0 && r[i] == r[i - 1] && !((m >> (8 * (i - 1))) & 0xffu)) s = (st >> (8 * (i - 1))) & 0xffu;
        else s = touch[tile + r[i]] ? 0u : 2u;
        st |= s << (8 * i);
    }
    *(uint32_t*)(state + tile + p) = st;
}

// Hole pixels (state 2) join the foreground components around them.  The forest was flattened by the previous
// pass, so L holds roots (or, under concurrent unions, nodes one hop from them): all neighbour states and parents
// of the word's hole pixels are loaded first (independent loads in flight together), duplicates are dropped (a
// hole is nearly always enclosed by ONE component) and the unions start at roots -- a short dependent chain
// instead of four full find/union walks per hole pixel.
__global__ void __launch_bounds__(256) k_fill_merge4(const uint8_t* __restrict__ state, int* __restrict__ L, int H, int W) {
    V4_COORDS
    if (!inb) return;
    const uint8_t* S = state + tile;
    const uint32_t w = *(const uint32_t*)(S + p);
    // any byte == 2 ?
    const uint32_t t = w ^ 0x02020202u;
    if (!((t - 0x01010101u) & ~t & 0x80808080u)) return;
    int* Lt = L + tile;
    int own[4], cand[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        own[i] = -1;
#pragma unroll
        for (int k = 0; k < 4; ++k) cand[i][k] = -1;
        if (((w >> (8 * i)) & 0xffu) != 2u) continue;
        const int q = p + i, x = x4 + i;
        own[i] = Lt[q];
        if (y > 0 && S[q - W] == 1) cand[i][0] = Lt[q - W];
        if (x > 0 && S[q - 1] == 1) cand[i][1] = Lt[q - 1];
        if (x + 1 < W && S[q + 1] == 1) cand[i][2] = Lt[q + 1];
        if (y + 1 < H && S[q + W] == 1) cand[i][3] = Lt[q + W];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (own[i] < 0) continue;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = cand[i][k];
            if (c < 0) continue;
            bool dup = false;
#pragma unroll
            for (int j = 0; j < k; ++j) dup |= (cand[i][j] == c);
            if (i > 0 && own[i - 1] == own[i]) {  // the previous pixel of the same hole already joined this one
#pragma unroll
                for (int j = 0; j < 4; ++j) dup |= (cand[i - 1][j] == c);
            }
            if (!dup) uf_union(Lt, own[i], c);
        }
    }
}

__global__ void __launch_bounds__(256) k_flatten_area4(const uint8_t* __restrict__ state, int* __restrict__ L,
                                                       int* __restrict__ area, int H, int W, int row_lo, int row_hi) {
    V4_COORDS
    if (!inb) return;
    const uint32_t w = *(const uint32_t*)(state + tile + p);
    if (w == 0) return;
    int* Lt = L + tile;
    const int4 l = *(const int4*)(Lt + p);
    const int q[4] = {l.x, l.y, l.z, l.w};
    int r[4];
    bool dirty = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (!((w >> (8 * i)) & 0xffu)) { r[i] = -1; continue; }
        if (i > 0 && r[i - 1] >= 0 && q[i] == q[i - 1]) r[i] = r[i - 1];
        else r[i] = (q[i] == p + i) ? q[i] : uf_find(Lt, q[i]);
        dirty |= (r[i] != q[i]);
    }
    if (dirty) *(int4*)(Lt + p) = make_int4(r[0] >= 0 ? r[0] : q[0], r[1] >= 0 ? r[1] : q[1], r[2] >= 0 ? r[2] : q[2],
                                            r[3] >= 0 ? r[3] : q[3]);
    if (y < row_lo || y >= row_hi) return;
    int run = 0, cur = -1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (r[i] == cur) { run += (cur >= 0); continue; }
        if (cur >= 0) atomicAdd(area + tile + cur, run);
        cur = r[i];
        run = 1;
    }
    if (cur >= 0) atomicAdd(area + tile + cur, run);
}

__global__ void __launch_bounds__(256) k_keep_large4(const uint8_t* __restrict__ state, const int* __restrict__ L,
                                                     const int* __restrict__ area, uint8_t* __restrict__ keep, int min_area,
                                                     int H, int W) {
    V4_COORDS
    if (!inb) return;
    const uint32_t w = *(const uint32_t*)(state + tile + p);
    uint32_t k = 0;
    if (w) {
        const int4 l = *(const int4*)(L + tile + p);
        const int q[4] = {l.x, l.y, l.z, l.w};
        int prev_root = -1;
        uint32_t prev_keep = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (!((w >> (8 * i)) & 0xffu)) continue;
            uint32_t ki;
            if (q[i] == prev_root) ki = prev_keep;
            else ki = area[tile + q[i]] >= min_area ? 1u : 0u;
            prev_root = q[i];
            prev_keep = ki;
            k |= ki << (8 * i);
        }
    }
    *(uint32_t*)(keep + tile + p) = k;
}

__global__ void __launch_bounds__(256) k_diag_merge4(const uint8_t* __restrict__ keep, int* __restrict__ L, int H, int W) {
    V4_COORDS
    if (!inb || y == 0) return;
    const uint8_t* K = keep + tile;
    const uint32_t V = nz_bits4(*(const uint32_t*)(K + p));
    if (!V) return;
    const uint32_t U = nz_bits4(*(const uint32_t*)(K + p - W));
    uint32_t lv = 0, lu = 0, okl = 0xEu, rv = 0, ru = 0, okr = 0x7u;
    if (x4 > 0) { lv = K[p - 1] != 0; lu = K[p - W - 1] != 0; okl = 0xFu; }
    if (x4 + 4 < W) { rv = K[p + 4] != 0; ru = K[p - W + 4] != 0; okr = 0xFu; }
    const uint32_t VL = ((V << 1) | lv) & 0xFu, UL = ((U << 1) | lu) & 0xFu;
    const uint32_t VR = (V >> 1) | (rv << 3), UR = (U >> 1) | (ru << 3);
    uint32_t dl = V & ~U & UL & ~VL & okl;
    uint32_t dr = V & ~U & UR & ~VR & okr;
    int* Lt = L + tile;
    while (dl) {
        const int i = __ffs(dl) - 1;
        dl &= dl - 1;
        uf_union(Lt, p + i, p + i - W - 1);
    }
    while (dr) {
        const int i = __ffs(dr) - 1;
        dr &= dr - 1;
        uf_union(Lt, p + i, p + i - W + 1);
    }
}

__global__ void __launch_bounds__(256) k_flatten_count4(int* __restrict__ L, const uint8_t* __restrict__ keep,
                                                        int* __restrict__ rowcnt, int H, int W,
                                                        const uint8_t* __restrict__ excluded) {
    V4_COORDS
    int cnt = 0;
    if (inb) {
        const uint32_t w = *(const uint32_t*)(keep + tile + p);
        if (w) {
            int* Lt = L + tile;
            const int4 l = *(const int4*)(Lt + p);
            const int q[4] = {l.x, l.y, l.z, l.w};
            int r[4];
            bool dirty = false;
            const uint32_t ex = excluded ? *(const uint32_t*)(excluded + tile + p) : 0u;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                r[i] = q[i];
                if (!((w >> (8 * i)) & 0xffu)) continue;
                if (i > 0 && ((w >> (8 * (i - 1))) & 0xffu) && q[i] == q[i - 1]) r[i] = r[i - 1];
                else r[i] = (q[i] == p + i) ? q[i] : uf_find(Lt, q[i]);
                dirty |= (r[i] != q[i]);
                cnt += (r[i] == p + i) && !((ex >> (8 * i)) & 0xffu);
            }
            if (dirty) *(int4*)(Lt + p) = make_int4(r[0], r[1], r[2], r[3]);
        }
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt && y < H) atomicAdd(rowcnt + (size_t)b * H + y, cnt);
}

__global__ void __launch_bounds__(256) k_relabel4(const int* __restrict__ L, const uint8_t* __restrict__ keep,
                                                  const int* __restrict__ idmap, int* __restrict__ out, int H, int W) {
    V4_COORDS
    if (!inb) return;
    const uint32_t w = *(const uint32_t*)(keep + tile + p);
    int o[4] = {0, 0, 0, 0};
    if (w) {
        const int4 l = *(const int4*)(L + tile + p);
        const int q[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (!((w >> (8 * i)) & 0xffu)) continue;
            if (i > 0 && ((w >> (8 * (i - 1))) & 0xffu) && q[i] == q[i - 1]) o[i] = o[i - 1];
            else o[i] = idmap[tile + q[i]];
        }
    }
    *(int4*)(out + tile + p) = make_int4(o[0], o[1], o[2], o[3]);
}

// ---- numbering ------------------------------------------------------------------------------------
// flatten + count roots of kept pixels per row.  keep == nullptr: every pixel with L-root semantics
// handled by the caller's mask `fg`.
__global__ void __launch_bounds__(kBX* kBY) k_flatten_count(int* __restrict__ L, const uint8_t* __restrict__ keep,
                                                            int* __restrict__ rowcnt, int H, int W,
                                                            const uint8_t* __restrict__ excluded) {
    CCL_COORDS
    bool root = false;
    if (inb && keep[tile + p]) {
        int* Lt = L + tile;
        const int r = uf_find(Lt, p);
        Lt[p] = r;
        root = (r == p) && !(excluded && excluded[tile + p]);
    }
    const unsigned m = __ballot_sync(0xffffffffu, root);
    if (lane == 0 && m && y < H) atomicAdd(rowcnt + (size_t)b * H + y, __popc(m));
}

// exclusive scan of rowcnt over the H rows of each tile (in place); n_out[b] = total
__global__ void __launch_bounds__(1024) k_scan_rows(int* __restrict__ rowcnt, int* __restrict__ n_out, int H) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int b = blockIdx.x;
    int* rc = rowcnt + (size_t)b * H;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < H; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < H ? rc[i] : 0;
        int s = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += t;
        }
        if (lane == 31) s_warp[wid] = s;
        __syncthreads();
        if (wid == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += t;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const int carry = s_carry;
        const int incl = s + (wid > 0 ? s_warp[wid - 1] : 0) + carry;
        if (i < H) rc[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0 && n_out) n_out[b] = s_carry;
}

// one warp per row: idmap[root pixel] = 1 + number of roots before it in raster order
__global__ void __launch_bounds__(256) k_assign_ids(const int* __restrict__ L, const uint8_t* __restrict__ keep,
                                                    const int* __restrict__ rowbase, int* __restrict__ idmap, int H, int W,
                                                    const uint8_t* __restrict__ excluded) {
    const int lane = threadIdx.x & 31;
    const int y = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int b = blockIdx.y;
    if (y >= H) return;
    const size_t tile = (size_t)b * H * W;
    int running = rowbase[(size_t)b * H + y] + 1;
    for (int x0 = 0; x0 < W; x0 += 32) {
        const int x = x0 + lane;
        const int p = y * W + x;
        const bool root = x < W && keep[tile + p] && L[tile + p] == p && !(excluded && excluded[tile + p]);
        const unsigned m = __ballot_sync(0xffffffffu, root);
        if (root) idmap[tile + p] = running + __popc(m & ((1u << lane) - 1));
        running += __popc(m);
    }
}

__global__ void __launch_bounds__(kBX* kBY) k_relabel(const int* __restrict__ L, const uint8_t* __restrict__ keep,
                                                      const int* __restrict__ idmap, int* __restrict__ out, int H, int W) {
    CCL_COORDS
    if (!inb) return;
    out[tile + p] = keep[tile + p] ? idmap[tile + L[tile + p]] : 0;
}

int scan_rows_launch(int32_t* rowcnt, int32_t* n_out, int B, int H, cudaStream_t st) {
    CDNET_LAUNCH(k_scan_rows, B, 1024, 0, st, rowcnt, n_out, H);
    return last_error();
}

static int number_roots(int32_t* L, const uint8_t* keep, const uint8_t* excluded, int32_t* idmap, int32_t* rowcnt,
                        int32_t* n_out, int B, int H, int W, cudaStream_t st) {
    CDNET_CUDA_OK(cudaMemsetAsync(rowcnt, 0, sizeof(int32_t) * (size_t)B * H, st));
    if (v4_ok(W, L, keep, excluded)) CDNET_LAUNCH(k_flatten_count4, v4_grid(B, H, W), v4_block(), 0, st, L, keep, rowcnt, H, W, excluded);
    else CDNET_LAUNCH(k_flatten_count, ccl_grid(B, H, W), ccl_block(), 0, st, L, keep, rowcnt, H, W, excluded);
    CDNET_LAUNCH(k_scan_rows, B, 1024, 0, st, rowcnt, n_out, H);
    CDNET_LAUNCH(k_assign_ids, dim3(ceil_div(H, 8), B), 256, 0, st, L, keep, rowcnt, idmap, H, W, excluded);
    return last_error();
}

static int number_and_relabel(int32_t* L, const uint8_t* keep, int32_t* idmap, int32_t* rowcnt, int32_t* labels,
                              int32_t* n_out, int B, int H, int W, cudaStream_t st) {
    int rc = number_roots(L, keep, nullptr, idmap, rowcnt, n_out, B, H, W, st);
    if (rc) return rc;
    if (v4_ok(W, L, keep, idmap) && ((uintptr_t)labels & 15) == 0) CDNET_LAUNCH(k_relabel4, v4_grid(B, H, W), v4_block(), 0, st, L, keep, idmap, labels, H, W);
    else CDNET_LAUNCH(k_relabel, ccl_grid(B, H, W), ccl_block(), 0, st, L, keep, idmap, labels, H, W);
    return last_error();
}

int ccl_forest_launch(const uint8_t* mask, int32_t* L, int B, int H, int W, int conn, cudaStream_t st) {
    if (conn == 4) return forest_build<false, 4>(mask, L, nullptr, nullptr, B, H, W, st);
    if (conn == 8) return forest_build<false, 8>(mask, L, nullptr, nullptr, B, H, W, st);
    return CDNET_E_BADARG;
}

int ccl_label_launch(const uint8_t* mask, int32_t* labels, int32_t* n_out, int32_t* L, int32_t* idmap,
                     int32_t* rowcnt, int B, int H, int W, int conn, cudaStream_t st) {
    int rc = ccl_forest_launch(mask, L, B, H, W, conn, st);
    if (rc) return rc;
    return number_and_relabel(L, mask, idmap, rowcnt, labels, n_out, B, H, W, st);
}

// ---- skimage.measure.label of a multi-valued image ------------------------------------------------------
// 8-connected components of EQUAL non-zero value (my_transforms_direction.py:723-725: the instance map of the
// out_c != 3 form of LabelEncoding).  Rare path, kept simple: identity forest, every pixel is united with its
// equal-valued left / upper-left / upper / upper-right neighbour, then the usual raster-order numbering.
__global__ void k_values_init(int* __restrict__ L, size_t n, int plane) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        L[i] = (int)(i % (size_t)plane);
}

__global__ void __launch_bounds__(kBX* kBY) k_values_link(const uint8_t* __restrict__ ids, int* __restrict__ L, int H, int W) {
    CCL_COORDS
    if (!inb) return;
    const uint8_t* I = ids + tile;
    const int v = I[p];
    if (v == 0) return;
    int* Lt = L + tile;
    const bool left = x > 0 && I[p - 1] == v;
    if (left) uf_union(Lt, p, p - 1);
    if (y > 0) {
        const bool up = I[p - W] == v;
        if (up) uf_union(Lt, p, p - W);
        // a diagonal link is implied when the pixel between the two is already joined to both
        if (x > 0 && I[p - W - 1] == v && !(up || left)) uf_union(Lt, p, p - W - 1);
        if (x + 1 < W && I[p - W + 1] == v && !up) uf_union(Lt, p, p - W + 1);
    }
}

int ccl_label_values_launch(const uint8_t* ids, int32_t* labels, int32_t* n_out, int32_t* L, int32_t* idmap,
                            int32_t* rowcnt, int B, int H, int W, cudaStream_t st) {
    const size_t n = (size_t)B * H * W;
    const size_t blocks = (n + 1023) / 1024;
    CDNET_LAUNCH(k_values_init, (unsigned)(blocks > (1u << 20) ? (1u << 20) : blocks), 256, 0, st, L, n, H * W);
    CDNET_LAUNCH(k_values_link, ccl_grid(B, H, W), ccl_block(), 0, st, ids, L, H, W);
    return number_and_relabel(L, ids, idmap, rowcnt, labels, n_out, B, H, W, st);
}

// ---- fill holes ----------------------------------------------------------------------------------
// frame pixels that are background mark the root of their background component
__global__ void k_border_touch(const uint8_t* __restrict__ mask, const int* __restrict__ L, int* __restrict__ touch,
                               int H, int W, int top_frame, int bottom_frame) {
    const int b = blockIdx.y;
    const size_t tile = (size_t)b * H * W;
    const int per = 2 * W + 2 * H;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per; i += gridDim.x * blockDim.x) {
        int y, x;
        if (i < W) { y = 0; x = i; }
        else if (i < 2 * W) { y = H - 1; x = i - W; }
        else if (i < 2 * W + H) { y = i - 2 * W; x = 0; }
        else { y = i - 2 * W - H; x = W - 1; }
        if ((i < W && !top_frame) || (i >= W && i < 2 * W && !bottom_frame)) continue;  // shard seam, not the slide frame
        const int p = y * W + x;
        if (mask[tile + p] == 0) touch[tile + uf_find(L + tile, p)] = 1;
    }
}

// L <- roots; state = 1 foreground, 2 hole (background component that never reaches the frame), 0 else
__global__ void __launch_bounds__(kBX* kBY) k_flatten_fill(const uint8_t* __restrict__ mask, int* __restrict__ L,
                                                           const int* __restrict__ touch, uint8_t* __restrict__ state,
                                                           int H, int W) {
    CCL_COORDS
    if (!inb) return;
    int* Lt = L + tile;
    const int r = uf_find(Lt, p);
    Lt[p] = r;
    const bool fg = mask[tile + p] != 0;
    state[tile + p] = fg ? 1 : (touch[tile + r] ? 0 : 2);
}

int fill_holes_state_launch(const uint8_t* mask, uint8_t* state, int32_t* L, int32_t* touch, int B, int H, int W,
                            cudaStream_t st) {
    { int rc0 = forest_build<true, 4>(mask, L, touch, nullptr, B, H, W, st); if (rc0) return rc0; }
    CDNET_LAUNCH(k_border_touch, dim3(ceil_div(2 * W + 2 * H, 256), B), 256, 0, st, mask, L, touch, H, W, 1, 1);
    if (v4_ok(W, L, mask, state)) CDNET_LAUNCH(k_flatten_fill4, v4_grid(B, H, W), v4_block(), 0, st, mask, L, touch, state, H, W);
    else CDNET_LAUNCH(k_flatten_fill, ccl_grid(B, H, W), ccl_block(), 0, st, mask, L, touch, state, H, W);
    return last_error();
}

// hole pixels join every 4-adjacent foreground pixel -> forest of the filled mask's 4-conn components
__global__ void __launch_bounds__(kBX* kBY) k_fill_merge(const uint8_t* __restrict__ state, int* __restrict__ L, int H, int W) {
    CCL_COORDS
    if (!inb) return;
    const uint8_t* S = state + tile;
    if (S[p] != 2) return;
    int* Lt = L + tile;
    if (y > 0 && S[p - W] == 1) uf_union(Lt, p, p - W);
    if (x > 0 && S[p - 1] == 1) uf_union(Lt, p, p - 1);
    if (x + 1 < W && S[p + 1] == 1) uf_union(Lt, p, p + 1);
    if (y + 1 < H && S[p + W] == 1) uf_union(Lt, p, p + W);
}

// flatten + per-root pixel count of the pixels with state != 0 (warp-aggregated atomics)
__global__ void __launch_bounds__(kBX* kBY) k_flatten_area(const uint8_t* __restrict__ state, int* __restrict__ L,
                                                           int* __restrict__ area, int H, int W, int row_lo, int row_hi) {
    CCL_COORDS
    int r = -1;
    if (inb && state[tile + p]) {
        int* Lt = L + tile;
        r = uf_find(Lt, p);
        Lt[p] = r;
        if (y < row_lo || y >= row_hi) r = -1;  // ghost rows of a slide shard are counted by their owner
    }
    // lanes of one warp that share a root add once
    const unsigned peers = __match_any_sync(0xffffffffu, r);
    if (r >= 0 && lane == (__ffs(peers) - 1)) atomicAdd(area + tile + r, __popc(peers));
}

__global__ void __launch_bounds__(kBX* kBY) k_keep_large(const uint8_t* __restrict__ state, const int* __restrict__ L,
                                                         const int* __restrict__ area, uint8_t* __restrict__ keep,
                                                         int min_area, int H, int W) {
    CCL_COORDS
    if (!inb) return;
    keep[tile + p] = (state[tile + p] && area[tile + L[tile + p]] >= min_area) ? 1 : 0;
}

// kept 4-connected components that touch only diagonally are one 8-connected component
__global__ void __launch_bounds__(kBX* kBY) k_diag_merge(const uint8_t* __restrict__ keep, int* __restrict__ L, int H, int W) {
    CCL_COORDS
    if (!inb || y == 0) return;
    const uint8_t* K = keep + tile;
    if (!K[p] || K[p - W]) return;
    int* Lt = L + tile;
    if (x > 0 && K[p - W - 1] && !K[p - 1]) uf_union(Lt, p, p - W - 1);
    if (x + 1 < W && K[p - W + 1] && !K[p + 1]) uf_union(Lt, p, p - W + 1);
}

size_t fill_remove_label_workspace(int B, int H, int W) {
    const size_t n = (size_t)B * H * W;
    return 3 * pad256(n * 4) + 2 * pad256(n) + pad256((size_t)B * H * 4);
}

int fill_remove_label_launch(const uint8_t* inside, int32_t* labels, uint8_t* pred2_out, int B, int H, int W,
                             int min_area, void* ws, size_t ws_bytes, cudaStream_t st) {
    const size_t n = (size_t)B * H * W;
    Arena ar(ws, ws_bytes);
    int32_t* L = ar.take<int32_t>(n);
    int32_t* aux1 = ar.take<int32_t>(n);  // frame-touch flags, later id map
    int32_t* aux2 = ar.take<int32_t>(n);  // areas
    uint8_t* state = ar.take<uint8_t>(n);
    uint8_t* keep_ws = ar.take<uint8_t>(n);
    int32_t* rowcnt = ar.take<int32_t>((size_t)B * H);
    if (!ar.ok) return CDNET_E_WORKSPACE;
    uint8_t* keep = pred2_out ? pred2_out : keep_ws;
    { int rc0 = forest_build<true, 4>(inside, L, aux1, aux2, B, H, W, st); if (rc0) return rc0; }
    CDNET_LAUNCH(k_border_touch, dim3(ceil_div(2 * W + 2 * H, 256), B), 256, 0, st, inside, L, aux1, H, W, 1, 1);
    if (v4_ok(W, L, inside, state)) CDNET_LAUNCH(k_flatten_fill4, v4_grid(B, H, W), v4_block(), 0, st, inside, L, aux1, state, H, W);
    else CDNET_LAUNCH(k_flatten_fill, ccl_grid(B, H, W), ccl_block(), 0, st, inside, L, aux1, state, H, W);
    if (v4_ok(W, L, state)) CDNET_LAUNCH(k_fill_merge4, v4_grid(B, H, W), v4_block(), 0, st, state, L, H, W);
    else CDNET_LAUNCH(k_fill_merge, ccl_grid(B, H, W), ccl_block(), 0, st, state, L, H, W);
    if (v4_ok(W, L, state)) CDNET_LAUNCH(k_flatten_area4, v4_grid(B, H, W), v4_block(), 0, st, state, L, aux2, H, W, 0, H);
    else CDNET_LAUNCH(k_flatten_area, ccl_grid(B, H, W), ccl_block(), 0, st, state, L, aux2, H, W, 0, H);
    if (v4_ok(W, L, state, keep)) CDNET_LAUNCH(k_keep_large4, v4_grid(B, H, W), v4_block(), 0, st, state, L, aux2, keep, min_area, H, W);
    else CDNET_LAUNCH(k_keep_large, ccl_grid(B, H, W), ccl_block(), 0, st, state, L, aux2, keep, min_area, H, W);
    if (v4_ok(W, L, keep)) CDNET_LAUNCH(k_diag_merge4, v4_grid(B, H, W), v4_block(), 0, st, keep, L, H, W);
    else CDNET_LAUNCH(k_diag_merge, ccl_grid(B, H, W), ccl_block(), 0, st, keep, L, H, W);
    return number_and_relabel(L, keep, aux1, rowcnt, labels, nullptr, B, H, W, st);
}

// ---- integer remove_small_objects ------------------------------------------------------------------
__global__ void k_label_hist(const int* __restrict__ labels, int* __restrict__ counts, size_t plane) {
    const int b = blockIdx.y;
    const int* Lb = labels + (size_t)b * plane;
    int* cb = counts + (size_t)b * (plane + 1);
    if ((plane & 3) == 0 && (((uintptr_t)Lb) & 15) == 0) {
        // 128-bit loads; equal neighbours inside the quad share one atomic (a quad inside a nucleus: one)
        for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < plane / 4; q += (size_t)gridDim.x * blockDim.x) {
            const int4 v = *(const int4*)(Lb + 4 * q);
            if ((v.x | v.y | v.z | v.w) == 0) continue;
            const int vs[4] = {v.x, v.y, v.z, v.w};
            int cur = vs[0], c = 1;
#pragma unroll
            for (int k = 1; k < 4; ++k) {
                if (vs[k] == cur) { ++c; continue; }
                if (cur > 0) atomicAdd(cb + cur, c);
                cur = vs[k];
                c = 1;
            }
            if (cur > 0) atomicAdd(cb + cur, c);
        }
        return;
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < ((plane + 31) & ~size_t(31));
         i += (size_t)gridDim.x * blockDim.x) {
        const int v = i < plane ? Lb[i] : 0;
        const unsigned nz = __ballot_sync(0xffffffffu, v > 0);
        if (!nz) continue;  // 32 unlabelled pixels: most of a tile
        const int lane = threadIdx.x & 31, first = __ffs(nz) - 1;
        const int v0 = __shfl_sync(0xffffffffu, v, first);
        if (__all_sync(0xffffffffu, v <= 0 || v == v0)) {  // one label among them: the inside of a nucleus
            if (lane == first) atomicAdd(cb + v0, __popc(nz));
            continue;
        }
        const unsigned peers = __match_any_sync(0xffffffffu, v);
        if (v > 0 && lane == (__ffs(peers) - 1)) atomicAdd(cb + v, __popc(peers));
    }
}
__global__ void k_label_drop_small(int* __restrict__ labels, const int* __restrict__ counts, size_t plane, int min_size) {
    const int b = blockIdx.y;
    int* Lb = labels + (size_t)b * plane;
    const int* cb = counts + (size_t)b * (plane + 1);
    if ((plane & 3) == 0 && (((uintptr_t)Lb) & 15) == 0) {  // 128-bit loads; a quad is stored only when it changes
        for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < plane / 4; q += (size_t)gridDim.x * blockDim.x) {
            int4 v = *(const int4*)(Lb + 4 * q);
            if ((v.x | v.y | v.z | v.w) == 0) continue;
            bool ch = false;
            if (v.x > 0 && cb[v.x] < min_size) { v.x = 0; ch = true; }
            if (v.y > 0 && cb[v.y] < min_size) { v.y = 0; ch = true; }
            if (v.z > 0 && cb[v.z] < min_size) { v.z = 0; ch = true; }
            if (v.w > 0 && cb[v.w] < min_size) { v.w = 0; ch = true; }
            if (ch) *(int4*)(Lb + 4 * q) = v;
        }
        return;
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (size_t)gridDim.x * blockDim.x) {
        const int v = Lb[i];
        if (v > 0 && cb[v] < min_size) Lb[i] = 0;
    }
}

int remove_small_labels_launch(int32_t* labels, int32_t* counts, int B, int H, int W, int min_size, cudaStream_t st) {
    const size_t plane = (size_t)H * W;
    CDNET_CUDA_OK(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)B * (plane + 1), st));
    dim3 grid((unsigned)((plane + 256 * 4 - 1) / (256 * 4)), B);
    CDNET_LAUNCH(k_label_hist, grid, 256, 0, st, labels, counts, plane);
    CDNET_LAUNCH(k_label_drop_small, grid, 256, 0, st, labels, counts, plane, min_size);
    return last_error();
}

__global__ void k_state_to_mask(const uint8_t* __restrict__ state, uint8_t* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = state[i] ? 1 : 0;
}

}  // namespace cdnet

using namespace cdnet;

static bool bad_dims(int B, int H, int W) { return B <= 0 || H <= 0 || W <= 0 || (double)H * W >= 2147483648.0; }

extern "C" size_t cdnet_ccl_workspace_bytes(int B, int H, int W) {
    if (bad_dims(B, H, W)) return 0;
    const size_t n = (size_t)B * H * W;
    return 2 * pad256(n * 4) + pad256((size_t)B * H * 4);
}

extern "C" int cdnet_ccl(const uint8_t* mask, int32_t* labels, int32_t* n_out, int B, int H, int W, int connectivity,
                         void* ws, size_t ws_bytes, void* stream) {
    if (!mask || !labels || bad_dims(B, H, W) || (connectivity != 4 && connectivity != 8)) return CDNET_E_BADARG;
    Arena ar(ws, ws_bytes);
    const size_t n = (size_t)B * H * W;
    int32_t* L = ar.take<int32_t>(n);
    int32_t* idmap = ar.take<int32_t>(n);
    int32_t* rowcnt = ar.take<int32_t>((size_t)B * H);
    if (!ar.ok) return CDNET_E_WORKSPACE;
    return ccl_label_launch(mask, labels, n_out, L, idmap, rowcnt, B, H, W, connectivity, (cudaStream_t)stream);
}

extern "C" int cdnet_label_values(const uint8_t* ids, int32_t* labels, int32_t* n_out, int B, int H, int W, void* ws,
                                  size_t ws_bytes, void* stream) {
    if (!ids || !labels || bad_dims(B, H, W)) return CDNET_E_BADARG;
    Arena ar(ws, ws_bytes);
    const size_t n = (size_t)B * H * W;
    int32_t* L = ar.take<int32_t>(n);
    int32_t* idmap = ar.take<int32_t>(n);
    int32_t* rowcnt = ar.take<int32_t>((size_t)B * H);
    if (!ar.ok) return CDNET_E_WORKSPACE;
    return ccl_label_values_launch(ids, labels, n_out, L, idmap, rowcnt, B, H, W, (cudaStream_t)stream);
}

extern "C" size_t cdnet_fill_holes_workspace_bytes(int B, int H, int W) {
    if (bad_dims(B, H, W)) return 0;
    const size_t n = (size_t)B * H * W;
    return 2 * pad256(n * 4) + pad256(n);
}

extern "C" int cdnet_fill_holes(const uint8_t* mask, uint8_t* out, int B, int H, int W, void* ws, size_t ws_bytes,
                                void* stream) {
    if (!mask || !out || bad_dims(B, H, W)) return CDNET_E_BADARG;
    Arena ar(ws, ws_bytes);
    const size_t n = (size_t)B * H * W;
    int32_t* L = ar.take<int32_t>(n);
    int32_t* touch = ar.take<int32_t>(n);
    uint8_t* state = ar.take<uint8_t>(n);
    if (!ar.ok) return CDNET_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = fill_holes_state_launch(mask, state, L, touch, B, H, W, st);
    if (rc) return rc;
    const size_t blocks = (n + 1023) / 1024;
    CDNET_LAUNCH(k_state_to_mask, (unsigned)(blocks > (1u << 20) ? (1u << 20) : blocks), 256, 0, st, state, out, n);
    return last_error();
}

extern "C" size_t cdnet_remove_small_mask_workspace_bytes(int B, int H, int W) {
    if (bad_dims(B, H, W)) return 0;
    const size_t n = (size_t)B * H * W;
    return 2 * pad256(n * 4);
}

extern "C" int cdnet_remove_small_mask(const uint8_t* mask, uint8_t* out, int B, int H, int W, int min_size, void* ws,
                                       size_t ws_bytes, void* stream) {
    if (!mask || !out || bad_dims(B, H, W)) return CDNET_E_BADARG;
    Arena ar(ws, ws_bytes);
    const size_t n = (size_t)B * H * W;
    int32_t* L = ar.take<int32_t>(n);
    int32_t* area = ar.take<int32_t>(n);
    if (!ar.ok) return CDNET_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    { int rc0 = forest_build<false, 4>(mask, L, area, nullptr, B, H, W, st); if (rc0) return rc0; }
    if (v4_ok(W, L, mask)) CDNET_LAUNCH(k_flatten_area4, v4_grid(B, H, W), v4_block(), 0, st, mask, L, area, H, W, 0, H);
    else CDNET_LAUNCH(k_flatten_area, ccl_grid(B, H, W), ccl_block(), 0, st, mask, L, area, H, W, 0, H);
    if (v4_ok(W, L, mask, out)) CDNET_LAUNCH(k_keep_large4, v4_grid(B, H, W), v4_block(), 0, st, mask, L, area, out, min_size, H, W);
    else CDNET_LAUNCH(k_keep_large, ccl_grid(B, H, W), ccl_block(), 0, st, mask, L, area, out, min_size, H, W);
    return last_error();
}

extern "C" size_t cdnet_remove_small_labels_workspace_bytes(int B, int H, int W) {
    if (bad_dims(B, H, W)) return 0;
    return pad256((size_t)B * ((size_t)H * W + 1) * 4);
}

extern "C" int cdnet_remove_small_labels(int32_t* labels, int B, int H, int W, int min_size, void* ws, size_t ws_bytes,
                                         void* stream) {
    if (!labels || bad_dims(B, H, W)) return CDNET_E_BADARG;
    Arena ar(ws, ws_bytes);
    int32_t* counts = ar.take<int32_t>((size_t)B * ((size_t)H * W + 1));
    if (!ar.ok) return CDNET_E_WORKSPACE;
    return remove_small_labels_launch(labels, counts, B, H, W, min_size, (cudaStream_t)stream);
}

// =====================================================================================================
// whole-slide row shards: the fill-holes / remove-small / 8-conn label chain cut
// into stages so that the host can reconcile components that straddle shard seams between the stages
// (cdnet_b200/sharded.py).  One extended tile per call: the shard's own rows plus one ghost row of the
// neighbouring shard on each inner side.  All planes are [He, W].
// =====================================================================================================
extern "C" int cdnet_shard_label_stage1(const uint8_t* inside, int32_t* L, int32_t* touch, int He, int W, int top_is_frame,
                                        int bottom_is_frame, void* stream) {
    if (!inside || !L || !touch || bad_dims(1, He, W)) return CDNET_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int B = 1, H = He;
    { int rc0 = forest_build<true, 4>(inside, L, touch, nullptr, B, H, W, st); if (rc0) return rc0; }
    CDNET_LAUNCH(k_border_touch, dim3(ceil_div(2 * W + 2 * H, 256), B), 256, 0, st, inside, L, touch, H, W, top_is_frame,
                 bottom_is_frame);
    CDNET_LAUNCH(k_flatten, ccl_grid(B, H, W), ccl_block(), 0, st, L, H, W);
    return last_error();
}

// touch[] now holds slide-global flags for the roots of seam components.  state, hole merge, areas of the
// rows [row_lo, row_hi) only (area must be zero-filled by the caller).
extern "C" int cdnet_shard_label_stage2(const uint8_t* inside, int32_t* L, const int32_t* touch, uint8_t* state,
                                        int32_t* area, int He, int W, int row_lo, int row_hi, void* stream) {
    if (!inside || !L || !touch || !state || !area || bad_dims(1, He, W)) return CDNET_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int B = 1, H = He;
    if (v4_ok(W, L, inside, state)) CDNET_LAUNCH(k_flatten_fill4, v4_grid(B, H, W), v4_block(), 0, st, inside, L, touch, state, H, W);
    else CDNET_LAUNCH(k_flatten_fill, ccl_grid(B, H, W), ccl_block(), 0, st, inside, L, touch, state, H, W);
    if (v4_ok(W, L, state)) CDNET_LAUNCH(k_fill_merge4, v4_grid(B, H, W), v4_block(), 0, st, state, L, H, W);
    else CDNET_LAUNCH(k_fill_merge, ccl_grid(B, H, W), ccl_block(), 0, st, state, L, H, W);
    if (v4_ok(W, L, state)) CDNET_LAUNCH(k_flatten_area4, v4_grid(B, H, W), v4_block(), 0, st, state, L, area, H, W, row_lo, row_hi);
    else CDNET_LAUNCH(k_flatten_area, ccl_grid(B, H, W), ccl_block(), 0, st, state, L, area, H, W, row_lo, row_hi);
    return last_error();
}

// area[] now holds slide-global pixel counts for the roots of seam components.
extern "C" int cdnet_shard_label_stage3(const uint8_t* state, int32_t* L, const int32_t* area, uint8_t* keep, int min_area,
                                        int He, int W, void* stream) {
    if (!state || !L || !area || !keep || bad_dims(1, He, W)) return CDNET_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int B = 1, H = He;
    if (v4_ok(W, L, state, keep)) CDNET_LAUNCH(k_keep_large4, v4_grid(B, H, W), v4_block(), 0, st, state, L, area, keep, min_area, H, W);
    else CDNET_LAUNCH(k_keep_large, ccl_grid(B, H, W), ccl_block(), 0, st, state, L, area, keep, min_area, H, W);
    if (v4_ok(W, L, keep)) CDNET_LAUNCH(k_diag_merge4, v4_grid(B, H, W), v4_block(), 0, st, keep, L, H, W);
    else CDNET_LAUNCH(k_diag_merge, ccl_grid(B, H, W), ccl_block(), 0, st, keep, L, H, W);
    CDNET_LAUNCH(k_flatten, ccl_grid(B, H, W), ccl_block(), 0, st, L, H, W);
    return last_error();
}

// local ids 1..n_owned (raster order) for the kept roots that are not `excluded`; rowcnt: int32 [He] scratch
extern "C" int cdnet_shard_label_stage4(int32_t* L, const uint8_t* keep, const uint8_t* excluded, int32_t* idmap,
                                        int32_t* rowcnt, int32_t* n_owned, int He, int W, void* stream) {
    if (!L || !keep || !idmap || !rowcnt || !n_owned || bad_dims(1, He, W)) return CDNET_E_BADARG;
    return number_roots(L, keep, excluded, idmap, rowcnt, n_owned, 1, He, W, (cudaStream_t)stream);
}

extern "C" int cdnet_shard_relabel(const int32_t* L, const uint8_t* keep, const int32_t* idmap, int32_t* labels, int He,
                                   int W, void* stream) {
    if (!L || !keep || !idmap || !labels || bad_dims(1, He, W)) return CDNET_E_BADARG;
    const int B = 1, H = He;
    if (v4_ok(W, L, keep, idmap) && ((uintptr_t)labels & 15) == 0) CDNET_LAUNCH(k_relabel4, v4_grid(B, H, W), v4_block(), 0, (cudaStream_t)stream, L, keep, idmap, labels, H, W);
    else CDNET_LAUNCH(k_relabel, ccl_grid(B, H, W), ccl_block(), 0, (cudaStream_t)stream, L, keep, idmap, labels, H, W);
    return last_error();
}
