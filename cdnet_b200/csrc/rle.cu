// rle.cu -- run-based tail of the inference post-processing (K6, K7, K8, K13 of SURVEY.md section 2.1 in one chain):
//   binary_fill_holes -> remove_small_objects(min_area, 4-connected) -> measure.label (8-connected, raster-first
//   numbering) -> dilation by disk(radius)                                   test_dam.py:546-563, test.py:277-295
//
// The per-pixel union-find chain of ccl.cu streams a 4-byte parent plane eight times.  Here the mask lives as ONE
// BIT per pixel (a 1000 x 1000 tile is 125 KB) and the graph nodes are the horizontal RUNS of equal mask value
// (a MoNuSeg-shaped row has ~30 of them, not 1000 pixels): every stage but the last is a pass over the bit-planes
// with one warp per row -- run starts, run membership and row-to-row adjacency are a handful of bit operations per
// 32-pixel word -- plus SPARSE accesses to two pixel-indexed int32 planes (parent / area-or-id), touched only at run
// starts.  The only pixel-granular pass is the last one, which writes the dilated labels.
//
//   k_rle_pack     mask bytes -> bit-plane M; parent[start] = start, aux[start] = 0 at every run start
//   k_rle_link     runs of equal value that overlap in adjacent rows are united (4-connectivity; foreground AND
//                  background components in one forest; the root is the component's first raster pixel)
//                  background runs on the image frame hang below the special root -1 ("outside") from the start
//   k_rle_holes    bit-plane F = M | (background runs whose component never reached -1) = binary_fill_holes(M)
//   k_rle_fill     hole runs are united with the foreground runs they touch (left / right / above / below)
//   k_rle_area     per-component pixel counts (one atomicAdd per filled segment of a word); flattens the run starts
//   k_rle_diag     components with area >= min_area that touch only diagonally are united (8-connectivity of the
//                  kept mask; removed components never join anything)
//   k_rle_count / scan / k_rle_assign   raster-order rank of the surviving roots = the label ids
//   k_rle_labels   every pixel looks up the id of its run's root (0 outside F / for removed components), a block
//                  stages TR + 2R label rows in shared memory and writes max over disk(R) as int32 or int64
//
// A pixel's run is found without any per-pixel table: the run start is the highest set bit of the row's transition
// mask at or below the pixel, or -- when the word holds none -- the carry-in start that a warp-wide max-scan over the
// row's words provides.  Two runs of adjacent rows overlap iff they share a column, and the first shared column is a
// run start of one of them, so the adjacency events of a word are  same_value & (starts_cur | starts_prev).
#include <stdlib.h>

#include "internal.h"

namespace cdnet {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kOutside = -1;  // the special root of background that touches the image frame (see rf_find)
// parent and aux (area, then id) of a run start sit side by side: node i = words 2 i, 2 i + 1 of ONE plane, so that the
// sparse accesses that need both (numbering, labels, the first store) touch one sector instead of two
constexpr size_t kNS = 2;  // size_t: 2 * (pixel index) leaves the int range on tiles beyond 2^30 pixels (a 40 000^2 slide)
constexpr int kRleWarps = 8;  // rows per block, one warp per row

struct RowScan {
    uint32_t m;      // mask bits of this lane's word (bit k = pixel wx + k), 0 beyond the row
    uint32_t t;      // run starts: pixels whose value differs from their left neighbour (and pixel 0)
    uint32_t valid;  // pixels of the word that exist
    int cin;         // latest run start left of this word (-1 for the first word) = start of the run of pixel wx - 1
    int wx;          // x of bit 0
};

// run starts of one word from the word, the last pixel of its left neighbour and the stored carry-in start
__device__ __forceinline__ RowScan row_word(uint32_t word, uint32_t left_word, int cin, int W, int wj) {
    RowScan r;
    r.wx = wj * 32;
    const int rem = W - r.wx;
    r.valid = rem >= 32 ? kFull : (rem > 0 ? ((1u << rem) - 1u) : 0u);
    r.m = word & r.valid;
    r.t = (r.m ^ ((r.m << 1) | (left_word >> 31))) & r.valid;
    if (wj == 0) r.t |= 1u;
    r.cin = cin;
    return r;
}

// row_bits / row_cin: the row's words of the mask bit-plane and of the carry-in plane; words beyond the row read as 0
__device__ __forceinline__ RowScan row_load(const uint32_t* __restrict__ row_bits, const int* __restrict__ row_cin, int NW,
                                            int W, int wj) {
    const bool in = wj < NW;
    return row_word(in ? row_bits[wj] : 0u, (in && wj > 0) ? row_bits[wj - 1] : 0u, in ? row_cin[wj] : -1, W, wj);
}

// start (x) of the run that contains bit k of the word
__device__ __forceinline__ int run_start(const RowScan& r, int k) {
    const uint32_t tt = r.t & (kFull >> (31 - k));
    return tt ? (r.wx + 31 - __clz(tt)) : r.cin;
}

// bit k = value of the LEFT neighbour of pixel k (the last pixel of the previous word for k = 0)
__device__ __forceinline__ uint32_t left_bits(const uint32_t* __restrict__ row, int NW, int wj, uint32_t own) {
    return (own << 1) | ((wj > 0 && wj <= NW) ? (row[wj - 1] >> 31) : 0u);
}

#define RLE_ROW_COORDS                                          \
    pdl_wait();                                                 \
    pdl_trigger();                                              \
    const int lane = threadIdx.x & 31;                          \
    const int y = blockIdx.x * kRleWarps + (threadIdx.x >> 5);  \
    const int b = blockIdx.y;                                   \
    if (y >= H) return;                                         \
    const size_t tile = (size_t)b * H * W;                      \
    const int NW = (W + 31) >> 5;                               \
    const int nchunks = (NW + 31) >> 5;                         \
    const size_t rowbits = ((size_t)b * H + y) * NW;            \
    (void)tile; (void)nchunks; (void)rowbits; (void)lane;

static inline dim3 rle_grid(int B, int H) { return dim3(ceil_div(H, kRleWarps), B); }
static inline size_t pack_smem(int rows) { return (size_t)rows * 96 * 4 + ((size_t)rows * 1024 + 2) * 2; }
static inline dim3 link_grid(int B, int H, int mod_lo) { return dim3(ceil_div(ceil_div(H, mod_lo), 2), B); }

// 32 mask bytes -> 32 bits
__device__ __forceinline__ uint32_t nz_nibble(uint32_t w) {  // 4 bytes -> 4 bits (bit i = byte i != 0)
    const uint32_t nz = ((((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) >> 7) & 0x01010101u;
    return (nz * 0x10204080u) >> 28;
}

// the only warp-synchronous kernel of the chain: the carry-in starts are a max-scan over the words of a row
__global__ void __launch_bounds__(32 * kRleWarps) k_rle_pack(const uint8_t* __restrict__ mask, uint32_t* __restrict__ M,
                                                             int* __restrict__ C, int* __restrict__ P, int* __restrict__ A,
                                                             int H, int W) {
    RLE_ROW_COORDS
    const uint8_t* row = mask + tile + (size_t)y * W;
    const bool vec = (W % 8 == 0) && (((uintptr_t)mask & 7) == 0);
    uint32_t carry_word = 0;  // last word of the previous chunk
    int carry_start = -1;     // latest run start of the previous chunks
    uint32_t last_word = 0;
    for (int ch = 0; ch < nchunks; ++ch) {
        const int wj = ch * 32 + lane;
        const int x0 = wj * 32;
        uint32_t word = 0;
        if (x0 < W) {
            if (vec) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (x0 + 8 * k < W) {
                        const uint2 v = __ldg((const uint2*)(row + x0 + 8 * k));
                        word |= (nz_nibble(v.x) | (nz_nibble(v.y) << 4)) << (8 * k);
                    }
                }
            } else {
                const int n = min(32, W - x0);
                for (int k = 0; k < n; ++k) word |= (uint32_t)(__ldg(row + x0 + k) != 0) << k;
            }
        }
        const uint32_t up = __shfl_up_sync(kFull, word, 1);
        RowScan r = row_word(word, lane ? up : carry_word, -1, W, wj);
        const int last = r.t ? (r.wx + 31 - __clz(r.t)) : -1;
        int incl = last;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl = max(incl, v);
        }
        int excl = __shfl_up_sync(kFull, incl, 1);
        if (lane == 0) excl = -1;
        r.cin = max(carry_start, excl);
        carry_start = max(carry_start, __shfl_sync(kFull, incl, 31));
        carry_word = __shfl_sync(kFull, word, 31);
        if (wj < NW) {
            M[rowbits + wj] = r.m;
            C[rowbits + wj] = r.cin;
        }
        uint32_t t = r.t;
        while (t) {
            const int k = __ffs(t) - 1;
            t &= t - 1;
            const int gid = y * W + r.wx + k;
            const bool frame_bg = !((r.m >> k) & 1u) && (y == 0 || y == H - 1 || gid == y * W);
            *(int2*)(P + kNS * (tile + gid)) = make_int2(frame_bg ? kOutside : gid, 0);
        }
        last_word = __shfl_sync(kFull, r.m, (NW - 1) & 31);  // meaningful after the last chunk
    }
    // the last run of the row, if background, touches the frame too
    __syncwarp();
    if (lane == 0 && !((last_word >> ((W - 1) & 31)) & 1u)) P[kNS * (tile + y * W + carry_start)] = kOutside;
}

// ---- tiles up to 1024 columns: pack + the links INSIDE groups of kRleWarps rows in one kernel ---------------------
// A block owns kRleWarps consecutive rows.  Their runs are united in shared memory first (parent slots indexed by the
// pixel position inside the group, ~30-cycle hops, no global atomics), then every run start is written out already
// pointing at the root of its group-local tree.  What is left for global memory are the links across the group seams
// (k_rle_link on rows y % kRleWarps == 0): an eighth of the unions, on trees of depth one.
// The forest of the link phase knows one special root, kOutside = -1: a background run that touches the image frame is
// born with it as its parent.  Hooking always goes to the smaller index, so -1 wins every union it takes part in, and
//   * "is this background run a hole?" is "does its find() end at a real root?" -- no frame-flag pass;
//   * the one huge component of a tile, its background, is never assembled: its runs hang two hops below -1, and a
//     union of two runs that are both outside finds equal roots and touches no atomic.
// Foreground runs never meet -1 (only runs of equal value are united before the holes are known).

// path halving with plain stores: parents only ever point at smaller indices of the same tree, so a stale store can at
// worst undo a little compression (shared or global memory)
__device__ __forceinline__ int rf_find(int* P, int p) {
    if (p < 0) return kOutside;
    int q = P[kNS * p];
    while (q != p) {
        if (q < 0) return kOutside;
        const int g = P[kNS * q];
        if (g != q) P[kNS * p] = g;
        p = q;
        q = g;
    }
    return p;
}

// read-only walk (foreground runs after the link phase: never outside)
__device__ __forceinline__ int rf_find_ro(const int* P, int p) {
    int q = P[kNS * p];
    while (q != p) {
        p = q;
        q = P[kNS * p];
    }
    return p;
}

// both finds of a union at once: the two chains of dependent loads overlap
__device__ __forceinline__ void rf_find2(int* P, int& a, int& b) {
    int qa = a >= 0 ? P[kNS * a] : a, qb = b >= 0 ? P[kNS * b] : b;
    while (qa != a || qb != b) {
        const int ga = (qa != a && qa >= 0) ? P[kNS * qa] : qa;
        const int gb = (qb != b && qb >= 0) ? P[kNS * qb] : qb;
        if (qa != a) {
            if (qa >= 0 && ga != qa) P[kNS * a] = ga;
            a = qa;
            qa = ga;
        }
        if (qb != b) {
            if (qb >= 0 && gb != qb) P[kNS * b] = gb;
            b = qb;
            qb = gb;
        }
    }
}

__device__ __forceinline__ void rf_union(int* P, int a, int b) {
    for (;;) {
        rf_find2(P, a, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }
        const int old = atomicMin(P + kNS * a, b);  // a > b >= -1: hang a under b
        if (old == a) return;
        a = old;
    }
}

// The group-local forest of k_rle_pack_link: 16-bit parents in shared memory (32 rows x 1024 positions + 1 fit), slot 0
// = outside (a root that is smaller than every other, so it wins every union by itself), slot 1 + row * 1024 + x = the
// run that starts at (row, x).  Hooking is a compare-and-swap of a root's self-link.
__device__ __forceinline__ int sf_find(unsigned short* P, int p) {
    int q = P[p];
    while (q != p) {
        const int g = P[q];
#ifndef CDNET_SF_NO_HALVING
        if (g != q) P[p] = (unsigned short)g;
#endif
        p = q;
        q = g;
    }
    return p;
}
__device__ __forceinline__ void sf_union(unsigned short* P, int a, int b) {
    for (;;) {
        a = sf_find(P, a);
        b = sf_find(P, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }
        const int old = atomicCAS(P + a, (unsigned short)a, (unsigned short)b);  // a > b: hang root a under b
        if (old == a) return;
        a = old;
    }
}

// ERODED: the input is not a byte mask but a bit-plane (the filled plane F of a previous chain, passed through `mask`),
// eroded on the fly with the cross structure and a zero border -- binary_erosion(iterations=1) of postproc_other.py:43
// between the hole filling and the labelling of the markers, without ever leaving the bit domain.
template <int ROWS, bool ERODED = false>
__global__ void __launch_bounds__(32 * ROWS) k_rle_pack_link(const uint8_t* __restrict__ mask, uint32_t* __restrict__ M,
                                                             int* __restrict__ C, int* __restrict__ P,
                                                             int* __restrict__ A, int H, int W) {
    pdl_wait();
    pdl_trigger();
    CDNET_DYN_SHARED(int, s_dyn);  // ROWS * 96 words, then ROWS * 1024 + 2 16-bit parents
    uint32_t(*s_m)[32] = reinterpret_cast<uint32_t(*)[32]>(s_dyn);
    uint32_t(*s_t)[32] = reinterpret_cast<uint32_t(*)[32]>(s_dyn + ROWS * 32);
    int(*s_c)[32] = reinterpret_cast<int(*)[32]>(s_dyn + ROWS * 64);
    unsigned short* s_par = reinterpret_cast<unsigned short*>(s_dyn + ROWS * 96);
    if (threadIdx.x == 0) s_par[0] = 0;  // outside
    const int lane = threadIdx.x & 31, rw = threadIdx.x >> 5;
    const int y0 = blockIdx.x * ROWS, y = y0 + rw;
    const int b = blockIdx.y;
    const size_t tile = (size_t)b * H * W;
    const int NW = (W + 31) >> 5;  // <= 32
    const bool row_ok = y < H;
    RowScan r = row_word(0u, 0u, -1, W, lane);
    if (row_ok) {
        const uint8_t* row = mask + tile + (size_t)y * W;
        const bool vec = (W % 8 == 0) && (((uintptr_t)mask & 7) == 0);
        const int x0 = lane * 32;
        uint32_t word = 0;
        if (ERODED) {
            const uint32_t* Fp = reinterpret_cast<const uint32_t*>(mask) + ((size_t)b * H + y) * NW;
            const uint32_t c = lane < NW ? Fp[lane] : 0u;
            const uint32_t lw = __shfl_up_sync(kFull, c, 1), rw_ = __shfl_down_sync(kFull, c, 1);
            if (lane < NW && y > 0 && y + 1 < H) {
                const uint32_t left = (c << 1) | (lane ? (lw >> 31) : 0u);                  // bit k = pixel k - 1
                const uint32_t right = (c >> 1) | (lane + 1 < NW ? (rw_ << 31) : 0u);       // bit k = pixel k + 1
                word = c & Fp[lane - NW] & Fp[lane + NW] & left & right;                    // pixels beyond W are 0 in F
            }
        } else if (x0 < W) {
            if (vec) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (x0 + 8 * k < W) {
                        const uint2 v = __ldg((const uint2*)(row + x0 + 8 * k));
                        word |= (nz_nibble(v.x) | (nz_nibble(v.y) << 4)) << (8 * k);
                    }
                }
            } else {
                const int n = min(32, W - x0);
                for (int k = 0; k < n; ++k) word |= (uint32_t)(__ldg(row + x0 + k) != 0) << k;
            }
        }
        const uint32_t up = __shfl_up_sync(kFull, word, 1);
        r = row_word(word, lane ? up : 0u, -1, W, lane);
        int incl = r.t ? (r.wx + 31 - __clz(r.t)) : -1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl = max(incl, v);
        }
        int excl = __shfl_up_sync(kFull, incl, 1);
        if (lane == 0) excl = -1;
        r.cin = excl;
        if (lane < NW) {
            const size_t rowbits = ((size_t)b * H + y) * NW;
            M[rowbits + lane] = r.m;
            C[rowbits + lane] = r.cin;
        }
        uint32_t t = r.t;
        while (t) {
            const int k = __ffs(t) - 1;
            t &= t - 1;
            const int li = 1 + rw * 1024 + r.wx + k;
            s_par[li] = (unsigned short)li;
        }
        // background runs on the image frame start out below kOutside: every background run of the first and the last
        // row, the first and the last run of the others
        __syncwarp();
        if (y == 0 || y == H - 1) {
            uint32_t t0 = r.t & ~r.m;
            while (t0) {
                const int k = __ffs(t0) - 1;
                t0 &= t0 - 1;
                s_par[1 + rw * 1024 + r.wx + k] = 0;
            }
        } else {
            if (lane == 0 && !(r.m & 1u)) s_par[1 + rw * 1024] = 0;
            if (lane == ((W - 1) >> 5) && !((r.m >> ((W - 1) & 31)) & 1u)) s_par[1 + rw * 1024 + run_start(r, (W - 1) & 31)] = 0;
        }
    }
    s_m[rw][lane] = r.m;
    s_t[rw][lane] = r.t;
    s_c[rw][lane] = r.cin;
    __syncthreads();
    if (row_ok && rw > 0) {
        RowScan prv = r;  // same wx / valid
        prv.m = s_m[rw - 1][lane];
        prv.t = s_t[rw - 1][lane];
        prv.cin = s_c[rw - 1][lane];
        uint32_t e = ~(r.m ^ prv.m) & r.valid & (r.t | prv.t);
        while (e) {
            const int k = __ffs(e) - 1;
            e &= e - 1;
            sf_union(s_par, 1 + rw * 1024 + run_start(r, k), 1 + (rw - 1) * 1024 + run_start(prv, k));
        }
    }
    __syncthreads();
    if (row_ok) {
        uint32_t t = r.t;
        while (t) {
            const int k = __ffs(t) - 1;
            t &= t - 1;
            const int root = sf_find(s_par, 1 + rw * 1024 + r.wx + k) - 1;
            const int gid = y * W + r.wx + k;
            *(int2*)(P + kNS * (tile + gid)) = make_int2(root < 0 ? kOutside : (y0 + (root >> 10)) * W + (root & 1023), 0);
        }
    }
}

// rows y with y % mod_lo == 0 and (mod_hi == 0 or y % mod_hi != 0) are linked to the row above them.  Three launches
// (rows inside groups of 8, the seams of those inside groups of 64, the remaining seams) keep the trees shallow and the
// atomics spread out: the one huge component of a tile -- its background -- is then assembled from a few dozen
// sub-trees instead of being fought over by every row at once.
constexpr int kLinkWarps = 2;  // few warps per block: the seam rows of a launch spread over all SMs
__global__ void __launch_bounds__(32 * kLinkWarps) k_rle_link(const uint32_t* __restrict__ M, const int* __restrict__ C,
                                                              int* __restrict__ P, int H, int W, int mod_lo, int mod_hi) {
    pdl_wait();
    pdl_trigger();
    // one warp per candidate row y = mod_lo, 2 mod_lo, ..
    const int lane = threadIdx.x & 31;
    const int y = (blockIdx.x * kLinkWarps + (threadIdx.x >> 5) + 1) * mod_lo;
    const int b = blockIdx.y;
    if (y >= H || (mod_hi && (y % mod_hi) == 0)) return;
    const size_t tile = (size_t)b * H * W;
    const int NW = (W + 31) >> 5;
    const size_t rowbits = ((size_t)b * H + y) * NW;
    int* Pt = P + kNS * tile;
    for (int wj = lane; wj < NW; wj += 32) {
        const RowScan cur = row_load(M + rowbits, C + rowbits, NW, W, wj);
        const RowScan prv = row_load(M + rowbits - NW, C + rowbits - NW, NW, W, wj);
        uint32_t e = ~(cur.m ^ prv.m) & cur.valid & (cur.t | prv.t);
        while (e) {
            // four events per round: the first two hops of their eight find chains are independent loads.  A union of
            // two ancestors unites the same sets, so the unions start from the grandparents.
            int ea[4], eb[4];
            bool ev[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                ev[i] = e != 0;
                ea[i] = eb[i] = kOutside;
                if (e) {
                    const int k = __ffs(e) - 1;
                    e &= e - 1;
                    ea[i] = y * W + run_start(cur, k);
                    eb[i] = (y - 1) * W + run_start(prv, k);
                }
            }
#pragma unroll
            for (int hop = 0; hop < 2; ++hop) {
                int na[4], nb[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    na[i] = ea[i] >= 0 ? Pt[kNS * ea[i]] : kOutside;
                    nb[i] = eb[i] >= 0 ? Pt[kNS * eb[i]] : kOutside;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    ea[i] = na[i];
                    eb[i] = nb[i];
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (ev[i] && ea[i] != eb[i]) rf_union(Pt, ea[i], eb[i]);
        }
    }
}

// segments of set bits inside one word: calls fn(first bit, bit mask of the segment)
template <typename Fn>
__device__ __forceinline__ void for_each_segment(uint32_t bits, Fn fn) {
    uint32_t s = bits & ~(bits << 1);
    while (s) {
        const int k0 = __ffs(s) - 1;
        s &= s - 1;
        const uint32_t inv = ~(bits >> k0);
        const int len = inv ? (__ffs(inv) - 1) : (32 - k0);
        const uint32_t seg = (len >= 32 ? kFull : ((1u << len) - 1u)) << k0;
        fn(k0, seg);
    }
}

// F = M | holes, and every hole run is united with the foreground runs it touches (left, right, above, below), so that
// afterwards the forest holds the 4-connected components of the FILLED mask.  A background run is a hole iff the root
// of its component carries no frame flag; once a hole has been united with foreground the root it reaches is a
// foreground root, whose aux entry is still 0 (areas come later) -- the test stays right under concurrent unions.
// the rare part of k_rle_holes, kept out of line so that the common loop (no hole anywhere near) stays small
__device__ __noinline__ void rle_join_hole(const uint32_t* __restrict__ M, const int* __restrict__ C, int* Pt,
                                           const RowScan r, uint32_t seg, int k0, int y, int H, int W, int NW,
                                           size_t rowbits, int wj) {
    const int sx = run_start(r, k0);
    const int me = y * W + sx;
    // a hole never touches the frame: it has a left and a right neighbour, a row above and a row below
    const int k1 = 31 - __clz(seg);  // last bit of the segment
    if (sx >= r.wx) rf_union(Pt, me, y * W + (k0 ? run_start(r, k0 - 1) : r.cin));  // the run starts here: left
    if (k1 < 31) {
        if ((r.valid >> (k1 + 1)) & 1u) rf_union(Pt, me, y * W + r.wx + k1 + 1);     // the run ends here: right
    } else if (wj + 1 < NW && (M[rowbits + wj + 1] & 1u)) {
        rf_union(Pt, me, y * W + r.wx + 32);
    }
    if (y > 0) {
        const RowScan up = row_load(M + rowbits - NW, C + rowbits - NW, NW, W, wj);
        for_each_segment(up.m & seg, [&](int j0, uint32_t) { rf_union(Pt, me, (y - 1) * W + run_start(up, j0)); });
    }
    if (y + 1 < H) {
        const RowScan dn = row_load(M + rowbits + NW, C + rowbits + NW, NW, W, wj);
        for_each_segment(dn.m & seg, [&](int j0, uint32_t) { rf_union(Pt, me, (y + 1) * W + run_start(dn, j0)); });
    }
}

__global__ void __launch_bounds__(32 * kRleWarps) k_rle_holes(const uint32_t* __restrict__ M, const int* __restrict__ C,
                                                              int* __restrict__ P, const int* __restrict__ A,
                                                              uint32_t* __restrict__ F, int H, int W) {
    RLE_ROW_COORDS
    int* Pt = P + kNS * tile;
    for (int wj = lane; wj < NW; wj += 32) {
        const RowScan r = row_load(M + rowbits, C + rowbits, NW, W, wj);
        uint32_t hole = 0;
        // common case first: which background segments of the word belong to a component without a frame flag?  The
        // first four segments are looked up together (parent, its parent, flag: three rounds of independent loads)
        uint32_t rest = ~r.m & r.valid;
        uint32_t seg4[4];
        int par[4], root[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            seg4[i] = 0;
            par[i] = -1;
            if (rest) {
                const int k0 = __ffs(rest) - 1;
                const uint32_t inv = ~(rest >> k0);
                const int len = inv ? (__ffs(inv) - 1) : (32 - k0);
                seg4[i] = (len >= 32 ? kFull : ((1u << len) - 1u)) << k0;
                rest &= ~seg4[i];
                par[i] = Pt[kNS * (y * W + run_start(r, k0))];
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) root[i] = par[i] >= 0 ? Pt[kNS * par[i]] : kOutside;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (root[i] >= 0 && root[i] != par[i]) root[i] = rf_find(Pt, root[i]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (seg4[i] && root[i] >= 0) hole |= seg4[i];  // its component never reached kOutside
        for_each_segment(rest, [&](int k0, uint32_t seg) {  // more than four background segments in one word
            if (rf_find(Pt, y * W + run_start(r, k0)) >= 0) hole |= seg;
        });
        F[rowbits + wj] = r.m | hole;
        if (hole) for_each_segment(hole, [&](int k0, uint32_t seg) { rle_join_hole(M, C, Pt, r, seg, k0, y, H, W, NW, rowbits, wj); });
    }
}

__global__ void __launch_bounds__(32 * kRleWarps) k_rle_area(const uint32_t* __restrict__ M, const int* __restrict__ C,
                                                             const uint32_t* __restrict__ F, int* __restrict__ P,
                                                             int* __restrict__ A, int H, int W) {
    RLE_ROW_COORDS
    int* Pt = P + kNS * tile;
    for (int wj = lane; wj < NW; wj += 32) {
        const uint32_t f = F[rowbits + wj];
        if (!f) continue;
        const RowScan r = row_load(M + rowbits, C + rowbits, NW, W, wj);
        // one atomicAdd per filled segment of the word: a foreground run next to a hole run is one component already
        for_each_segment(f, [&](int k0, uint32_t seg) {
            const int sx = run_start(r, k0);
            const int s = y * W + sx;
            const int root = rf_find(Pt, s);
            if (sx >= r.wx && root != s) Pt[kNS * s] = root;  // flatten the starts that live in this word
            atomicAdd(A + kNS * (tile + root), __popc(seg));
        });
    }
}

__global__ void __launch_bounds__(32 * kRleWarps) k_rle_diag(const uint32_t* __restrict__ M, const int* __restrict__ C,
                                                             const uint32_t* __restrict__ F, int* __restrict__ P,
                                                             const int* __restrict__ A, int min_area, int H, int W) {
    RLE_ROW_COORDS
    if (y == 0) return;
    int* Pt = P + kNS * tile;
    const int* At = A + kNS * tile;
    for (int wj = lane; wj < NW; wj += 32) {
        const uint32_t fc = F[rowbits + wj], fp = F[rowbits - NW + wj];
        const uint32_t fcl = left_bits(F + rowbits, NW, wj, fc), fpl = left_bits(F + rowbits - NW, NW, wj, fp);
        // (y, x) and (y-1, x-1) filled, (y, x-1) and (y-1, x) not: only the diagonal joins them
        uint32_t dl = fc & fpl & ~fcl & ~fp;
        // (y-1, x) and (y, x-1) filled, (y-1, x-1) and (y, x) not: the other diagonal, seen from the upper pixel
        uint32_t dr = fp & fcl & ~fpl & ~fc;
        if (wj == 0) { dl &= ~1u; dr &= ~1u; }
        if (!(dl | dr)) continue;
        const RowScan cur = row_load(M + rowbits, C + rowbits, NW, W, wj);
        const RowScan prv = row_load(M + rowbits - NW, C + rowbits - NW, NW, W, wj);
        while (dl) {
            const int k = __ffs(dl) - 1;
            dl &= dl - 1;
            const int ra = rf_find_ro(Pt, y * W + run_start(cur, k));
            const int rb = rf_find_ro(Pt, (y - 1) * W + (k ? run_start(prv, k - 1) : prv.cin));
            if (ra != rb && At[kNS * ra] >= min_area && At[kNS * rb] >= min_area) rf_union(Pt, ra, rb);
        }
        while (dr) {
            const int k = __ffs(dr) - 1;
            dr &= dr - 1;
            const int ra = rf_find_ro(Pt, (y - 1) * W + run_start(prv, k));
            const int rb = rf_find_ro(Pt, y * W + (k ? run_start(cur, k - 1) : cur.cin));
            if (ra != rb && At[kNS * ra] >= min_area && At[kNS * rb] >= min_area) rf_union(Pt, ra, rb);
        }
    }
}

// surviving roots of a row: foreground run starts that are their own parent and whose component is large enough
// The counting pass leaves the roots it found (and the removed ones) as two bit-planes, so that the assigning pass does
// not repeat the sparse parent / area loads.
template <bool ASSIGN>
__global__ void __launch_bounds__(32 * kRleWarps) k_rle_number(const uint32_t* __restrict__ M, const int* __restrict__ P,
                                                               int* __restrict__ A, int* __restrict__ rowcnt,
                                                               uint32_t* __restrict__ RB, uint32_t* __restrict__ DB,
                                                               int min_area, int H, int W) {
    RLE_ROW_COORDS
    const int* Pt = P + kNS * tile;
    int* At = A + kNS * tile;
    int running = 0;
    if (ASSIGN) {
        // id base of this row = surviving roots of the rows above (the per-row counts of a tile are a few KB in L2:
        // summing them here is cheaper than a separate scan launch)
        const int* rc = rowcnt + (size_t)b * H;
        int sum = 0;
        for (int i = lane; i < y; i += 32) sum += rc[i];
        running = __reduce_add_sync(kFull, sum) + 1;
    }
    for (int ch = 0; ch < nchunks; ++ch) {
        const int wj = ch * 32 + lane;
        const bool in = wj < NW;
        uint32_t roots = 0, dead = 0;
        if (ASSIGN) {
            roots = in ? RB[rowbits + wj] : 0u;
            dead = in ? DB[rowbits + wj] : 0u;
        }
        const RowScan r = ASSIGN ? row_word(0u, 0u, -1, W, wj)
                                 : row_word(in ? M[rowbits + wj] : 0u, (in && wj > 0) ? M[rowbits + wj - 1] : 0u, -1, W, wj);
        uint32_t s = ASSIGN ? 0u : (r.t & r.m);
        while (s) {
            // four run starts per round: their parent and area loads are independent of each other
            int kk[4], pv[4], av[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                kk[i] = -1;
                if (s) { kk[i] = __ffs(s) - 1; s &= s - 1; }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int gid = y * W + r.wx + max(kk[i], 0);
                const int2 node = kk[i] >= 0 ? *(const int2*)(Pt + kNS * gid) : make_int2(-1, 0);
                pv[i] = node.x;
                av[i] = node.y;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (kk[i] >= 0 && pv[i] == y * W + r.wx + kk[i]) {
                    if (av[i] >= min_area) roots |= 1u << kk[i];
                    else dead |= 1u << kk[i];
                }
            }
        }
        if (!ASSIGN && in) {
            RB[rowbits + wj] = roots;
            DB[rowbits + wj] = dead;
        }
        const int n = __popc(roots);
        int incl = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += v;
        }
        if (ASSIGN) {
            int id = running + incl - n;
            while (roots) {
                const int k = __ffs(roots) - 1;
                roots &= roots - 1;
                At[kNS * (y * W + r.wx + k)] = id++;
            }
            while (dead) {  // removed components label their pixels 0
                const int k = __ffs(dead) - 1;
                dead &= dead - 1;
                At[kNS * (y * W + r.wx + k)] = 0;
            }
        }
        running += __shfl_sync(kFull, incl, 31);
    }
    if (!ASSIGN && lane == 0) rowcnt[(size_t)b * H + y] = running;
}

// ---- labels + dilation ------------------------------------------------------------------------------------------
// A block owns TR output rows of one column chunk.  Phase 1: warp per staged row, lane per pixel of a 32-pixel
// word: label = id of the root of the pixel's run (F bit set) or 0; the row goes to shared memory (4 zero columns
// of padding left and right).  Phase 2: max over disk(R) from shared memory, 4 pixels per thread, written as OUT.
#ifndef CDNET_LAB_MINB
#define CDNET_LAB_MINB 4  // resident blocks per SM the register allocation aims at
#endif
#ifndef CDNET_LAB_ROWS
#define CDNET_LAB_ROWS 4
#endif
constexpr int kLabRows = CDNET_LAB_ROWS;
constexpr int kLabPad = 4;
constexpr int kLabTW = 1024 + 2 * kLabPad;

template <int R, typename OUT>
__global__ void __launch_bounds__(256, CDNET_LAB_MINB) k_rle_labels(const uint32_t* __restrict__ M, const int* __restrict__ C,
                                                    const uint32_t* __restrict__ F, const int* __restrict__ P,
                                                    const int* __restrict__ A, OUT* __restrict__ out, int H, int W,
                                                    int chunk_px, int halo) {
    pdl_wait();
    pdl_trigger();
    CDNET_DYN_SHARED(int, s_lab);  // [(kLabRows + 2 R)][kLabTW]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int b = blockIdx.z;
    const int y0 = blockIdx.y * kLabRows;
    const int NW = (W + 31) >> 5;
    const int w0 = blockIdx.x * (chunk_px >> 5) - halo;  // first word of the staged window (may be -1)
    const size_t tile = (size_t)b * H * W;
    const int* Pt = P + kNS * tile;
    const int* At = A + kNS * tile;
    constexpr int SR = kLabRows + 2 * R;
    // zero padding columns
    for (int i = threadIdx.x; i < SR * 2 * kLabPad; i += 256) {
        const int rr = i / (2 * kLabPad), c = i % (2 * kLabPad);
        s_lab[rr * kLabTW + (c < kLabPad ? c : 1024 + c)] = 0;
    }
    for (int rr = wid; rr < SR; rr += 8) {
        const int y = y0 - R + rr;
        int* dst = s_lab + rr * kLabTW + kLabPad;
        if (y < 0 || y >= H) {
#pragma unroll
            for (int q = 0; q < 8; ++q) *(int4*)(dst + 32 * lane + 4 * ((q + lane) & 7)) = make_int4(0, 0, 0, 0);
            continue;
        }
        const size_t rowbits = ((size_t)b * H + y) * NW;
        const int wj = w0 + lane;
        const bool in = wj >= 0 && wj < NW;
        const RowScan r = in ? row_load(M + rowbits, C + rowbits, NW, W, wj) : row_word(0u, 0u, -1, W, NW);
        const uint32_t f = in ? F[rowbits + wj] : 0u;
        // lane = word: the labels of (up to four) filled segments of the word are looked up with every lane's loads in
        // flight together; a filled segment is one component (a foreground run and a hole run side by side have
        // been united).  Then the 32 pixels of the word go to shared memory, skewed by the lane so that the 32 lanes
        // hit 32 different banks.
        uint32_t smask[4];
        int slab[4], sidx[4];
        uint32_t rest = f;
#pragma unroll
        for (int sgi = 0; sgi < 4; ++sgi) {
            smask[sgi] = 0;
            sidx[sgi] = -1;
            if (rest) {
                const int k0 = __ffs(rest) - 1;
                const uint32_t inv = ~(rest >> k0);
                const int len = inv ? (__ffs(inv) - 1) : (32 - k0);
                const uint32_t seg = (len >= 32 ? kFull : ((1u << len) - 1u)) << k0;
                smask[sgi] = seg;
                rest &= ~seg;
                sidx[sgi] = y * W + run_start(r, k0);
            }
        }
        // the look-ups of the four segments are issued together: parent of the run start (the area pass left it pointing
        // at its root, the diagonal pass may have re-rooted that) -> root -> id.  Three rounds of independent loads
        // instead of up to twelve dependent ones.
        int par[4];
        int2 node[4];
#pragma unroll
        for (int sgi = 0; sgi < 4; ++sgi) par[sgi] = sidx[sgi] >= 0 ? Pt[kNS * sidx[sgi]] : -1;
        // parent and id sit side by side: when the parent is the root (the common case) its node holds the answer
#pragma unroll
        for (int sgi = 0; sgi < 4; ++sgi) node[sgi] = par[sgi] >= 0 ? *(const int2*)(Pt + kNS * par[sgi]) : make_int2(-1, 0);
#pragma unroll
        for (int sgi = 0; sgi < 4; ++sgi) {
            slab[sgi] = node[sgi].y;
            if (par[sgi] >= 0 && node[sgi].x != par[sgi]) slab[sgi] = At[kNS * rf_find_ro(Pt, node[sgi].x)];  // rare: a longer chain
        }
        // the row is zero-filled with 128-bit stores (skewed by the lane: each quarter-warp covers all 32 banks), then
        // only the pixels of labelled segments are written -- a quarter of a tile is nucleus, not all of it
#pragma unroll
        for (int q = 0; q < 8; ++q) *(int4*)(dst + 32 * lane + 4 * ((q + lane) & 7)) = make_int4(0, 0, 0, 0);
        __syncwarp();
#pragma unroll
        for (int sgi = 0; sgi < 4; ++sgi) {
            if (slab[sgi]) {
                const int k0 = __ffs(smask[sgi]) - 1, len = __popc(smask[sgi]);
                for (int k = 0; k < len; ++k) dst[32 * lane + k0 + k] = slab[sgi];
            }
        }
        while (rest) {  // more than four segments in one word: rare
            const int pbit = __ffs(rest) - 1;
            rest &= rest - 1;
            dst[32 * lane + pbit] = At[kNS * rf_find_ro(Pt, y * W + run_start(r, pbit))];
        }
    }
    __syncthreads();
    // ---- phase 2
    const int xbase = w0 * 32;                       // x of staged column 0
    const int own_lo = (w0 + halo) * 32;             // first owned x
    const int own_hi = min(W, own_lo + chunk_px);    // one past the last owned x
    const int rows = min(kLabRows, H - y0);
    OUT* ot = out + tile;
    if ((W & 3) == 0 && (((uintptr_t)out) & 15) == 0) {
        // a thread owns a 4-pixel column of the block and walks down the staged rows ONCE: per staged row the centre
        // values, their 3-wide and their 5-wide horizontal maxima are formed once and serve every output row whose
        // disk touches that row (disk(2): rows y-2 / y+2 contribute their centre, y-1 / y+1 three pixels, y five)
        const int nq = (own_hi - own_lo) >> 2;
        for (int q = threadIdx.x; q < nq; q += 256) {
            const int x4 = own_lo + 4 * q;
            const int col = x4 - xbase + kLabPad;  // staged column of pixel x4 (multiple of 4)
            int a1[SR][4], a3[SR][4], a5[SR][4];
#pragma unroll
            for (int sr = 0; sr < SR; ++sr) {
                const int* rowp = s_lab + sr * kLabTW + col;
                const int4 cv = *(const int4*)rowp;
                a1[sr][0] = cv.x; a1[sr][1] = cv.y; a1[sr][2] = cv.z; a1[sr][3] = cv.w;
                if (R >= 1) {
                    const int2 l = *(const int2*)(rowp - 2), rg = *(const int2*)(rowp + 4);
                    a3[sr][0] = max(max(l.y, cv.x), cv.y);
                    a3[sr][1] = max(max(cv.x, cv.y), cv.z);
                    a3[sr][2] = max(max(cv.y, cv.z), cv.w);
                    a3[sr][3] = max(max(cv.z, cv.w), rg.x);
                    if (R >= 2) {
                        a5[sr][0] = max(a3[sr][0], max(l.x, cv.z));
                        a5[sr][1] = max(a3[sr][1], max(l.y, cv.w));
                        a5[sr][2] = max(a3[sr][2], max(cv.x, rg.x));
                        a5[sr][3] = max(a3[sr][3], max(cv.y, rg.y));
                    }
                }
                if (sr >= 2 * R) {
                    const int ry = sr - 2 * R;  // output row whose disk ends at staged row sr
                    if (ry < rows) {
                        int m[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (R == 0) m[j] = a1[ry][j];
                            else if (R == 1) m[j] = max(a3[ry + 1][j], max(a1[ry][j], a1[ry + 2][j]));
                            else m[j] = max(max(a5[ry + 2][j], max(a3[ry + 1][j], a3[ry + 3][j])), max(a1[ry][j], a1[ry + 4][j]));
                        }
                        OUT* o = ot + (size_t)(y0 + ry) * W + x4;
                        if (sizeof(OUT) == 4) {
                            *(int4*)o = make_int4(m[0], m[1], m[2], m[3]);
                        } else {
                            *(longlong2*)o = make_longlong2((long long)m[0], (long long)m[1]);
                            *(longlong2*)(o + 2) = make_longlong2((long long)m[2], (long long)m[3]);
                        }
                    }
                }
            }
        }
    } else {
        const int nx = own_hi - own_lo;
        for (int i = threadIdx.x; i < rows * nx; i += 256) {
            const int ry = i / nx, x = own_lo + (i - ry * nx);
            const int col = x - xbase + kLabPad;
            int m = 0;
#pragma unroll
            for (int dy = -R; dy <= R; ++dy)
#pragma unroll
                for (int dx = -R; dx <= R; ++dx)
                    if (dx * dx + dy * dy <= R * R) m = max(m, s_lab[(ry + R + dy) * kLabTW + col + dx]);
            ot[(size_t)(y0 + ry) * W + x] = (OUT)m;
        }
    }
}

template <int R, typename OUT>
static int rle_labels_launch(const uint32_t* M, const int* C, const uint32_t* F, const int* P, const int* A, OUT* out,
                             int B, int H, int W, cudaStream_t st) {
    // one chunk of 1024 columns when the tile fits; otherwise chunks of 32 words whose first and last word are halo
    // (30 owned words): the window then starts at word 30 * blockIdx.x - 1
    const int halo = W > 1024 ? 1 : 0;
    const int chunk_px = halo ? 960 : 1024;
    const size_t smem = (size_t)(kLabRows + 2 * R) * kLabTW * sizeof(int);
    static bool attr_done = false;
    if (!attr_done) {
        CDNET_CUDA_OK(cudaFuncSetAttribute(k_rle_labels<R, OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        attr_done = true;
    }
    dim3 grid(ceil_div(W, chunk_px), ceil_div(H, kLabRows), B);
    CDNET_LAUNCH_PDL((k_rle_labels<R, OUT>), grid, 256, smem, st, M, C, F, P, A, out, H, W, chunk_px, halo);
    return last_error();
}

size_t rle_tail_workspace(int B, int H, int W) {
    const size_t n = (size_t)B * H * W;
    const size_t nbits = (size_t)B * H * ((W + 31) / 32) * sizeof(uint32_t);
    return 2 * pad256(n * 4) + 5 * pad256(nbits) + pad256((size_t)B * H * 4);
}

bool rle_tail_supported(int radius) {
    static int off = -1;
    if (off < 0) off = getenv("CDNET_NO_RLE") ? 1 : 0;
    return !off && radius >= 0 && radius <= 2;
}

// inside (0 / non-zero bytes) -> out = dilation(label8(remove_small(fill_holes(inside), min_area)), disk(radius))
// label4: plain 4-connected labelling of the mask (no hole filling, no size filter, no diagonal joins, no dilation) with
// the same raster-first numbering -- what process() needs for its markers (postproc_other.py:44)
// markers: fill_holes -> cross erosion -> label4 (postproc_other.py:42-44) as two chains that meet in the bit domain
enum { kChainTail = 0, kChainLabel4 = 1, kChainMarkers = 2 };
static bool rle_fused_path(int W) {
    static int fused = -1;  // CDNET_RLE_NO_LOCAL=1: every link through global memory
    if (fused < 0) fused = getenv("CDNET_RLE_NO_LOCAL") ? 0 : 1;
    return fused && W <= 1024;
}
static int rle_chain(const uint8_t* inside, void* out, int out_elem_bytes, int B, int H, int W, int min_area, int radius,
                     int mode, void* ws, size_t ws_bytes, cudaStream_t st) {
    const size_t n = (size_t)B * H * W;
    const size_t nbits = (size_t)B * H * ((W + 31) / 32);
    Arena ar(ws, ws_bytes);
    int* P = ar.take<int>(kNS * n);  // nodes: parent at 2 i, aux at 2 i + 1
    int* A = P + 1;
    uint32_t* M = ar.take<uint32_t>(nbits);
    uint32_t* F = ar.take<uint32_t>(nbits);
    int* C = ar.take<int>(nbits);
    uint32_t* RB = ar.take<uint32_t>(nbits);
    uint32_t* DB = ar.take<uint32_t>(nbits);
    int* rowcnt = ar.take<int>((size_t)B * H);
    if (!ar.ok) return CDNET_E_WORKSPACE;
    const dim3 grid = rle_grid(B, H);
    const int threads = 32 * kRleWarps;
    CDNET_RANGE("run-based tail (rle.cu)");
    if (mode == kChainMarkers && !rle_fused_path(W)) return CDNET_E_BADARG;  // callers ask rle_markers_supported first
    const uint8_t* src = inside;
    bool eroded = false;
second_chain:
    if (rle_fused_path(W)) {
        // rows per block of the shared-memory union-find (CDNET_RLE_PACK_ROWS = 8 | 16 | 32); the seams between the
        // blocks are what k_rle_link joins through global memory afterwards
        // Default: 16 rows when that still gives every SM two blocks, else 8 (measured on 14 x 1000^2, whole step: 16 rows
        // 0.274 ms, 32 rows 0.278 ms, 8 rows 0.277 ms; a single 1000^2 tile keeps 8 rows = 125 blocks).
        static int prow_env = -1;
        if (prow_env < 0) {
            const char* e = getenv("CDNET_RLE_PACK_ROWS");
            prow_env = e ? atoi(e) : 0;
            if (prow_env != 8 && prow_env != 16 && prow_env != 32) prow_env = 0;
            CDNET_CUDA_OK(cudaFuncSetAttribute(k_rle_pack_link<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pack_smem(16)));
            CDNET_CUDA_OK(cudaFuncSetAttribute(k_rle_pack_link<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pack_smem(32)));
            CDNET_CUDA_OK(cudaFuncSetAttribute(k_rle_pack_link<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pack_smem(16)));
            CDNET_CUDA_OK(cudaFuncSetAttribute(k_rle_pack_link<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pack_smem(32)));
        }
        const long long rows_total = (long long)B * H;
        const int prow = prow_env ? prow_env : (rows_total >= 16LL * 296 ? 16 : 8);
        const dim3 pgrid(ceil_div(H, prow), B);
        const size_t psm = pack_smem(prow);
        if (eroded) {
            if (prow == 8) CDNET_LAUNCH_PDL((k_rle_pack_link<8, true>), pgrid, 256, psm, st, src, M, C, P, A, H, W);
            else if (prow == 16) CDNET_LAUNCH_PDL((k_rle_pack_link<16, true>), pgrid, 512, psm, st, src, M, C, P, A, H, W);
            else CDNET_LAUNCH_PDL((k_rle_pack_link<32, true>), pgrid, 1024, psm, st, src, M, C, P, A, H, W);
        } else {
            if (prow == 8) CDNET_LAUNCH_PDL((k_rle_pack_link<8, false>), pgrid, 256, psm, st, src, M, C, P, A, H, W);
            else if (prow == 16) CDNET_LAUNCH_PDL((k_rle_pack_link<16, false>), pgrid, 512, psm, st, src, M, C, P, A, H, W);
            else CDNET_LAUNCH_PDL((k_rle_pack_link<32, false>), pgrid, 1024, psm, st, src, M, C, P, A, H, W);
        }
        // the seams between the groups in one launch (CDNET_RLE_SEAM_PHASES=2 links the seams inside super-groups of 64
        // rows first: measured slower on B200, 0.057 vs 0.042 ms for 14 x 1000^2)
        static int seam2 = -1;
        if (seam2 < 0) { const char* e = getenv("CDNET_RLE_SEAM_PHASES"); seam2 = (e && atoi(e) == 2) ? 1 : 0; }
        if (H > prow) {
            if (seam2 && H > 64 && prow < 64) {
                CDNET_LAUNCH_PDL(k_rle_link, link_grid(B, H, prow), 32 * kLinkWarps, 0, st, M, C, P, H, W, prow, 64);
                CDNET_LAUNCH_PDL(k_rle_link, link_grid(B, H, 64), 32 * kLinkWarps, 0, st, M, C, P, H, W, 64, 0);
            } else {
                CDNET_LAUNCH_PDL(k_rle_link, link_grid(B, H, prow), 32 * kLinkWarps, 0, st, M, C, P, H, W, prow, 0);
            }
        }
    } else {
        CDNET_LAUNCH_PDL(k_rle_pack, grid, threads, 0, st, src, M, C, P, A, H, W);
        static int phases = 0;  // CDNET_RLE_LINK_PHASES=3: rows inside groups of 8, then of 64, then the rest (slower on B200)
        if (!phases) { const char* e = getenv("CDNET_RLE_LINK_PHASES"); phases = (e && atoi(e) == 3) ? 3 : 1; }
        if (phases == 1 || H <= 8) {
            CDNET_LAUNCH_PDL(k_rle_link, link_grid(B, H, 1), 32 * kLinkWarps, 0, st, M, C, P, H, W, 1, 0);
        } else {
            CDNET_LAUNCH_PDL(k_rle_link, link_grid(B, H, 1), 32 * kLinkWarps, 0, st, M, C, P, H, W, 1, 8);
            CDNET_LAUNCH_PDL(k_rle_link, link_grid(B, H, 8), 32 * kLinkWarps, 0, st, M, C, P, H, W, 8, 64);
            if (H > 64) CDNET_LAUNCH_PDL(k_rle_link, link_grid(B, H, 64), 32 * kLinkWarps, 0, st, M, C, P, H, W, 64, 0);
        }
    }
    if (mode == kChainMarkers && !eroded) {
        // first chain done up to the filled plane F; the second chain reads it back through the erosion
        CDNET_LAUNCH_PDL(k_rle_holes, grid, threads, 0, st, M, C, P, A, F, H, W);
        src = reinterpret_cast<const uint8_t*>(F);
        eroded = true;
        goto second_chain;
    }
    if (mode != kChainTail) {
        // every foreground root survives (aux is still 0 >= 0); the filled plane is the mask itself
        CDNET_LAUNCH_PDL(k_rle_number<false>, grid, threads, 0, st, M, P, A, rowcnt, RB, DB, 0, H, W);
        CDNET_LAUNCH_PDL(k_rle_number<true>, grid, threads, 0, st, M, P, A, rowcnt, RB, DB, 0, H, W);
        return rle_labels_launch<0, int32_t>(M, C, M, P, A, (int32_t*)out, B, H, W, st);
    }
    CDNET_LAUNCH_PDL(k_rle_holes, grid, threads, 0, st, M, C, P, A, F, H, W);
    CDNET_LAUNCH_PDL(k_rle_area, grid, threads, 0, st, M, C, F, P, A, H, W);
    CDNET_LAUNCH_PDL(k_rle_diag, grid, threads, 0, st, M, C, F, P, A, min_area, H, W);
    CDNET_LAUNCH_PDL(k_rle_number<false>, grid, threads, 0, st, M, P, A, rowcnt, RB, DB, min_area, H, W);
    CDNET_LAUNCH_PDL(k_rle_number<true>, grid, threads, 0, st, M, P, A, rowcnt, RB, DB, min_area, H, W);
    if (out_elem_bytes == 4) {
        if (radius == 0) return rle_labels_launch<0, int32_t>(M, C, F, P, A, (int32_t*)out, B, H, W, st);
        if (radius == 1) return rle_labels_launch<1, int32_t>(M, C, F, P, A, (int32_t*)out, B, H, W, st);
        return rle_labels_launch<2, int32_t>(M, C, F, P, A, (int32_t*)out, B, H, W, st);
    }
    if (radius == 0) return rle_labels_launch<0, long long>(M, C, F, P, A, (long long*)out, B, H, W, st);
    if (radius == 1) return rle_labels_launch<1, long long>(M, C, F, P, A, (long long*)out, B, H, W, st);
    return rle_labels_launch<2, long long>(M, C, F, P, A, (long long*)out, B, H, W, st);
}

int rle_tail_launch(const uint8_t* inside, void* out, int out_elem_bytes, int B, int H, int W, int min_area, int radius,
                    void* ws, size_t ws_bytes, cudaStream_t st) {
    return rle_chain(inside, out, out_elem_bytes, B, H, W, min_area, radius, kChainTail, ws, ws_bytes, st);
}

int rle_label4_launch(const uint8_t* mask, int32_t* labels, int B, int H, int W, void* ws, size_t ws_bytes, cudaStream_t st) {
    return rle_chain(mask, labels, 4, B, H, W, 0, 0, kChainLabel4, ws, ws_bytes, st);
}

bool rle_markers_supported(int W) { return rle_tail_supported(0) && rle_fused_path(W); }

int rle_markers_launch(const uint8_t* marker0, int32_t* labels, int B, int H, int W, void* ws, size_t ws_bytes, cudaStream_t st) {
    return rle_chain(marker0, labels, 4, B, H, W, 0, 0, kChainMarkers, ws, ws_bytes, st);
}

}  // namespace cdnet
