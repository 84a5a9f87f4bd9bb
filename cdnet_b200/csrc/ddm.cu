// ddm.cu -- direction-difference map (K1/K2 of SURVEY.md section 2.1).
//
// Replaces generate_dd_map / circshift / label_to_vector
// (data_prepare/getDirectionDiffMap.py:14-108, data_prepare/SegFix_offset_helper.py:50-89,246-261).
//
// The reference builds int64 vector planes, eight shifted copies, f64 cosines stored to f32, a
// channel minimum, np.around and a per-image min-max normalisation.  Exact reformulation
// (SURVEY.md Appendix B.1): around(f32(cos(v_a, v_b))) in {-1,0,1} depends only on the two class
// ids, and around is monotone, so d = 1 - min_k LUT[a][b_k] in {0,1,2}.  The LUT is evaluated on
// the host with the reference's own arithmetic and shipped as two bit-sets per class (pos / neg);
// a pixel only needs the SET S of classes present in its neighbourhood:
//     any b in S with LUT[a][b] = -1  -> d = 2 ;  all b in S have LUT = +1 -> d = 0 ;  else d = 1
// Zero padding == class 0 (zero vector, cos 0).  Output of the first pass is a 2-bit code per
// (pixel, map) plus three "value present" bits per map for the normalisation.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace cdnet {

struct DdmLut {
    uint32_t pos[32];
    uint32_t neg[32];
    int n;           // number of classes incl. background (5, 9, 17)
    int axial;       // 5 classes: only the 4 axial neighbours (getDirectionDiffMap.py:56-67)
    int force_zero;  // 17 classes: 8 of 16 cosine channels stay 0 (getDirectionDiffMap.py:90 vs 69-88)
};

static const int kRing8[8][2] = {{0, -1}, {-1, -1}, {-1, 0}, {-1, 1}, {0, 1}, {1, 1}, {1, 0}, {1, -1}};
static const int kRing16[16][2] = {{0, -2}, {-1, -2}, {-2, -2}, {-2, -1}, {-2, 0}, {-2, 1}, {-2, 2}, {-1, 2},
                                   {0, 2},  {1, 2},   {2, 2},   {2, 1},   {2, 0},  {2, -1}, {2, -2}, {1, -2}};
static const int kDiag4[4][2] = {{-1, -1}, {-1, 1}, {1, 1}, {1, -1}};

// class -> (dh, dw), data_prepare/SegFix_offset_helper.py:50-89 (keys 5, 9, 17)
bool ddm_build_lut(int n_classes, DdmLut* lut) {
    int vec[32][2] = {{0, 0}};
    if (n_classes == 5) {
        for (int i = 0; i < 4; ++i) { vec[i + 1][0] = kDiag4[i][0]; vec[i + 1][1] = kDiag4[i][1]; }
    } else if (n_classes == 9) {
        for (int i = 0; i < 8; ++i) { vec[i + 1][0] = kRing8[i][0]; vec[i + 1][1] = kRing8[i][1]; }
    } else if (n_classes == 17) {
        for (int i = 0; i < 16; ++i) { vec[i + 1][0] = kRing16[i][0]; vec[i + 1][1] = kRing16[i][1]; }
    } else {
        return false;
    }
    lut->n = n_classes;
    lut->axial = (n_classes == 5);
    lut->force_zero = (n_classes == 17);
    for (int a = 0; a < 32; ++a) { lut->pos[a] = 0; lut->neg[a] = 0; }
    for (int a = 0; a < n_classes; ++a) {
        for (int b = 0; b < n_classes; ++b) {
            // getDirectionDiffMap.py:92-97: int64 dot, f64 norms, +1e-6, f64 divide, store to f32
            const double num = (double)(vec[a][0] * vec[b][0] + vec[a][1] * vec[b][1]);
            const double den = sqrt((double)(vec[a][0] * vec[a][0] + vec[a][1] * vec[a][1])) *
                                   sqrt((double)(vec[b][0] * vec[b][0] + vec[b][1] * vec[b][1])) +
                               0.000001;
            const float c = (float)(num / den);
            const float r = nearbyintf(c);  // np.around: half to even (default rounding mode)
            if (r > 0.5f) lut->pos[a] |= (1u << b);
            if (r < -0.5f) lut->neg[a] |= (1u << b);
        }
    }
    return true;
}

constexpr int kRows = 8;  // rows per thread strip

template <bool FAST>
__device__ __forceinline__ void load_row6(const uint8_t* __restrict__ plane, int H, int W, int y, int x4,
                                          int n, uint32_t cb[6], uint32_t cls[4]) {
    // class bits of columns x4-1 .. x4+4 of row y (class 0 outside the image / for invalid ids)
    uint32_t c[6] = {0, 0, 0, 0, 0, 0};
    if (y >= 0 && y < H) {
        const uint8_t* row = plane + (size_t)y * W;
        if (FAST && x4 + 3 < W) {
            const uint32_t w = __ldg((const uint32_t*)(row + x4));
            c[1] = w & 0xff; c[2] = (w >> 8) & 0xff; c[3] = (w >> 16) & 0xff; c[4] = w >> 24;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (x4 + i < W) c[1 + i] = __ldg(row + x4 + i);
        }
        if (x4 > 0) c[0] = __ldg(row + x4 - 1);
        if (x4 + 4 < W) c[5] = __ldg(row + x4 + 4);
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const uint32_t ci = c[i] < (uint32_t)n ? c[i] : 0u;  // unknown id: zero vector, like class 0
        cb[i] = 1u << ci;
        if (i >= 1 && i <= 4) cls[i - 1] = c[i];
    }
}

// codes: uint16 [B,H,W], bits 2t..2t+1 = d of map t.  flags: uint32 [B], bit 3t+d = "map t has value d".
template <int T, bool FAST>
__global__ void __launch_bounds__(128) k_ddm_codes(const uint8_t* __restrict__ cls_maps, uint16_t* __restrict__ codes,
                                                   uint32_t* __restrict__ flags, int H, int W, DdmLut lut, int row_lo,
                                                   int row_hi) {
    __shared__ uint32_t s_pos[32], s_neg[32];
    const int tid = threadIdx.y * 32 + threadIdx.x;
    if (tid < 32) { s_pos[tid] = lut.pos[tid]; s_neg[tid] = lut.neg[tid]; }
    __syncthreads();
    const int b = blockIdx.z;
    const int x4 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int y0 = (blockIdx.y * blockDim.y + threadIdx.y) * kRows;
    uint32_t seen = 0;
    if (x4 < W && y0 < H) {
        uint32_t acc[kRows][2];  // 4 x uint16 per row
#pragma unroll
        for (int r = 0; r < kRows; ++r) { acc[r][0] = 0; acc[r][1] = 0; }
        const size_t plane_sz = (size_t)H * W;
#pragma unroll 1
        for (int t = 0; t < T; ++t) {
            const uint8_t* plane = cls_maps + ((size_t)b * T + t) * plane_sz;
            uint32_t cb_prev[6], cb_cur[6], cb_next[6], cls_prev[4], cls_cur[4], cls_next[4];
            load_row6<FAST>(plane, H, W, y0 - 1, x4, lut.n, cb_prev, cls_prev);
            load_row6<FAST>(plane, H, W, y0, x4, lut.n, cb_cur, cls_cur);
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                const int y = y0 + r;
                load_row6<FAST>(plane, H, W, y + 1, x4, lut.n, cb_next, cls_next);
                if (y < H) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint32_t S;
                        if (lut.axial) {
                            S = cb_prev[i + 1] | cb_next[i + 1] | cb_cur[i] | cb_cur[i + 2];
                        } else {
                            S = cb_prev[i] | cb_prev[i + 1] | cb_prev[i + 2] | cb_cur[i] | cb_cur[i + 2] |
                                cb_next[i] | cb_next[i + 1] | cb_next[i + 2];
                        }
                        const uint32_t a = cls_cur[i];
                        uint32_t d = 0;
                        if (a != 0) {
                            const uint32_t ai = a < (uint32_t)lut.n ? a : 31u;  // row 31 is empty
                            if (s_neg[ai] & S) d = 2;
                            else if ((S & ~s_pos[ai]) != 0 || lut.force_zero) d = 1;
                        }
                        if (x4 + i < W) {
                            if (y >= row_lo && y < row_hi) seen |= 1u << (3 * t + d);
                            acc[r][i >> 1] |= d << (2 * t + 16 * (i & 1));
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < 6; ++i) { cb_prev[i] = cb_cur[i]; cb_cur[i] = cb_next[i]; }
#pragma unroll
                for (int i = 0; i < 4; ++i) cls_cur[i] = cls_next[i];
            }
        }
        uint16_t* cout = codes + (size_t)b * plane_sz;
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            const int y = y0 + r;
            if (y < H) {
                uint16_t* dst = cout + (size_t)y * W + x4;
                if (FAST && x4 + 3 < W) {
                    *(uint2*)dst = make_uint2(acc[r][0], acc[r][1]);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (x4 + i < W) dst[i] = (uint16_t)(acc[r][i >> 1] >> (16 * (i & 1)));
                }
            }
        }
    }
    seen = __reduce_or_sync(0xffffffffu, seen);
    if (threadIdx.x == 0 && seen) atomicOr(flags + b, seen);
}

// ---- byte-SIMD variant for the 5- and 9-class tables (<= 8 direction classes) ----------------------
// Four pixels travel in one 32-bit register.  Per pixel the neighbourhood class set is ONE BYTE (bit k =
// direction class k+1 present) plus one "background/zero-vector present" bit; class -> one-hot byte,
// class -> neg/pos byte masks are 8-entry byte LUTs evaluated for 4 pixels at once with PRMT
// (__byte_perm), the 3-wide horizontal OR is two funnel shifts + one LOP3, and "byte != 0" tests use
// the carry-free haszero trick.  ~14 thread-instructions per (pixel, map) instead of ~53.
__device__ __forceinline__ uint32_t nonzero7(uint32_t v) {  // bit 7 of each byte set iff the byte != 0
    return (((v & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v) & 0x80808080u;
}

struct RowSets {
    uint32_t onehot;  // per byte: 1 << (class-1) for a valid direction class, else 0
    uint32_t bg;      // per byte: 1 where the pixel acts as the zero vector (class 0, unknown id, outside)
    uint32_t act7;    // bit 7 per byte: valid direction class
    uint32_t odd7;    // bit 7 per byte: non-zero but unknown id (centre code is always 1)
    uint32_t sel;     // PRMT selector nibbles (class-1) & 7
};

__device__ __forceinline__ RowSets classify4(uint32_t w, uint32_t ge_add) {
    RowSets r;
    const uint32_t inv7 = (((w & 0x7f7f7f7fu) + ge_add) | w) & 0x80808080u;  // byte >= n
    const uint32_t nz7 = nonzero7(w);
    r.act7 = nz7 & ~inv7;
    r.odd7 = nz7 & inv7;
    const uint32_t actmask = (r.act7 >> 7) * 0xffu;
    uint32_t km = ((w & 0x0f0f0f0fu) + 0x07070707u) & 0x07070707u;  // (class-1) & 7 per byte
    km = __byte_perm(km, 0u, 0x3120);
    r.sel = km | (km >> 12);
    r.onehot = __byte_perm(0x08040201u, 0x80402010u, r.sel) & actmask;
    r.bg = ((~r.act7) & 0x80808080u) >> 7;
    return r;
}

__device__ __forceinline__ void classify1(uint32_t c, int n, uint32_t& onehot, uint32_t& bg) {
    const bool act = c >= 1u && c < (uint32_t)n;
    onehot = act ? (1u << (c - 1u)) : 0u;
    bg = act ? 0u : 1u;
}

struct RawRow {
    uint32_t w, e;  // 4 class bytes; edge byte: the pixel left of the word (lane 0) / right of it (lane 31)
};

// colmode: 0 = word outside the image, 1 = full word (32-bit load), 2 = ragged (byte loads); edge_off: offset of
// this lane's edge byte relative to x4 (-1 / +4) or 0 for "no edge byte" -- all row-invariant, computed once
template <bool FAST>
__device__ __forceinline__ RawRow load_raw(const uint8_t* __restrict__ plane, int H, int W, int y, int x4, int colmode,
                                           int edge_off) {
    RawRow r;
    r.w = 0; r.e = 0;
    if ((unsigned)y < (unsigned)H && colmode) {
        const uint8_t* row = plane + (size_t)y * W + x4;
        if (FAST && colmode == 1) {
            r.w = __ldg((const uint32_t*)row);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (x4 + i < W) r.w |= (uint32_t)__ldg(row + i) << (8 * i);
        }
        if (edge_off) r.e = __ldg(row + edge_off);
    }
    return r;
}

__device__ __forceinline__ void classify_row(const RawRow& raw, int n, uint32_t ge_add, int lane, RowSets& rs,
                                             uint32_t& h3, uint32_t& hb3, uint32_t& lr, uint32_t& lrb) {
    rs = classify4(raw.w, ge_add);
    uint32_t ohl = __shfl_up_sync(0xffffffffu, rs.onehot, 1), bgl = __shfl_up_sync(0xffffffffu, rs.bg, 1);
    uint32_t ohr = __shfl_down_sync(0xffffffffu, rs.onehot, 1), bgr = __shfl_down_sync(0xffffffffu, rs.bg, 1);
    {
        // warp edges: the neighbour pixel comes from the edge byte instead of the adjacent lane (selects, no branch)
        uint32_t o, g;
        classify1(raw.e, n, o, g);
        ohl = lane == 0 ? (o << 24) : ohl;
        bgl = lane == 0 ? (g << 24) : bgl;
        ohr = lane == 31 ? o : ohr;
        bgr = lane == 31 ? g : bgr;
    }
    const uint32_t L = __funnelshift_l(ohl, rs.onehot, 8), R = __funnelshift_r(rs.onehot, ohr, 8);
    const uint32_t Lb = __funnelshift_l(bgl, rs.bg, 8), Rb = __funnelshift_r(rs.bg, bgr, 8);
    lr = L | R;
    lrb = Lb | Rb;
    h3 = lr | rs.onehot;
    hb3 = lrb | rs.bg;
}

struct DdmLut8 {
    uint32_t neg_lo, neg_hi, pos_lo, pos_hi;  // byte k = neg/pos set of class k+1 over classes 1..8
    int n, axial;
};

#ifndef CDNET_DDM_PB
#define CDNET_DDM_PB 2
#endif
#ifndef CDNET_DDM_MINB
#define CDNET_DDM_MINB 5
#endif
#ifndef CDNET_DDM_ROWS_DEFAULT
#define CDNET_DDM_ROWS_DEFAULT 4
#endif
template <int T, bool FAST, bool AXIAL, int ROWS>
__global__ void __launch_bounds__(128, CDNET_DDM_MINB) k_ddm_codes_simd(const uint8_t* __restrict__ cls_maps, uint16_t* __restrict__ codes,
                                                        uint32_t* __restrict__ flags, int H, int W, DdmLut8 lut,
                                                        int row_lo, int row_hi) {
    const int b = blockIdx.z;
    const int lane = threadIdx.x;
    const int x4 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int y0 = (blockIdx.y * blockDim.y + threadIdx.y) * ROWS;
    uint32_t seen = 0;
    const uint32_t ge_add = (uint32_t)(0x80 - lut.n) * 0x01010101u;
    // pixels of this word that lie inside the image (bit 7 per byte)
    uint32_t inimg7 = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (x4 + i < W) inimg7 |= 0x80u << (8 * i);
    if (y0 >= H) return;  // whole warp shares y0 (one warp per threadIdx.y): uniform exit, nothing to reduce
    const int colmode = x4 >= W ? 0 : (x4 + 3 < W ? 1 : 2);
    const int edge_off = (lane == 0 && x4 > 0 && x4 < W) ? -1 : ((lane == 31 && x4 + 4 < W) ? 4 : 0);
    {
        uint32_t acc[ROWS][2];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) { acc[r][0] = 0; acc[r][1] = 0; }
        const size_t plane_sz = (size_t)H * W;
        // all row words of a plane pair are requested before any of them is consumed: the kernel is bound by
        // load latency, not by issue slots, so memory-level parallelism is what buys time here
        constexpr int PB = (T >= 2) ? CDNET_DDM_PB : 1;
#pragma unroll 1
        for (int t0 = 0; t0 < T; t0 += PB) {
            RawRow raw[PB][ROWS + 2];
#pragma unroll
            for (int u = 0; u < PB; ++u) {
                const uint8_t* plane = cls_maps + ((size_t)b * T + t0 + u) * plane_sz;
#pragma unroll
                for (int r = 0; r < ROWS + 2; ++r) raw[u][r] = load_raw<FAST>(plane, H, W, y0 - 1 + r, x4, colmode, edge_off);
            }
#pragma unroll
            for (int u = 0; u < PB; ++u) {
                const int t = t0 + u;
                RowSets rp, rc, rn;
                uint32_t h3p, hb3p, lrp, lrbp, h3c, hb3c, lrc, lrbc, h3n, hb3n, lrn, lrbn;
                uint32_t m0 = 0, m1 = 0, m2 = 0;  // "value d present" masks of this map (bit 7 per pixel byte)
                classify_row(raw[u][0], lut.n, ge_add, lane, rp, h3p, hb3p, lrp, lrbp);
                classify_row(raw[u][1], lut.n, ge_add, lane, rc, h3c, hb3c, lrc, lrbc);
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    classify_row(raw[u][r + 2], lut.n, ge_add, lane, rn, h3n, hb3n, lrn, lrbn);
                    uint32_t S, BG;
                    if (AXIAL) { S = rp.onehot | rn.onehot | lrc; BG = rp.bg | rn.bg | lrbc; }
                    else { S = h3p | h3c | h3n; BG = hb3p | hb3c | hb3n; }
                    const uint32_t NEGW = __byte_perm(lut.neg_lo, lut.neg_hi, rc.sel);
                    const uint32_t POSW = __byte_perm(lut.pos_lo, lut.pos_hi, rc.sel);
                    const uint32_t anyneg7 = nonzero7(NEGW & S) & rc.act7;
                    const uint32_t notall7 = nonzero7((S & ~POSW) | BG);
                    const uint32_t b1 = anyneg7;
                    const uint32_t b0 = (notall7 & ~anyneg7 & rc.act7) | rc.odd7;
                    const uint32_t cw = ((b1 >> 6) | (b0 >> 7)) & 0x03030303u;
                    acc[r][0] |= __byte_perm(cw, 0u, 0x4140) << (2 * t);
                    acc[r][1] |= __byte_perm(cw, 0u, 0x4342) << (2 * t);
                    if (y0 + r >= row_lo && y0 + r < row_hi) {
                        m2 |= b1;
                        m1 |= b0;
                        m0 |= ~(b1 | b0);
                    }
                    rp = rc; rc = rn;
                    h3p = h3c; hb3p = hb3c; h3c = h3n; hb3c = hb3n; lrc = lrn; lrbc = lrbn;
                }
                seen |= (((m0 & inimg7) ? 1u : 0u) | (m1 ? 2u : 0u) | (m2 ? 4u : 0u)) << (3 * t);
            }
        }
        if (x4 < W) {
            uint16_t* cout = codes + (size_t)b * plane_sz;
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                const int y = y0 + r;
                if (y < H) {
                    uint16_t* dst = cout + (size_t)y * W + x4;
                    if (FAST && x4 + 3 < W) {
                        *(uint2*)dst = make_uint2(acc[r][0], acc[r][1]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (x4 + i < W) dst[i] = (uint16_t)(acc[r][i >> 1] >> (16 * (i & 1)));
                    }
                }
            }
        }
    }
    seen = __reduce_or_sync(0xffffffffu, seen);
    if (threadIdx.x == 0 && seen) atomicOr(flags + b, seen);
}

// ---- bit-sliced variant (5 / 9 classes, W % 8 == 0) ---------------------------------------------------
// One 32-bit register holds ONE BIT of 32 consecutive pixels, so the whole stencil is LOP3 work on 32 pixels at
// a time: ~6 thread-instructions per (pixel, map) instead of ~31 for the byte-SIMD kernel above.
//   * a lane owns 32 pixels of a row (a warp spans 1024 columns); per row and map it loads the 32 class bytes
//     (4 x 64-bit loads), packs two pixels into one byte (ids < 16: one multiply-add per word pair, which is also
//     the first transpose stage) and transposes 16 bytes x 8 bits -> the 4 id-bit planes (a 4x4 byte transpose
//     with PRMT + two mask/shift stages);
//   * class planes P_k = "pixel has direction class k" are decoded from the four id bits, ids >= n decode to the
//     "unknown" plane (zero vector as a neighbour, centre code 1), ids >= 16 are mapped to 15 first (rare branch);
//   * 3-wide horizontal OR: neighbour bits come from the adjacent lanes (rotating shuffle + funnel shift); the
//     3-row OR is a rolling window down the strip;
//   * d = 2 iff the centre's class has an opposing class (ring distance 3..5 of 8, or 2 of 4) in its neighbourhood
//     set, d = 0 iff the set lies within ring distance 1 (resp. 0) and holds no zero vector -- the same table
//     ddm_build_lut derives from the reference's arithmetic (the launcher checks that it has this ring form);
//   * warp w of a block handles map w (T = 8) or the w-th sub-strip of rows (T = 1); the two code bit-planes of
//     every (row, map) go to shared memory, then the block transposes 16 planes x 32 pixels -> 32 uint16 code words
//     and stores them with 128-bit stores in the layout k_boost_inside reads.
// Columns: a tile up to 1023 px wide is one chunk (the rotating shuffle hands lane 0 the pixel right of the
// chunk, which lies outside the image = zero vector, exactly what the left border needs); wider tiles are cut
// into chunks of 960 owned pixels with one halo lane on either side.
constexpr int kBitsWarps = 8;

// ids >= 16 -> 15 (still "unknown", but it survives the nibble packing)
__device__ __forceinline__ void bits_clamp_ids(uint32_t w[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint32_t m = w[i] & 0xf0f0f0f0u;
        m |= m >> 1;
        m |= m >> 2;
        m = ((m >> 4) & 0x01010101u) * 0xffu;  // 0xff in every byte whose high nibble is non-zero
        w[i] = (w[i] & ~m) | (m & 0x0f0f0f0fu);
    }
}

// 32 class bytes of one row (8 words, pixel 4i + b in byte b of word i) -> the 4 id-bit planes (bit p = pixel p)
__device__ __forceinline__ void bits_transpose_row(uint32_t w[8], uint32_t G[4]) {
    if ((w[0] | w[1] | w[2] | w[3] | w[4] | w[5] | w[6] | w[7]) & 0xf0f0f0f0u)
        bits_clamp_ids(w);  // ids >= 16 somewhere in these 32 pixels (never, for an argmax over <= 9 channels)
    // nibble packing does the first transpose stage for free: byte b of E[m] = pixel 8m+b (low) | pixel 8m+4+b (high)
    uint32_t E[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) E[m] = w[2 * m + 1] * 16u + w[2 * m];
    // 4x4 byte transpose: G[b] = bytes b of E[0..3]
    const uint32_t t0 = __byte_perm(E[0], E[1], 0x5140), t1 = __byte_perm(E[2], E[3], 0x5140);
    const uint32_t t2 = __byte_perm(E[0], E[1], 0x7362), t3 = __byte_perm(E[2], E[3], 0x7362);
    uint32_t g0 = __byte_perm(t0, t1, 0x5410), g1 = __byte_perm(t0, t1, 0x7632);
    uint32_t g2 = __byte_perm(t2, t3, 0x5410), g3 = __byte_perm(t2, t3, 0x7632);
    // two bit stages inside every nibble: register index (b1 b0) <-> id bit index (j1 j0)
    {
        const uint32_t a0 = g0, c0 = g2, a1 = g1, c1 = g3;
        g0 = bitsel(0x33333333u, a0, c0 << 2);
        g2 = bitsel(0x33333333u, a0 >> 2, c0);
        g1 = bitsel(0x33333333u, a1, c1 << 2);
        g3 = bitsel(0x33333333u, a1 >> 2, c1);
    }
    G[0] = bitsel(0x55555555u, g0, g1 << 1);
    G[1] = bitsel(0x55555555u, g0 >> 1, g1);
    G[2] = bitsel(0x55555555u, g2, g3 << 1);
    G[3] = bitsel(0x55555555u, g2 >> 1, g3);
}

// NC direction classes (8: ring of 45-degree steps, 4: ring of 90-degree steps) + plane NC = "acts as the zero vector"
template <int NC>
struct BitRow {
    uint32_t H[NC + 1];  // NC == 8: 3-wide horizontal OR of the class planes; NC == 4: left | right only
    uint32_t P[NC + 1];  // the pixel's own class planes (NC == 4 also uses them as the vertical neighbours)
    uint32_t act, odd;   // valid direction class / non-zero unknown id
    uint32_t idb[3];     // the three low id bits of the pixel's class (NC == 8: selectors of the per-centre muxes)
};

template <int NC>
__device__ __forceinline__ void bits_decode_row(const uint32_t B[4], int lane, BitRow<NC>& r) {
    const uint32_t B0 = B[0], B1 = B[1], B2 = B[2], B3 = B[3];
    r.idb[0] = B0; r.idb[1] = B1; r.idb[2] = B2;
    const uint32_t m0 = ~B3 & ~B2, m1 = ~B3 & B2;
    r.P[0] = m0 & ~B1 & B0;
    r.P[1] = m0 & B1 & ~B0;
    r.P[2] = m0 & B1 & B0;
    r.P[3] = m1 & ~B1 & ~B0;
    if (NC == 8) {
        const uint32_t nz = B2 | B1 | B0;
        r.P[4] = m1 & ~B1 & B0;
        r.P[5] = m1 & B1 & ~B0;
        r.P[6] = m1 & B1 & B0;
        r.P[7] = B3 & ~nz;
        r.act = B3 ^ nz;
        r.odd = B3 & nz;
    } else {
        r.act = r.P[0] | r.P[1] | r.P[2] | r.P[3];
        r.odd = (B3 | B2 | B1 | B0) & ~r.act;
    }
    r.P[NC] = ~r.act;  // class 0, unknown ids and everything outside the image: the zero vector
    const int src_l = (lane + 31) & 31, src_r = (lane + 1) & 31;
#pragma unroll
    for (int k = 0; k <= NC; ++k) {
        const uint32_t x = r.P[k];
        const uint32_t l = __shfl_sync(0xffffffffu, x, src_l), rr = __shfl_sync(0xffffffffu, x, src_r);
        const uint32_t lr = __funnelshift_l(l, x, 1) | __funnelshift_r(x, rr, 1);
        r.H[k] = NC == 8 ? (lr | x) : lr;
    }
}

// code bit-planes of the centre row c; upV / dnV = the planes the rows above / below contribute (their 3-wide OR
// for the 8-neighbourhood, their own planes for the axial one)
template <int NC>
__device__ __forceinline__ void bits_codes(const uint32_t upV[NC + 1], const BitRow<NC>& c, const uint32_t dnV[NC + 1],
                                           uint32_t& b0, uint32_t& b1) {
    uint32_t S[NC + 1];
#pragma unroll
    for (int k = 0; k <= NC; ++k) S[k] = upV[k] | c.H[k] | dnV[k];
    if (NC == 8) {
        // Per-centre selection with muxes on the centre's id bits instead of eight AND-OR terms per test.  For class k
        // (a = k - 1 on the ring) the opposing set is the 3-window centred at a + 4 and the perpendicular pair is
        // {a + 2, a + 6}.  With the mux index i = k mod 8 (class 8 -> 0; the pixel's low id bits as they are):
        //   neg  <- T[(i + 3) & 7],  T[c] = S[c-1] | S[c] | S[c+1]
        //   zero <- X[(i + 1) & 3],  X[j] = S[j] | S[j+4]
        uint32_t Tw[8], X[4];
#pragma unroll
        for (int cidx = 0; cidx < 8; ++cidx) Tw[cidx] = S[(cidx + 7) & 7] | S[cidx] | S[(cidx + 1) & 7];
#pragma unroll
        for (int j = 0; j < 4; ++j) X[j] = S[j] | S[j + 4];
        const uint32_t i0 = c.idb[0], i1 = c.idb[1], i2 = c.idb[2];
        uint32_t m4[4], m2[2];
#pragma unroll
        for (int j = 0; j < 4; ++j) m4[j] = bitsel(i0, Tw[(2 * j + 1 + 3) & 7], Tw[(2 * j + 3) & 7]);  // index 2j+1 : 2j
        m2[0] = bitsel(i1, m4[1], m4[0]);
        m2[1] = bitsel(i1, m4[3], m4[2]);
        const uint32_t tsel = bitsel(i2, m2[1], m2[0]);
        const uint32_t x0 = bitsel(i0, X[(1 + 1) & 3], X[(0 + 1) & 3]), x1 = bitsel(i0, X[(3 + 1) & 3], X[(2 + 1) & 3]);
        const uint32_t xsel = bitsel(i1, x1, x0);
        b1 = c.act & tsel;
        b0 = (c.act & (xsel | S[NC]) & ~tsel) | c.odd;
        return;
    }
    uint32_t neg = 0, np = 0;
#pragma unroll
    for (int a = 0; a < NC; ++a) {
        const uint32_t N = S[(a + 2) & 3];                           // opposite diagonal: cos rounds to -1
        const uint32_t NZ = N | S[(a + 1) & 3] | S[(a + 3) & 3];     // ... or perpendicular: cos rounds to 0
        neg |= c.P[a] & N;
        np |= c.P[a] & NZ;
    }
    np |= c.act & S[NC];
    b1 = neg;
    b0 = (np & ~neg) | c.odd;
}

template <int T, int NC, int MB>
__global__ void __launch_bounds__(32 * kBitsWarps, MB) k_ddm_bits(const uint8_t* __restrict__ cls_maps,
                                                                 uint16_t* __restrict__ codes, uint32_t* __restrict__ flags,
                                                                 int H, int W, int R, int chunk_px, int halo, int row_lo,
                                                                 int row_hi) {
    CDNET_DYN_SHARED(uint32_t, s_planes);  // [block rows][2 T planes][32 lanes]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int b = blockIdx.z;
    constexpr int SUB = kBitsWarps / T;  // row sub-strips per block (1 for T = 8, 8 for T = 1)
    const int t = T == 8 ? wid : 0;
    const int sub = T == 8 ? 0 : wid;
    const int yb = blockIdx.y * (R * SUB);  // first row of the block
    const int y0 = yb + sub * R;            // first row of this warp
    const int xl = blockIdx.x * chunk_px + (lane - halo) * 32;  // first pixel of this lane's word
    const size_t plane_sz = (size_t)H * W;
    const uint8_t* plane = cls_maps + ((size_t)b * T + t) * plane_sz;
    // pixels of this word inside the image and owned by this lane (halo lanes own nothing)
    uint32_t own = 0;
    if (!(halo && (lane == 0 || lane == 31))) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (xl + 8 * k >= 0 && xl + 8 * k < W) own |= 0xffu << (8 * k);
    }
    auto load_row = [&](int y, uint32_t w[8]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = 0;
        if ((unsigned)y < (unsigned)H) {
            const uint8_t* row = plane + (size_t)y * W + xl;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int x = xl + 8 * k;
                if (x >= 0 && x < W) {  // W % 8 == 0: an 8-pixel group is inside or outside as a whole
                    const uint2 v = __ldg((const uint2*)(row + 8 * k));
                    w[2 * k] = v.x;
                    w[2 * k + 1] = v.y;
                }
            }
        }
    };
    uint32_t seen0 = 0, seen1 = 0, seen2 = 0;
    if (y0 < H) {
        uint32_t w[8], nw[8];
        load_row(y0 - 1, w);
        const int rows_here = min(R, H - y0);
        auto emit = [&](const uint32_t upV[NC + 1], const BitRow<NC>& c, const uint32_t dnV[NC + 1], int y) {
            // y = row of c, y0 <= y < y0 + rows_here
            uint32_t b0, b1;
            bits_codes<NC>(upV, c, dnV, b0, b1);
            uint32_t* dst = s_planes + ((size_t)(y - yb) * (2 * T) + 2 * t) * 32 + lane;
            dst[0] = b0;
            dst[32] = b1;
            if (y >= row_lo && y < row_hi) {
                seen2 |= b1 & own;
                seen1 |= b0 & own;
                seen0 |= ~(b1 | b0) & own;
            }
        };
        // rolling window: the vertical-neighbour planes of the row two back, the whole previous row, the new row
        uint32_t upV[NC + 1];
        BitRow<NC> prev;
#pragma unroll
        for (int k = 0; k <= NC; ++k) { upV[k] = 0; prev.P[k] = 0; prev.H[k] = 0; }
        prev.act = prev.odd = 0;
        prev.idb[0] = prev.idb[1] = prev.idb[2] = 0;
        // step i brings in row y0 - 1 + i and emits the centre row y0 + i - 2 (three steps rotate the window once)
#pragma unroll 3
        for (int i = 0; i < rows_here + 2; ++i) {
            load_row(y0 + i, nw);  // request the next row before the arithmetic
            uint32_t G[4];
            bits_transpose_row(w, G);
            BitRow<NC> rc;
            bits_decode_row<NC>(G, lane, rc);
            if (i >= 2) emit(upV, prev, NC == 8 ? rc.H : rc.P, y0 + i - 2);
#pragma unroll
            for (int k = 0; k <= NC; ++k) upV[k] = NC == 8 ? prev.H[k] : prev.P[k];
            prev = rc;
#pragma unroll
            for (int k = 0; k < 8; ++k) w[k] = nw[k];
        }
    }
    {
        const uint32_t s = ((seen0 ? 1u : 0u) | (seen1 ? 2u : 0u) | (seen2 ? 4u : 0u)) << (3 * t);
        const uint32_t all = __reduce_or_sync(0xffffffffu, s);
        if (lane == 0 && all) atomicOr(flags + b, all);
    }
    __syncthreads();
    // ---- 2 T bit-planes x 32 pixels -> 32 code words per (row, lane) cell --------------------------------
    const int block_rows = min(R * SUB, H - yb);
    uint16_t* cout = codes + (size_t)b * plane_sz;
    for (int cell = threadIdx.x; cell < block_rows * 32; cell += 32 * kBitsWarps) {
        const int r = cell >> 5, l = cell & 31;
        if (halo && (l == 0 || l == 31)) continue;
        const int x = blockIdx.x * chunk_px + (l - halo) * 32;
        if (x >= W) continue;
        const uint32_t* src = s_planes + (size_t)r * (2 * T) * 32 + l;
        uint32_t p[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) p[j] = j < 2 * T ? src[j * 32] : 0u;
        // 16x16 bit transpose of both halves: afterwards p[i] = code(pixel i) | code(pixel i + 16) << 16
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t a = p[j], c = p[j + 8];
            p[j] = __byte_perm(a, c, 0x6240);
            p[j + 8] = __byte_perm(a, c, 0x7351);
        }
#pragma unroll
        for (int h = 0; h < 16; h += 8)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t a = p[h + j], c = p[h + j + 4];
                p[h + j] = bitsel(0x0f0f0f0fu, a, c << 4);
                p[h + j + 4] = bitsel(0x0f0f0f0fu, a >> 4, c);
            }
#pragma unroll
        for (int h = 0; h < 16; h += 4)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const uint32_t a = p[h + j], c = p[h + j + 2];
                p[h + j] = bitsel(0x33333333u, a, c << 2);
                p[h + j + 2] = bitsel(0x33333333u, a >> 2, c);
            }
#pragma unroll
        for (int h = 0; h < 16; h += 2) {
            const uint32_t a = p[h], c = p[h + 1];
            p[h] = bitsel(0x55555555u, a, c << 1);
            p[h + 1] = bitsel(0x55555555u, a >> 1, c);
        }
        uint16_t* dst = cout + (size_t)(yb + r) * W + x;
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
            if (x + 8 * g8 >= W) break;
            // pixels 8 g8 .. 8 g8 + 7: low halves of p[8 g8 ..] for g8 < 2, high halves of p[8 (g8 - 2) ..] otherwise
            const int base = 8 * (g8 & 1);
            const uint32_t sel = g8 < 2 ? 0x5410u : 0x7632u;
            uint4 v;
            v.x = __byte_perm(p[base + 0], p[base + 1], sel);
            v.y = __byte_perm(p[base + 2], p[base + 3], sel);
            v.z = __byte_perm(p[base + 4], p[base + 5], sel);
            v.w = __byte_perm(p[base + 6], p[base + 7], sel);
            *(uint4*)(dst + 8 * g8) = v;
        }
    }
}

// the bit-sliced kernel hard-wires the ring structure of the table: check it against the table derived from the
// reference's arithmetic before trusting it
static bool lut_is_ring(const DdmLut& lut) {
    const int nc = lut.n - 1;
    if (nc != 8 && nc != 4) return false;
    for (int a = 1; a <= nc; ++a)
        for (int bq = 0; bq <= nc; ++bq) {
            int want;  // rounded cosine
            if (bq == 0) want = 0;
            else {
                const int d = ((a - bq) % nc + nc) % nc;
                const int dist = d < nc - d ? d : nc - d;
                if (nc == 8) want = dist <= 1 ? 1 : (dist == 2 ? 0 : -1);
                else want = dist == 0 ? 1 : (dist == 1 ? 0 : -1);
            }
            const int have = ((lut.pos[a] >> bq) & 1u) ? 1 : (((lut.neg[a] >> bq) & 1u) ? -1 : 0);
            if (have != want) return false;
        }
    return true;
}

static int bits_rows_per_warp() {
    static int rows = 0;
    if (!rows) {
        const char* e = getenv("CDNET_DDM_BITS_ROWS");
        rows = e ? atoi(e) : 16;
        if (rows < 2 || rows > 64 || (rows & 1)) rows = 16;
    }
    return rows;
}

template <int T, int NC, int MB>
static int launch_bits_mb(const uint8_t* cls_maps, uint16_t* codes, uint32_t* flags, int B, int H, int W, cudaStream_t st,
                          int row_lo, int row_hi) {
    const int R = bits_rows_per_warp();
    constexpr int SUB = kBitsWarps / T;
    const int halo = W > 1023 ? 1 : 0;
    const int chunk_px = halo ? 960 : 1024;
    const size_t smem = (size_t)R * SUB * 2 * T * 32 * sizeof(uint32_t);
    static bool attr_done = false;
    if (!attr_done) {
        CDNET_CUDA_OK(cudaFuncSetAttribute(k_ddm_bits<T, NC, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr_done = true;
    }
    dim3 grid(ceil_div(W, chunk_px), ceil_div(H, R * SUB), B);
    CDNET_LAUNCH((k_ddm_bits<T, NC, MB>), grid, 32 * kBitsWarps, smem, st, cls_maps, codes, flags, H, W, R, chunk_px, halo,
                 row_lo, row_hi);
    return last_error();
}

template <int T, int NC>
static int launch_bits(const uint8_t* cls_maps, uint16_t* codes, uint32_t* flags, int B, int H, int W, cudaStream_t st,
                       int row_lo, int row_hi) {
    static int mb = 0;  // resident blocks per SM the kernel is compiled for: 3 (80 registers) or 2 (96)
    if (!mb) { const char* e = getenv("CDNET_DDM_BITS_MB"); mb = (e && atoi(e) == 2) ? 2 : 3; }
    if (mb == 2) return launch_bits_mb<T, NC, 2>(cls_maps, codes, flags, B, H, W, st, row_lo, row_hi);
    return launch_bits_mb<T, NC, 3>(cls_maps, codes, flags, B, H, W, st, row_lo, row_hi);
}

// normalised value of code d for a map whose present-value bits are f (3 bits): (d-min)/(max-min)
// in f32 (getDirectionDiffMap.py:104-106); constant map -> 0/0 = NaN.
__device__ __forceinline__ float ddm_value(uint32_t d, uint32_t f) {
    const int mn = (f & 1) ? 0 : ((f & 2) ? 1 : 2);
    const int mx = (f & 4) ? 2 : ((f & 2) ? 1 : 0);
    return __fdiv_rn((float)((int)d - mn), (float)(mx - mn));
}

__global__ void k_ddm_normalize(const uint16_t* __restrict__ codes, const uint32_t* __restrict__ flags,
                                float* __restrict__ out, int32_t* __restrict__ status, size_t plane_sz) {
    const int b = blockIdx.y;
    const uint32_t f = flags[b] & 7u;
    const bool constant = (f == 1u || f == 2u || f == 4u || f == 0u);
    if (constant && status && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(status + b, CDNET_S_DDM_CONSTANT);
    const float v0 = ddm_value(0, f), v1 = ddm_value(1, f), v2 = ddm_value(2, f);
    const uint16_t* c = codes + (size_t)b * plane_sz;
    float* o = out + (size_t)b * plane_sz;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane_sz; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t d = c[i] & 3u;
        o[i] = d == 0 ? v0 : (d == 1 ? v1 : v2);
    }
}

// launcher shared with postproc.cu
template <int T>
static void launch_simd(const uint8_t* cls_maps, uint16_t* codes, uint32_t* flags, int H, int W, const DdmLut& lut,
                        bool fast, dim3 grid, dim3 block, cudaStream_t st, int row_lo, int row_hi) {
    DdmLut8 l8;
    l8.n = lut.n;
    l8.axial = lut.axial;
    l8.neg_lo = l8.neg_hi = l8.pos_lo = l8.pos_hi = 0;
    for (int k = 0; k < 8; ++k) {
        const uint32_t ng = (k + 1 < lut.n) ? ((lut.neg[k + 1] >> 1) & 0xffu) : 0u;
        const uint32_t ps = (k + 1 < lut.n) ? ((lut.pos[k + 1] >> 1) & 0xffu) : 0u;
        if (k < 4) { l8.neg_lo |= ng << (8 * k); l8.pos_lo |= ps << (8 * k); }
        else { l8.neg_hi |= ng << (8 * (k - 4)); l8.pos_hi |= ps << (8 * (k - 4)); }
    }
    static int rows = 0;
    if (!rows) {
        const char* e = getenv("CDNET_DDM_ROWS");
        rows = e ? atoi(e) : CDNET_DDM_ROWS_DEFAULT;
        if (rows != 8 && rows != 4 && rows != 2) rows = 4;
    }
    grid.y = ceil_div(H, 4 * rows);
    if (rows == 8) {
        if (lut.axial) {
            if (fast) CDNET_LAUNCH((k_ddm_codes_simd<T, true, true, 8>), grid, block, 0, st, cls_maps, codes, flags, H, W, l8, row_lo, row_hi);
            else CDNET_LAUNCH((k_ddm_codes_simd<T, false, true, 8>), grid, block, 0, st, cls_maps, codes, flags, H, W, l8, row_lo, row_hi);
        } else {
            if (fast) CDNET_LAUNCH((k_ddm_codes_simd<T, true, false, 8>), grid, block, 0, st, cls_maps, codes, flags, H, W, l8, row_lo, row_hi);
            else CDNET_LAUNCH((k_ddm_codes_simd<T, false, false, 8>), grid, block, 0, st, cls_maps, codes, flags, H, W, l8, row_lo, row_hi);
        }
    } else if (rows == 2) {
        if (lut.axial) {
            if (fast) CDNET_LAUNCH((k_ddm_codes_simd<T, true, true, 2>), grid, block, 0, st, cls_maps, codes, flags, H, W, l8, row_lo, row_hi);
            else CDNET_LAUNCH((k_ddm_codes_simd<T, false, true, 2>), grid, block, 0, st, cls_maps, codes, flags, H, W, l8, row_lo, row_hi);
        } else {
            if (fast) CDNET_LAUNCH((k_ddm_codes_simd<T, true, false, 2>), grid, block, 0, st, cls_maps, codes, flags, H, W, l8, row_lo, row_hi);
            else CDNET_LAUNCH((k_ddm_codes_simd<T, false, false, 2>), grid, block, 0, st, cls_maps, codes, flags, H, W, l8, row_lo, row_hi);
        }
    } else {
        if (lut.axial) {
            if (fast) CDNET_LAUNCH((k_ddm_codes_simd<T, true, true, 4>), grid, block, 0, st, cls_maps, codes, flags, H, W, l8, row_lo, row_hi);
            else CDNET_LAUNCH((k_ddm_codes_simd<T, false, true, 4>), grid, block, 0, st, cls_maps, codes, flags, H, W, l8, row_lo, row_hi);
        } else {
            if (fast) CDNET_LAUNCH((k_ddm_codes_simd<T, true, false, 4>), grid, block, 0, st, cls_maps, codes, flags, H, W, l8, row_lo, row_hi);
            else CDNET_LAUNCH((k_ddm_codes_simd<T, false, false, 4>), grid, block, 0, st, cls_maps, codes, flags, H, W, l8, row_lo, row_hi);
        }
    }
}

// launcher shared with postproc.cu
int ddm_codes_launch(const uint8_t* cls_maps, uint16_t* codes, uint32_t* flags, int B, int T, int H, int W,
                     int n_classes, cudaStream_t st, int row_lo, int row_hi) {
    if (row_hi < 0) row_hi = H;
    DdmLut lut;
    if (!ddm_build_lut(n_classes, &lut)) return CDNET_E_BADARG;
    if (T != 1 && T != 8) return CDNET_E_BADARG;
    CDNET_CUDA_OK(cudaMemsetAsync(flags, 0, sizeof(uint32_t) * (size_t)B, st));
    dim3 block(32, 4);
    dim3 grid(ceil_div(W, 128), ceil_div(H, 4 * kRows), B);
    const bool fast = (W % 4 == 0) && (((uintptr_t)cls_maps & 3) == 0) && (((uintptr_t)codes & 7) == 0);
    static int impl = -1;  // CDNET_DDM_IMPL=simd keeps the byte-SIMD kernel (A/B timing)
    if (impl < 0) { const char* e = getenv("CDNET_DDM_IMPL"); impl = (e && e[0] == 's') ? 1 : 0; }
    if (n_classes <= 9 && impl == 0 && W % 8 == 0 && (((uintptr_t)cls_maps & 7) == 0) && (((uintptr_t)codes & 15) == 0) &&
        lut_is_ring(lut)) {
        // bit-sliced kernel: 32 pixels per register
        if (n_classes == 9) return T == 8 ? launch_bits<8, 8>(cls_maps, codes, flags, B, H, W, st, row_lo, row_hi)
                                          : launch_bits<1, 8>(cls_maps, codes, flags, B, H, W, st, row_lo, row_hi);
        return T == 8 ? launch_bits<8, 4>(cls_maps, codes, flags, B, H, W, st, row_lo, row_hi)
                      : launch_bits<1, 4>(cls_maps, codes, flags, B, H, W, st, row_lo, row_hi);
    }
    if (n_classes <= 9) {
        // ragged widths / unaligned planes: byte-SIMD kernel
        if (T == 8) launch_simd<8>(cls_maps, codes, flags, H, W, lut, fast, grid, block, st, row_lo, row_hi);
        else launch_simd<1>(cls_maps, codes, flags, H, W, lut, fast, grid, block, st, row_lo, row_hi);
    } else if (T == 8) {
        if (fast) CDNET_LAUNCH((k_ddm_codes<8, true>), grid, block, 0, st, cls_maps, codes, flags, H, W, lut, row_lo, row_hi);
        else CDNET_LAUNCH((k_ddm_codes<8, false>), grid, block, 0, st, cls_maps, codes, flags, H, W, lut, row_lo, row_hi);
    } else {
        if (fast) CDNET_LAUNCH((k_ddm_codes<1, true>), grid, block, 0, st, cls_maps, codes, flags, H, W, lut, row_lo, row_hi);
        else CDNET_LAUNCH((k_ddm_codes<1, false>), grid, block, 0, st, cls_maps, codes, flags, H, W, lut, row_lo, row_hi);
    }
    return last_error();
}

}  // namespace cdnet

using namespace cdnet;

extern "C" size_t cdnet_ddm_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    return pad256((size_t)B * H * W * sizeof(uint16_t)) + pad256((size_t)B * sizeof(uint32_t));
}

extern "C" int cdnet_ddm(const uint8_t* cls, float* out, int32_t* status, int B, int H, int W, int n_classes,
                         void* ws, size_t ws_bytes, void* stream) {
    if (!cls || !out || B <= 0 || H <= 0 || W <= 0 || (double)H * W >= 2147483648.0) return CDNET_E_BADARG;
    if (ws_bytes < cdnet_ddm_workspace_bytes(B, H, W)) return CDNET_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    Arena ar(ws, ws_bytes);
    uint16_t* codes = ar.take<uint16_t>((size_t)B * H * W);
    uint32_t* flags = ar.take<uint32_t>(B);
    if (!ar.ok) return CDNET_E_WORKSPACE;
    int rc = ddm_codes_launch(cls, codes, flags, B, 1, H, W, n_classes, st, 0, -1);
    if (rc) return rc;
    const size_t plane = (size_t)H * W;
    dim3 grid((unsigned)((plane + 256 * 8 - 1) / (256 * 8)), B);
    CDNET_LAUNCH(k_ddm_normalize, grid, 256, 0, st, codes, flags, out, status, plane);
    return last_error();
}

// ---- circshift, data_prepare/getDirectionDiffMap.py:14-42 -------------------------------------
namespace cdnet {
template <typename E>
__global__ void k_circshift(const E* __restrict__ in, E* __restrict__ out, int H, int W, int dy, int dx) {
    // out[y][x] = in[y+dy][x+dx], zero outside
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const size_t plane = (size_t)H * W * blockIdx.z;
    if (x >= W) return;
    const int sy = y + dy, sx = x + dx;
    E v = 0;
    if (sy >= 0 && sy < H && sx >= 0 && sx < W) v = in[plane + (size_t)sy * W + sx];
    out[plane + (size_t)y * W + x] = v;
}
}  // namespace cdnet

extern "C" int cdnet_circshift(const void* in, void* out, int C, int H, int W, int elem_bytes, int direction,
                               int shift1, int shift2, void* stream) {
    if (!in || !out || C <= 0 || H <= 0 || W <= 0 || direction < 1 || direction > 4 || shift1 < 0 || shift2 < 0)
        return CDNET_E_BADARG;
    // direction 1/2: rows move up (content of row y+s1 lands on y); 3/4: down.  1/3: columns move
    // left; 2/4: right.
    const int dy = (direction <= 2) ? shift1 : -shift1;
    const int dx = (direction == 1 || direction == 3) ? shift2 : -shift2;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(ceil_div(W, 256), H, C);
    switch (elem_bytes) {
        case 1: CDNET_LAUNCH(k_circshift<uint8_t>, grid, 256, 0, st, (const uint8_t*)in, (uint8_t*)out, H, W, dy, dx); break;
        case 2: CDNET_LAUNCH(k_circshift<uint16_t>, grid, 256, 0, st, (const uint16_t*)in, (uint16_t*)out, H, W, dy, dx); break;
        case 4: CDNET_LAUNCH(k_circshift<uint32_t>, grid, 256, 0, st, (const uint32_t*)in, (uint32_t*)out, H, W, dy, dx); break;
        case 8: CDNET_LAUNCH(k_circshift<unsigned long long>, grid, 256, 0, st, (const unsigned long long*)in,
                             (unsigned long long*)out, H, W, dy, dx); break;
        default: return CDNET_E_BADARG;
    }
    return last_error();
}

// whole-slide shard: codes for T (1 or 8) maps of one extended tile [T,He,W]; the "value present" flags are
// accumulated over the rows [row_lo, row_hi) only (the shard's own rows) and are all-reduced by the host.
extern "C" int cdnet_shard_ddm_codes(const uint8_t* dcm, uint16_t* codes, uint32_t* flags, int T, int He, int W,
                                     int n_classes, int row_lo, int row_hi, void* stream) {
    if (!dcm || !codes || !flags || He <= 0 || W <= 0 || (double)He * W >= 2147483648.0) return CDNET_E_BADARG;
    return ddm_codes_launch(dcm, codes, flags, 1, T, He, W, n_classes, (cudaStream_t)stream, row_lo, row_hi);
}
