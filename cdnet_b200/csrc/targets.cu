// targets.cu -- instance-label -> training-target transform (K14-K19).
//
// Replaces LabelEncoding.__call__ (my_transforms_direction.py:697-885, out_c = 3, do_direction = 1)
// and get_centerpoint2 (my_transforms_direction.py:651-685), with the quantiser of
// data_prepare/SegFix_offset_helper.py:311-341,423-450,486-506 and the Sobel kernel of :102-132.
//
// O(H*W) reformulation of the reference's per-nucleus full-canvas loop (:800-835), exact by
// SURVEY.md Appendix B:
//   * centre of nucleus k = first raster pixel of maximum centerness; computed for ALL labels at
//     once: every foreground pixel bisects its 8 rays against `inst == own label`, a u64 atomicMax
//     on the f64 bit pattern picks the maximum per label and an atomicMin the first pixel;
//   * EDT(1 - point) is the Euclidean distance to the centre; its per-nucleus maximum over the
//     cross-dilated support is an integer atomicMax of dy^2+dx^2 (sqrt is monotone);
//   * "last writer wins" compositing == the winner w(p) = max label in p's cross neighbourhood;
//     dir(p) = sum over the 11x11 taps q of K[q-p] * f32(dc_w(q)), evaluated as the sequential f32
//     FMA chain in (kh, kw) order that torch's CPU conv2d was measured to produce (zero taps are
//     exact no-ops, so only taps inside w's dilated support are touched);
//   * the double quantisation (:853-855) is idempotent: class = align_angle(angle) index;
//   * Gaussian point map = separable 17-tap f64 correlation (scipy's symmetric summation order,
//     reflect border) of 255 at the centres, cast to f16.
#include <cuda_fp16.h>
#include <math.h>

#include <string.h>

#include <mutex>

#include "internal.h"

namespace cdnet {

__constant__ float c_sobel[2][121];
__constant__ double c_gauss[17];
// host-side record of what has been uploaded to the constants of each device
constexpr int kMaxDevices = 64;
struct TargetConsts {
    bool sobel_done = false, gauss_done = false;
    double gauss[17];
    int n_sm = 0;
};
// (sin, cos)(2*pi/8*k) exactly as CPython/numba's libm produces them (my_transforms_direction.py:657-658)
__constant__ double c_rays[8][2] = {
    {0x0.0p+0, 0x1.0000000000000p+0},
    {0x1.6a09e667f3bccp-1, 0x1.6a09e667f3bcdp-1},
    {0x1.0000000000000p+0, 0x1.1a62633145c07p-54},
    {0x1.6a09e667f3bcdp-1, -0x1.6a09e667f3bccp-1},
    {0x1.1a62633145c07p-53, -0x1.0000000000000p+0},
    {-0x1.6a09e667f3bccp-1, -0x1.6a09e667f3bcep-1},
    {-0x1.0000000000000p+0, -0x1.a79394c9e8a0ap-53},
    {-0x1.6a09e667f3bcep-1, 0x1.6a09e667f3bcbp-1},
};

constexpr int kBX = 128, kBY = 4;
static inline dim3 px_grid(int B, int H, int W) { return dim3(ceil_div(W, kBX), ceil_div(H, kBY), B); }
static inline dim3 px_block() { return dim3(kBX, kBY); }
#define PX_COORDS                                              \
    const int x = blockIdx.x * kBX + threadIdx.x;              \
    const int y = blockIdx.y * kBY + threadIdx.y;              \
    const int b = blockIdx.z;                                  \
    const bool inb = (x < W) && (y < H);                       \
    const size_t tile = (size_t)b * H * W;                     \
    const int p = y * W + x;                                   \
    const int lane = threadIdx.x & 31;                         \
    (void)lane; (void)p; (void)tile; (void)inb;

// ---- label statistics ------------------------------------------------------------------------------
// presence[b][v] = 1 if value v occurs; fg[b] = number of non-zero pixels
__global__ void __launch_bounds__(256) k_label_stats(const uint8_t* __restrict__ ids, int* __restrict__ presence,
                                                     int* __restrict__ fg, size_t plane) {
    __shared__ int s_pres[256];
    __shared__ int s_cnt;
    const int b = blockIdx.y;
    s_pres[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const uint8_t* I = ids + (size_t)b * plane;
    int cnt = 0;
    if (plane % 16 == 0 && (((uintptr_t)I) & 15) == 0) {
        // 16 pixels per load; words that are all background (most of a tile) touch shared memory once at most
        const uint4* I4 = (const uint4*)I;
        const size_t n4 = plane / 16;
        bool zero_seen = false;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
            const uint4 q = __ldg(I4 + i);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t nz = (((w[j] & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w[j]) & 0x80808080u;
                cnt += __popc(nz);
                zero_seen |= nz != 0x80808080u;
                if (nz) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t v = (w[j] >> (8 * k)) & 0xffu;
                        if (v) s_pres[v] = 1;
                    }
                }
            }
        }
        if (zero_seen) s_pres[0] = 1;
    } else {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (size_t)gridDim.x * blockDim.x) {
            const int v = I[i];
            s_pres[v] = 1;
            cnt += v != 0;
        }
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_cnt, cnt);
    __syncthreads();
    if (s_pres[threadIdx.x]) presence[b * 256 + threadIdx.x] = 1;
    if (threadIdx.x == 0 && s_cnt) atomicAdd(fg + b, s_cnt);
}

// int32 ids (label images loaded without the uint8 truncation of data_folder.py:26-37): stats[b] = {non-zero pixels,
// zero present, smallest non-zero id, largest id}; a second pass looks for a third distinct non-zero value
__global__ void __launch_bounds__(256) k_label_stats_wide(const int* __restrict__ ids, int* __restrict__ stats, size_t plane) {
    const int b = blockIdx.y;
    const int* I = ids + (size_t)b * plane;
    int cnt = 0, zero = 0, mn = 0x7fffffff, mx = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (size_t)gridDim.x * blockDim.x) {
        const int v = I[i];
        if (v != 0) { ++cnt; mn = min(mn, v); mx = max(mx, v); } else zero = 1;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    zero = __reduce_max_sync(0xffffffffu, zero);
    mn = __reduce_min_sync(0xffffffffu, mn);
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0) {
        if (cnt) atomicAdd(stats + 4 * b, cnt);
        if (zero) atomicMax(stats + 4 * b + 1, 1);
        atomicMin(stats + 4 * b + 2, mn);
        atomicMax(stats + 4 * b + 3, mx);
    }
}
__global__ void __launch_bounds__(256) k_label_third_wide(const int* __restrict__ ids, const int* __restrict__ stats,
                                                          int* __restrict__ third, size_t plane) {
    const int b = blockIdx.y;
    const int* I = ids + (size_t)b * plane;
    const int mn = stats[4 * b + 2], mx = stats[4 * b + 3];
    int hit = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (size_t)gridDim.x * blockDim.x) {
        const int v = I[i];
        hit |= (v != 0 && v != mn && v != mx);
    }
    if (__any_sync(0xffffffffu, hit) && (threadIdx.x & 31) == 0) atomicMax(third + b, 1);
}
__global__ void k_stats_init(int* __restrict__ stats, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) { stats[4 * b] = 0; stats[4 * b + 1] = 0; stats[4 * b + 2] = 0x7fffffff; stats[4 * b + 3] = 0; }
}
__global__ void k_stats_distinct(const int* __restrict__ stats, const int* __restrict__ third, int* __restrict__ n_distinct,
                                 int* __restrict__ fg, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int cnt = stats[4 * b], zero = stats[4 * b + 1], mn = stats[4 * b + 2], mx = stats[4 * b + 3];
    int d = zero + (cnt > 0 ? 1 : 0) + ((cnt > 0 && mn != mx) ? 1 : 0) + third[b];
    n_distinct[b] = d > 3 ? 3 : d;
    fg[b] = cnt;
}
// fg[b] <- stats[b][0]
__global__ void k_stats_to_fg(const int* __restrict__ stats, int* __restrict__ fg, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) fg[b] = stats[4 * b];
}

// ---- ternary label / interior (my_transforms_direction.py:743-751, 763-770, 781) --------------------
template <typename IdT>
__global__ void __launch_bounds__(kBX* kBY) k_t_ternary(const IdT* __restrict__ ids, const int* __restrict__ fg,
                                                        int instance_level, uint8_t* __restrict__ ternary,
                                                        uint8_t* __restrict__ inside, uint8_t* __restrict__ interior,
                                                        int H, int W) {
    PX_COORDS
    if (!inb) return;
    const IdT* I = ids + tile;
    if (instance_level >= 2) {
        // out_c != 3 (:721-739): no boundary class.  Mode 2: 2 where an instance id is set; mode 3: erosion (cross
        // minimum) of 2 * (label > 127.5), the caller hands in max(channel 0, channel 1).  new_label_inside is a
        // copy of new_label, and the instances are labelled from it (mode 3) or from the ids themselves (mode 2).
        bool on;
        if (instance_level == 2) {
            on = I[p] > 0;
        } else {
            on = I[p] > 127;
            if (y > 0) on = on && I[p - W] > 127;
            if (y + 1 < H) on = on && I[p + W] > 127;
            if (x > 0) on = on && I[p - 1] > 127;
            if (x + 1 < W) on = on && I[p + 1] > 127;
        }
        ternary[tile + p] = on ? 255 : 0;
        inside[tile + p] = on;
        interior[tile + p] = on;
        return;
    }
    auto val = [&](int q) -> int { const int v = (int)I[q]; return instance_level ? v : (v > 127 ? 1 : 0); };
    const int v = val(p);
    int mx = v, mn = v;
    if (y > 0) { const int u = val(p - W); mx = max(mx, u); mn = min(mn, u); }
    if (y + 1 < H) { const int u = val(p + W); mx = max(mx, u); mn = min(mn, u); }
    if (x > 0) { const int u = val(p - 1); mx = max(mx, u); mn = min(mn, u); }
    if (x + 1 < W) { const int u = val(p + 1); mx = max(mx, u); mn = min(mn, u); }
    int nl = v > 0;
    // remove_small_objects(new_label, 5) on a {0,1} uint8 image treats the value 1 as ONE object (:746)
    if (instance_level && fg[b] < 5) nl = 0;
    const int ins = nl;
    if (mx != mn) nl = 2;  // dilation(ids) & ~erosion(ids) > 0  <=>  cross-max != cross-min
    ternary[tile + p] = nl == 0 ? 0 : (nl == 1 ? 127 : 255);
    inside[tile + p] = ins;
    interior[tile + p] = nl == 1;
}

// ---- a tile without background (out_c != 3 only) ---------------------------------------------------------
// The reference iterates `np.unique(label_instance)[1:]` (:797-800): it takes the first value for background.
// Where the instance map has no zero pixel that silently drops the smallest instance id (no centre, no support;
// its pixels stay foreground).  tile_min[b] = smallest id of tile b, then that id is cleared when it is not 0.
__global__ void k_t_inst_min(const int* __restrict__ inst, int* __restrict__ tile_min, size_t plane) {
    const int b = blockIdx.y;
    const int* I = inst + (size_t)b * plane;
    int m = 0x7fffffff;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (size_t)gridDim.x * blockDim.x)
        m = min(m, I[i]);
    m = __reduce_min_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0) atomicMin(tile_min + b, m);
}

__global__ void k_t_drop_first(int* __restrict__ inst, const int* __restrict__ tile_min, size_t plane) {
    const int b = blockIdx.y;
    const int m = tile_min[b];
    if (m <= 0) return;
    int* I = inst + (size_t)b * plane;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (size_t)gridDim.x * blockDim.x)
        if (I[i] == m) I[i] = 0;
}

// ---- centre search (my_transforms_direction.py:651-685) ---------------------------------------------
// The reference bisects [l, r] = [0, 1000] thirty times with mid = (l + r) / 2 and returns r (:668-678).  Every
// l, r, mid is a multiple of 1000 / 2^30 below 2^11, so all of that arithmetic is exact in f64 and the pair can be
// carried as (lo, width): mid = lo + width/2, r_final = lo_final + 1000 / 2^30.  Python's round() (half to even)
// of the f64 position is taken with the 1.5 * 2^52 trick: one DADD in round-to-nearest-even leaves the integer
// in the low word -- no FRND / F2I / f64 compares in the loop.
__device__ __forceinline__ double ray_reach(const int* __restrict__ L, int H, int W, int i, int j, int own, double sy,
                                            double sx) {
    const double kMagic = 6755399441055744.0;  // 1.5 * 2^52
    double lo = 0.0, half = 500.0;
    const double fi = (double)i, fj = (double)j;
    CDNET_KEEP_IN_REG64(L);  // keep the tile base in one register pair: one IMAD.WIDE per probe
#pragma unroll
    for (int it = 0; it < 30; ++it) {
        const double mid = __dadd_rn(lo, half);
        const int ry = __double2loint(__dadd_rn(__dadd_rn(fi, __dmul_rn(sy, mid)), kMagic));
        const int rx = __double2loint(__dadd_rn(__dadd_rn(fj, __dmul_rn(sx, mid)), kMagic));
        bool hit = false;
        if ((unsigned)ry < (unsigned)H && (unsigned)rx < (unsigned)W) hit = __ldg(L + (ry * W + rx)) == own;
        lo = hit ? mid : lo;
        half = half * 0.5;  // exact; folds to a constant per unrolled step
    }
    return __dadd_rn(lo, 0x1.f4p-21);  // + 1000 / 2^30
}

// Rays 0, 2, 4, 6 run along an image axis: (sin, cos) is (0 | +-1, +-1 | ~1e-16), the product with mid is +-mid
// exactly on the axis and below 2e-13 across it, which rounds away -- the cross coordinate stays put.
// Probe index = r * stride + off with r = round(c +- mid) in [0, limit).
__device__ __forceinline__ double ray_reach_axis(const int* __restrict__ L, int limit, int stride, int off, int c, int own,
                                                 long long sign) {
    const double kMagic = 6755399441055744.0;
    double lo = 0.0, half = 500.0;
    const double fc = (double)c;
    CDNET_KEEP_IN_REG64(L);
#pragma unroll
    for (int it = 0; it < 30; ++it) {
        const double mid = __dadd_rn(lo, half);
        const double step = __longlong_as_double(__double_as_longlong(mid) ^ sign);
        const int r = __double2loint(__dadd_rn(__dadd_rn(fc, step), kMagic));
        bool hit = false;
        if ((unsigned)r < (unsigned)limit) hit = __ldg(L + (r * stride + off)) == own;
        lo = hit ? mid : lo;
        half = half * 0.5;
    }
    return __dadd_rn(lo, 0x1.f4p-21);
}

// table layout per tile: entries 0..tab-1 (label ids).  A block owns a 64 x 32 pixel region, queues its
// instance pixels in shared memory and lets all 256 threads work through the queue, so that lanes are not
// idle on background (only ~25 % of a tile is nucleus).
constexpr int kCW = 64, kCH = 32;
__global__ void __launch_bounds__(256) k_t_centerness(const int* __restrict__ inst, double* __restrict__ cness,
                                                      unsigned long long* __restrict__ best, int tab, int H, int W) {
    __shared__ int s_q[kCW * kCH];
    __shared__ int s_n;
    const int b = blockIdx.z;
    const size_t tile = (size_t)b * H * W;
    const int* L = inst + tile;
    const int bx0 = blockIdx.x * kCW, by0 = blockIdx.y * kCH;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < kCW * kCH; i += 256) {
        const int x = bx0 + (i % kCW), y = by0 + (i / kCW);
        bool fg = false;
        if (x < W && y < H) {
            const int own = L[y * W + x];
            fg = own > 0 && own < tab;
        }
        // warp-aggregated append
        const unsigned m = __ballot_sync(0xffffffffu, fg);
        int base = 0;
        if ((threadIdx.x & 31) == 0 && m) base = atomicAdd(&s_n, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (fg) s_q[base + __popc(m & ((1u << (threadIdx.x & 31)) - 1))] = y * W + x;
    }
    __syncthreads();
    const int n = s_n;
    for (int it = threadIdx.x; it < n; it += 256) {
        const int p = s_q[it];
        const int y = p / W, x = p - y * W;
        const int own = L[p];
        double far = 0.0, near = 10000000.0;
#pragma unroll 1
        for (int a = 0; a < 4; ++a) {  // rays 0 (+x), 2 (+y), 4 (-x), 6 (-y)
            const bool vert = a & 1;
            const double r = ray_reach_axis(L, vert ? H : W, vert ? W : 1, vert ? x : y * W, vert ? y : x, own,
                                            (a & 2) ? (long long)0x8000000000000000ull : 0ll);
            far = fmax(far, r);
            near = fmin(near, r);
        }
#pragma unroll 1
        for (int k = 1; k < 8; k += 2) {
            const double r = ray_reach(L, H, W, y, x, own, c_rays[k][0], c_rays[k][1]);
            far = fmax(far, r);
            near = fmin(near, r);
        }
        const double c = __ddiv_rn(near, far);
        cness[tile + p] = c;
        atomicMax(best + (size_t)b * tab + own, (unsigned long long)__double_as_longlong(c));
    }
}

__global__ void __launch_bounds__(kBX* kBY) k_t_center_pick(const int* __restrict__ inst, const double* __restrict__ cness,
                                                            const unsigned long long* __restrict__ best,
                                                            int* __restrict__ centre, int tab, int H, int W) {
    PX_COORDS
    if (!inb) return;
    const int own = inst[tile + p];
    if (own <= 0 || own >= tab) return;
    if ((unsigned long long)__double_as_longlong(cness[tile + p]) == best[(size_t)b * tab + own])
        atomicMin(centre + (size_t)b * tab + own, p);  // strict '>' keeps the first raster maximum
}

__global__ void k_t_table_init(unsigned long long* __restrict__ best, int* __restrict__ centre, int* __restrict__ maxd2,
                               size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (best) best[i] = 0ull;
        centre[i] = 0x7fffffff;
        if (maxd2) maxd2[i] = 0;
    }
}

// centres -> [B, tab, 2] (row, col), (-1,-1) for absent ids
__global__ void k_t_centres_out(const int* __restrict__ centre, int* __restrict__ out, int W, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int c = centre[i];
        out[2 * i] = c == 0x7fffffff ? -1 : c / W;
        out[2 * i + 1] = c == 0x7fffffff ? -1 : c % W;
    }
}

// distinct non-zero labels in the cross neighbourhood of p (dilation(nucleus, disk(1)) support, :819)
__device__ __forceinline__ int cross_labels(const int* __restrict__ L, int H, int W, int y, int x, int out[5]) {
    const int p = y * W + x;
    int n = 0;
    auto add = [&](int v) {
        if (v <= 0) return;
        for (int i = 0; i < n; ++i)
            if (out[i] == v) return;
        out[n++] = v;
    };
    add(L[p]);
    if (y > 0) add(L[p - W]);
    if (y + 1 < H) add(L[p + W]);
    if (x > 0) add(L[p - 1]);
    if (x + 1 < W) add(L[p + 1]);
    return n;
}

// maxd2[k] = max over the dilated support of k of |q - c_k|^2 ; cflag marks the centres (label_point, :816)
__global__ void __launch_bounds__(kBX* kBY) k_t_support_max(const int* __restrict__ inst, const int* __restrict__ centre,
                                                            int* __restrict__ maxd2, uint8_t* __restrict__ cflag, int tab,
                                                            int H, int W) {
    PX_COORDS
    if (!inb) return;
    const int* L = inst + tile;
    int labs[5];
    const int n = cross_labels(L, H, W, y, x, labs);
    const int own = L[p];
    // the farthest support pixel from the centre is never an interior pixel of the nucleus (one of its four
    // neighbours is farther), so interior pixels skip the atomic
    const bool interior = n == 1 && labs[0] == own && y > 0 && y + 1 < H && x > 0 && x + 1 < W &&
                          L[p - W] == own && L[p + W] == own && L[p - 1] == own && L[p + 1] == own;
    uint8_t flag = 0;
    for (int i = 0; i < n; ++i) {
        const int k = labs[i];
        if (k >= tab) continue;
        const int c = centre[(size_t)b * tab + k];
        if (k == own && c == p) flag = 1;
        if (interior) continue;
        const int dy = y - c / W, dx = x - c % W;
        atomicMax(maxd2 + (size_t)b * tab + k, dy * dy + dx * dx);
    }
    cflag[tile + p] = flag;
}

__device__ __forceinline__ int align_index(float a, int n) {
    // SegFix_offset_helper.py:311-341 (upper-inclusive bins; all thresholds exact in f32 for n = 8, 16)
    const float step = 360.0f / (float)n, half = step * 0.5f;
    if (a <= -180.0f + half || a > 180.0f - half) return 0;
    for (int i = 1; i < n; ++i) {
        const float mid = -180.0f + step * (float)i;
        if (a > mid - half && a <= mid + half) return i;
    }
    return 0;
}

// defaults for the pixels no instance support covers: dir = (0, 0) -> angle 0 -> class index n/2 (+1) on
// foreground, 0 on background (:848-871); the per-label kernel overwrites the covered pixels
__global__ void __launch_bounds__(kBX* kBY) k_t_dir_default(const uint8_t* __restrict__ inside,
                                                            long long* __restrict__ direction, float* __restrict__ dir_out,
                                                            int n_classes, int H, int W) {
    PX_COORDS
    if (!inb) return;
    direction[tile + p] = inside[tile + p] ? (long long)(n_classes / 2 + 1) : 0ll;
    if (dir_out) { dir_out[(tile + p) * 2] = 0.0f; dir_out[(tile + p) * 2 + 1] = 0.0f; }
}

// compact list of the labels that exist: (tile << 32) | label
__global__ void k_t_label_list(const int* __restrict__ centre, unsigned long long* __restrict__ list, int* __restrict__ count,
                               int tab, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < ((n + 31) & ~size_t(31));
         i += (size_t)gridDim.x * blockDim.x) {
        const bool present = i < n && (i % tab) != 0 && centre[i] != 0x7fffffff;
        const unsigned m = __ballot_sync(0xffffffffu, present);
        int base = 0;
        if ((threadIdx.x & 31) == 0 && m) base = atomicAdd(count, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (present)
            list[base + __popc(m & ((1u << (threadIdx.x & 31)) - 1))] =
                ((unsigned long long)(i / tab) << 32) | (unsigned long long)(i % tab);
    }
}

constexpr int kLC = 32;              // output chunk edge
constexpr int kLV = kLC + 10;        // value tile edge (5-pixel tap halo)
constexpr int kLI = kLC + 12;        // label tile edge (+1 for the cross dilation)

// Direction map per LABEL (:819-834, :848-871).  "Last writer wins" means pixel p takes the Sobel response of
// the highest label w whose cross-dilated support contains p, and that response only sees w's own
// centre-distance map.  So a block takes one label k at a time (persistent, work-stealing over the label
// list): it stages the label tile, evaluates f32((1 - |q - c_k| / (M_k + 1e-7))) for the support pixels of k
// in shared memory (0 elsewhere) and every pixel whose winner is k runs the 121-tap sequential f32 FMA chain
// in (kh, kw) order straight from that tile -- unconditionally, because a zero operand leaves the chain
// bit-identical.  All support pixels of k lie within floor(sqrt(M_k^2)) of its centre, which bounds the window.
__global__ void __launch_bounds__(128) k_t_direction_lab(const int* __restrict__ inst, const int* __restrict__ centre,
                                                         const int* __restrict__ maxd2, const uint8_t* __restrict__ inside,
                                                         long long* __restrict__ direction, float* __restrict__ dir_out,
                                                         const unsigned long long* __restrict__ list,
                                                         const int* __restrict__ count, int* __restrict__ cursor, int tab,
                                                         int n_classes, int H, int W) {
    __shared__ int s_l[kLI][kLI + 1];
    __shared__ float s_v[kLV][kLV + 1];
    __shared__ int s_item;
    const int tid = threadIdx.x;
    const int total = *count;
    for (;;) {
        if (tid == 0) s_item = atomicAdd(cursor, 1);
        __syncthreads();
        const int item = s_item;
        __syncthreads();
        if (item >= total) break;
        const unsigned long long e = list[item];
        const int b = (int)(e >> 32), k = (int)(e & 0xffffffffu);
        const size_t tile = (size_t)b * H * W;
        const int* L = inst + tile;
        const int c = centre[(size_t)b * tab + k];
        const int cy = c / W, cx = c % W;
        const int m2 = maxd2[(size_t)b * tab + k];
        const double denom = __dadd_rn(__dsqrt_rn((double)m2), 0.0000001);
        int R = (int)sqrt((double)m2);
        while ((long long)(R + 1) * (R + 1) <= (long long)m2) ++R;
        while ((long long)R * R > (long long)m2) --R;
        const int wy0 = max(cy - R, 0), wy1 = min(cy + R, H - 1), wx0 = max(cx - R, 0), wx1 = min(cx + R, W - 1);
        for (int oy = wy0; oy <= wy1; oy += kLC) {
            for (int ox = wx0; ox <= wx1; ox += kLC) {
                // label tile: rows oy-6 .. oy+kLC+5
                for (int i = tid; i < kLI * kLI; i += 128) {
                    const int ly = i / kLI, lx = i % kLI;
                    const int gy = oy - 6 + ly, gx = ox - 6 + lx;
                    s_l[ly][lx] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? L[gy * W + gx] : 0;
                }
                __syncthreads();
                // value tile: rows oy-5 .. oy+kLC+4  (label tile index = value tile index + 1)
                for (int i = tid; i < kLV * kLV; i += 128) {
                    const int vy = i / kLV, vx = i % kLV;
                    const int ly = vy + 1, lx = vx + 1;
                    const bool sup = s_l[ly][lx] == k || s_l[ly - 1][lx] == k || s_l[ly + 1][lx] == k ||
                                     s_l[ly][lx - 1] == k || s_l[ly][lx + 1] == k;
                    float v = 0.0f;
                    if (sup) {
                        const int gy = oy - 5 + vy, gx = ox - 5 + vx;
                        // an out-of-image pixel can have an in-image neighbour with label k: it is not part of
                        // the image, hence not part of the support
                        if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
                            const int dy = gy - cy, dx = gx - cx;
                            v = __double2float_rn(
                                __dadd_rn(1.0, -__ddiv_rn(__dsqrt_rn((double)(dy * dy + dx * dx)), denom)));
                        }
                    }
                    s_v[vy][vx] = v;
                }
                __syncthreads();
                for (int i = tid; i < kLC * kLC; i += 128) {
                    const int ty = i / kLC, tx = i % kLC;
                    const int y = oy + ty, x = ox + tx;
                    if (y > wy1 || x > wx1) continue;
                    const int ly = ty + 6, lx = tx + 6;
                    const int w = max(max(s_l[ly][lx], max(s_l[ly - 1][lx], s_l[ly + 1][lx])),
                                      max(s_l[ly][lx - 1], s_l[ly][lx + 1]));
                    if (w != k) continue;
                    float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll 1
                    for (int kh = 0; kh < 11; ++kh) {
#pragma unroll
                        for (int kw = 0; kw < 11; ++kw) {
                            const float v = s_v[ty + kh][tx + kw];
                            acc0 = __fmaf_rn(c_sobel[0][kh * 11 + kw], v, acc0);
                            acc1 = __fmaf_rn(c_sobel[1][kh * 11 + kw], v, acc1);
                        }
                    }
                    const int p = y * W + x;
                    if (dir_out) {
                        dir_out[(tile + p) * 2] = acc0;
                        dir_out[(tile + p) * 2 + 1] = acc1;
                    }
                    long long cls = 0;
                    if (inside[tile + p]) {
                        // angle = degrees(arctan2(dir0, dir1)) in f32 (:848); the reference's libm/SVML atan2f is
                        // not correctly rounded, this is (f64 atan2 rounded to f32) -- differences are confined to
                        // a few ulp of the angle, i.e. to pixels within ~1e-5 degrees of a bin edge
                        const float ang = __fmul_rn(__double2float_rn(atan2((double)acc0, (double)acc1)), 57.295776f);
                        cls = align_index(ang, n_classes) + 1;
                    }
                    direction[tile + p] = cls;
                }
                __syncthreads();
            }
        }
    }
}

// ---- Gaussian point map (:842) ----------------------------------------------------------------------
__device__ __forceinline__ int reflect_idx(int i, int n) {
    // scipy 'reflect': (d c b a | a b c d | d c b a)
    if (n == 1) return 0;
    const int period = 2 * n;
    i %= period;
    if (i < 0) i += period;
    return i < n ? i : period - 1 - i;
}

constexpr int kGX = 32, kGY = 16, kGR = 8;

// The point map is sparse (one pixel per nucleus) and every term of scipy's correlate1d is >= +0: adding a zero
// term never changes the sum, so elements whose 17-tap window holds no non-zero input are written as +0 without
// arithmetic and the others run the full tap sequence in scipy's order.  Presence is tracked as bit masks: one
// 32-bit word per window column (the window is exactly 32 rows), one 64-bit word per row of the first pass.
static_assert(kGY + 2 * kGR == 32 && kGX + 2 * kGR <= 64, "mask widths");
__global__ void __launch_bounds__(kGX* kGY) k_t_gauss(const uint8_t* __restrict__ cflag, __half* __restrict__ out, int H,
                                                      int W) {
    __shared__ uint32_t s_col[kGX + 2 * kGR];
    __shared__ unsigned long long s_row[kGY];
    __shared__ double s_v[kGY][kGX + 2 * kGR];
    const int b = blockIdx.z;
    const size_t tile = (size_t)b * H * W;
    const uint8_t* F = cflag + tile;
    const int bx0 = blockIdx.x * kGX, by0 = blockIdx.y * kGY;
    const int tid = threadIdx.y * kGX + threadIdx.x;
    if (tid < kGX + 2 * kGR) s_col[tid] = 0u;
    if (tid < kGY) s_row[tid] = 0ull;
    __syncthreads();
    int any = 0;
    constexpr int kWW = (kGX + 2 * kGR) / 4;  // window words per row
    if (by0 >= kGR && by0 + kGY + kGR <= H && bx0 >= kGR && bx0 + kGX + kGR <= W && (W & 3) == 0 &&
        ((uintptr_t)F & 3) == 0) {
        // window inside the image: one aligned 4-byte load per thread, no reflection
        if (tid < (kGY + 2 * kGR) * kWW) {
            const int ly = tid / kWW, wx = tid % kWW;
            const uint32_t w = __ldg((const uint32_t*)(F + (size_t)(by0 + ly - kGR) * W + (bx0 - kGR + 4 * wx)));
            if (w) {
                any = 1;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if ((w >> (8 * k)) & 0xffu) atomicOr(&s_col[4 * wx + k], 1u << ly);
            }
        }
    } else {
        for (int i = tid; i < (kGY + 2 * kGR) * (kGX + 2 * kGR); i += kGX * kGY) {
            const int ly = i / (kGX + 2 * kGR), lx = i % (kGX + 2 * kGR);
            const int gy = reflect_idx(by0 + ly - kGR, H), gx = reflect_idx(bx0 + lx - kGR, W);
            if (F[gy * W + gx]) {
                atomicOr(&s_col[lx], 1u << ly);
                any = 1;
            }
        }
    }
    const int x = bx0 + threadIdx.x, y = by0 + threadIdx.y;
    if (!__syncthreads_or(any)) {  // no centre in the halo window: the whole block is zero
        if (x < W && y < H) out[tile + (size_t)y * W + x] = __double2half(0.0);
        return;
    }
    // axis 0 first (scipy.ndimage.gaussian_filter iterates axes in order), symmetric summation order of
    // correlate1d: centre term, then (in[l-j] + in[l+j]) * w[8-j] for j = 8 .. 1
    for (int i = tid; i < kGY * (kGX + 2 * kGR); i += kGX * kGY) {
        const int ly = i / (kGX + 2 * kGR), lx = i % (kGX + 2 * kGR);
        const uint32_t win = (s_col[lx] >> ly) & 0x1ffffu;  // window rows ly .. ly+16, centre = bit 8
        double t = 0.0;
        if (win) {
            t = __dmul_rn((win >> 8) & 1u ? 255.0 : 0.0, c_gauss[8]);
#pragma unroll
            for (int j = 8; j >= 1; --j) {
                const double a = (win >> (8 - j)) & 1u ? 255.0 : 0.0, c = (win >> (8 + j)) & 1u ? 255.0 : 0.0;
                t = __dadd_rn(t, __dmul_rn(__dadd_rn(a, c), c_gauss[8 - j]));
            }
            if (t != 0.0) atomicOr(&s_row[ly], 1ull << lx);
        }
        s_v[ly][lx] = t;
    }
    __syncthreads();
    if (x >= W || y >= H) return;
    const int lx = threadIdx.x + kGR;
    double t = 0.0;
    if ((s_row[threadIdx.y] >> threadIdx.x) & 0x1ffffull) {  // columns lx-8 .. lx+8 of the first pass
        t = __dmul_rn(s_v[threadIdx.y][lx], c_gauss[8]);
#pragma unroll
        for (int j = 8; j >= 1; --j)
            t = __dadd_rn(t, __dmul_rn(__dadd_rn(s_v[threadIdx.y][lx - j], s_v[threadIdx.y][lx + j]), c_gauss[8 - j]));
    }
    out[tile + (size_t)y * W + x] = __double2half(t);
}

static void sobel_weights(float out[2][121]) {
    // SegFix_offset_helper.py:102-132: k[j,i] = (i_ or j_) / (i_^2 + j_^2) in f64, stored to f32;
    // channel 0 = sobel_y (row offset j_), channel 1 = sobel_x (column offset i_)
    for (int j = 0; j < 11; ++j)
        for (int i = 0; i < 11; ++i) {
            const int j_ = j - 5, i_ = i - 5;
            float ky = 0.f, kx = 0.f;
            if (i_ != 0 || j_ != 0) {
                const double r2 = (double)(i_ * i_ + j_ * j_);
                ky = (float)((double)j_ / r2);
                kx = (float)((double)i_ / r2);
            }
            out[0][j * 11 + i] = ky;
            out[1][j * 11 + i] = kx;
        }
}

// ---- centres of my_transforms.LabelEncoding with do_direction = 1 (my_transforms.py:771-778) -------------------------
// `peak_local_max(distance_transform_edt(nucleus), exclude_border=0, num_peaks=1)`: the maximum of the nucleus's OWN
// distance transform (distance to the nearest pixel that is not part of this instance -- background or another
// nucleus -- with no implicit background outside the frame), first in raster order among equals.  One separable,
// exact pass for all instances at once: a pixel of another instance is a zero for this one.
__global__ void __launch_bounds__(kBX* kBY) k_ledt_cols(const int* __restrict__ inst, int* __restrict__ g2, int H, int W) {
    PX_COORDS
    if (!inb) return;
    const int* L = inst + tile;
    const int own = L[p];
    int r = 0;
    if (own > 0) {
        r = 1 << 30;
        const int kmax = max(y, H - 1 - y);
        for (int k = 1; k <= kmax; ++k) {
            const bool up = (y - k >= 0) && L[p - k * W] != own;
            const bool dn = (y + k < H) && L[p + k * W] != own;
            if (up || dn) { r = k * k; break; }
        }
    }
    g2[tile + p] = r;
}

__global__ void __launch_bounds__(kBX* kBY) k_ledt_rows(const int* __restrict__ inst, const int* __restrict__ g2,
                                                        double* __restrict__ cness, unsigned long long* __restrict__ best,
                                                        int tab, int H, int W) {
    PX_COORDS
    if (!inb) return;
    const int* L = inst + tile + (size_t)y * W;
    const int* G = g2 + tile + (size_t)y * W;
    const int own = L[x];
    if (own <= 0 || own >= tab) return;
    const int kInf = 1 << 30;
    int d2 = G[x];
    const int kmax = max(x, W - 1 - x);
    for (int k = 1; k <= kmax; ++k) {
        const int kk = k * k;
        if (kk >= d2) break;
        if (x - k >= 0) d2 = min(d2, kk + (L[x - k] == own ? min(G[x - k], kInf - kk) : 0));
        if (x + k < W) d2 = min(d2, kk + (L[x + k] == own ? min(G[x + k], kInf - kk) : 0));
    }
    const double c = (double)d2;  // sqrt is monotone: the maximum of d is the maximum of d^2
    cness[tile + p] = c;
    atomicMax(best + (size_t)b * tab + own, (unsigned long long)__double_as_longlong(c));
}

static int centres_edt_launch(const int32_t* inst, int32_t* g2, double* cness, unsigned long long* best, int32_t* centre,
                              int32_t* maxd2, int tab, int B, int H, int W, cudaStream_t st) {
    const size_t nt = (size_t)B * tab;
    const size_t blocks = (nt + 255) / 256;
    CDNET_LAUNCH(k_t_table_init, (unsigned)(blocks > 65535 ? 65535 : blocks), 256, 0, st, best, centre, maxd2, nt);
    CDNET_LAUNCH(k_ledt_cols, px_grid(B, H, W), px_block(), 0, st, inst, g2, H, W);
    CDNET_LAUNCH(k_ledt_rows, px_grid(B, H, W), px_block(), 0, st, inst, g2, cness, best, tab, H, W);
    CDNET_LAUNCH(k_t_center_pick, px_grid(B, H, W), px_block(), 0, st, inst, cness, best, centre, tab, H, W);
    return last_error();
}

// new_label_inside of my_transforms.LabelEncoding: ids > 0 (instance ids, :717) or label > 127.5 ({0,255} label, :733)
__global__ void k_t_inside_plain(const uint8_t* __restrict__ ids, uint8_t* __restrict__ inside, size_t n, int thr) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        inside[i] = ids[i] > thr;
}

static int centres_launch(const int32_t* inst, double* cness, unsigned long long* best, int32_t* centre, int32_t* maxd2,
                          int tab, int B, int H, int W, cudaStream_t st) {
    const size_t nt = (size_t)B * tab;
    const size_t blocks = (nt + 255) / 256;
    CDNET_LAUNCH(k_t_table_init, (unsigned)(blocks > 65535 ? 65535 : blocks), 256, 0, st, best, centre, maxd2, nt);
    CDNET_LAUNCH(k_t_centerness, dim3(ceil_div(W, kCW), ceil_div(H, kCH), B), 256, 0, st, inst, cness, best, tab, H, W);
    CDNET_LAUNCH(k_t_center_pick, px_grid(B, H, W), px_block(), 0, st, inst, cness, best, centre, tab, H, W);
    return last_error();
}

}  // namespace cdnet

using namespace cdnet;

static bool bad_dims(int B, int H, int W) { return B <= 0 || H <= 0 || W <= 0 || (double)H * W >= 2147483648.0; }

extern "C" int cdnet_label_stats(const uint8_t* ids, int32_t* presence, int32_t* fg_count, int B, int H, int W,
                                 void* stream) {
    if (!ids || !presence || !fg_count || bad_dims(B, H, W)) return CDNET_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    CDNET_CUDA_OK(cudaMemsetAsync(presence, 0, sizeof(int32_t) * 256 * (size_t)B, st));
    CDNET_CUDA_OK(cudaMemsetAsync(fg_count, 0, sizeof(int32_t) * (size_t)B, st));
    const size_t plane = (size_t)H * W;
    int gx = (int)((plane + 256 * 16 - 1) / (256 * 16));
    CDNET_LAUNCH(k_label_stats, dim3(gx, B), 256, 0, st, ids, presence, fg_count, plane);
    return last_error();
}

extern "C" size_t cdnet_center_points_workspace_bytes(int B, int H, int W, int max_label) {
    if (bad_dims(B, H, W) || max_label < 0) return 0;
    const size_t n = (size_t)B * H * W, nt = (size_t)B * ((size_t)max_label + 1);
    return pad256(n * 8) + pad256(nt * 8) + pad256(nt * 4);
}

extern "C" int cdnet_center_points(const int32_t* labels, int32_t* centres, int B, int H, int W, int max_label, void* ws,
                                   size_t ws_bytes, void* stream) {
    if (!labels || !centres || bad_dims(B, H, W) || max_label < 0) return CDNET_E_BADARG;
    const int tab = max_label + 1;
    Arena ar(ws, ws_bytes);
    double* cness = ar.take<double>((size_t)B * H * W);
    unsigned long long* best = ar.take<unsigned long long>((size_t)B * tab);
    int32_t* centre = ar.take<int32_t>((size_t)B * tab);
    if (!ar.ok) return CDNET_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = centres_launch(labels, cness, best, centre, nullptr, tab, B, H, W, st);
    if (rc) return rc;
    const size_t nt = (size_t)B * tab;
    const size_t blocks = (nt + 255) / 256;
    CDNET_LAUNCH(k_t_centres_out, (unsigned)(blocks > 65535 ? 65535 : blocks), 256, 0, st, centre, centres, W, nt);
    return last_error();
}

extern "C" size_t cdnet_encode_targets_workspace_bytes(int B, int H, int W) {
    if (bad_dims(B, H, W)) return 0;
    const size_t n = (size_t)B * H * W, nt = (size_t)B * ((size_t)H * W / 2 + 2);
    return pad256(n * 8) + pad256(nt * 8) + 2 * pad256(nt * 4) + 2 * pad256(n * 4) + 3 * pad256(n) +
           pad256((size_t)B * 4) + pad256((size_t)B * 1024) + pad256((size_t)B * H * 4) + ws_process_workspace(B, H, W);
}

template <typename IdT>
static int encode_targets_impl(const IdT* ids, int instance_level, uint8_t* ternary, uint16_t* point, int64_t* direction,
                               int32_t* inst_out, float* dir_out, int32_t* status, int B, int H, int W, int num_classes,
                               const double* gauss_w, void* ws, size_t ws_bytes, void* stream) {
    constexpr bool kWide = sizeof(IdT) == 4;
    if (kWide && instance_level > 1) return CDNET_E_BADARG;  // int32 ids: the out_c = 3 transform only
    if (!ids || !ternary || !point || !direction || bad_dims(B, H, W)) return CDNET_E_BADARG;
    if (num_classes != 8 && num_classes != 16) return CDNET_E_BADARG;
    if (instance_level < 0 || instance_level > 5) return CDNET_E_BADARG;
    if (ws_bytes < cdnet_encode_targets_workspace_bytes(B, H, W)) return CDNET_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)B * H * W;
    // instance ids are handed out 1..M by a 4- or 8-connected labelling (watershed keeps the marker ids):
    // M <= ceil(H*W/2), so the per-label tables need H*W/2 + 2 entries per tile
    const int tab = (int)((size_t)H * W / 2 + 2);
    const size_t nt = (size_t)B * tab;
    Arena ar(ws, ws_bytes);
    double* cness = ar.take<double>(n);
    unsigned long long* best = ar.take<unsigned long long>(nt);
    int32_t* centre = ar.take<int32_t>(nt);
    int32_t* maxd2 = ar.take<int32_t>(nt);
    int32_t* inst_raw = ar.take<int32_t>(n);
    int32_t* inst_ws = ar.take<int32_t>(n);
    uint8_t* inside = ar.take<uint8_t>(n);
    uint8_t* interior = ar.take<uint8_t>(n);
    uint8_t* cflag = ar.take<uint8_t>(n);
    int32_t* fg = ar.take<int32_t>(B);
    int32_t* pres = ar.take<int32_t>((size_t)B * 256);
    int32_t* rowcnt = ar.take<int32_t>((size_t)B * H < 2 ? 2 : (size_t)B * H);  // [0], [1] double as list count / cursor
    if (!ar.ok) return CDNET_E_WORKSPACE;
    void* sub_ws = (char*)ws + ar.off;
    const size_t sub_bytes = ws_bytes - ar.off;
    int32_t* inst = inst_out ? inst_out : inst_ws;
    if (status) CDNET_CUDA_OK(cudaMemsetAsync(status, 0, sizeof(int32_t) * (size_t)B, st));

    // constants: uploaded once per device (and again only when the caller's Gaussian weights change) -- a
    // pageable-source cudaMemcpyToSymbolAsync on every call would serialise the stream with the host and rewrite
    // tables that kernels of another stream may be reading
    int dev_id = 0;
    CDNET_CUDA_OK(cudaGetDevice(&dev_id));
    double gw[17];
    if (gauss_w) {
        for (int i = 0; i < 17; ++i) gw[i] = gauss_w[i];
    } else {
        // scipy.ndimage._filters._gaussian_kernel1d(sigma=2, order=0, radius=8)
        double s = 0.0;
        for (int i = 0; i < 17; ++i) { const double t = (double)(i - 8); gw[i] = exp(-0.5 / 4.0 * t * t); s += gw[i]; }
        for (int i = 0; i < 17; ++i) gw[i] /= s;
    }
    int n_sm = 148;
    {
        static std::mutex mu;
        static TargetConsts consts[kMaxDevices];
        std::lock_guard<std::mutex> lock(mu);
        TargetConsts& c = consts[dev_id % kMaxDevices];
        if (!c.sobel_done) {
            float sob[2][121];
            sobel_weights(sob);
            CDNET_CUDA_OK(cudaMemcpyToSymbol(c_sobel, sob, sizeof sob));
            cudaDeviceGetAttribute(&c.n_sm, cudaDevAttrMultiProcessorCount, dev_id);
            if (c.n_sm <= 0) c.n_sm = 148;
            c.sobel_done = true;
        }
        if (!c.gauss_done || memcmp(c.gauss, gw, sizeof gw) != 0) {
            CDNET_CUDA_OK(cudaStreamSynchronize(st));  // nothing of this stream still reads the old weights
            CDNET_CUDA_OK(cudaMemcpyToSymbol(c_gauss, gw, sizeof gw));
            memcpy(c.gauss, gw, sizeof gw);
            c.gauss_done = true;
        }
        n_sm = c.n_sm;
    }

    CDNET_RANGE("cdnet_encode_targets");
    // 1. ternary label, fg mask, interior
    CDNET_CUDA_OK(cudaMemsetAsync(fg, 0, sizeof(int32_t) * (size_t)B, st));
    {
        const size_t plane = (size_t)H * W;
        int gx = (int)((plane + 256 * 16 - 1) / (256 * 16));
        if (kWide) {
            int* stats = pres;  // [B][4] in the (otherwise unused) presence table
            CDNET_LAUNCH(k_stats_init, ceil_div(B, 256), 256, 0, st, stats, B);
            CDNET_LAUNCH(k_label_stats_wide, dim3(gx, B), 256, 0, st, (const int*)ids, stats, plane);
            CDNET_LAUNCH(k_stats_to_fg, ceil_div(B, 256), 256, 0, st, stats, fg, B);
        } else {
            CDNET_LAUNCH(k_label_stats, dim3(gx, B), 256, 0, st, (const uint8_t*)ids, pres, fg, plane);
        }
    }
    if (instance_level >= 4) {
        // my_transforms.LabelEncoding (do_direction = 1, out_c = 3): its own ternary rule (csrc/training.cu) and
        // new_label_inside; modes 4 / 5 = instance ids / {0,255} label
        if constexpr (!kWide) {
            int rc0 = cdnet_ternary_label(ids, nullptr, instance_level == 4 ? 0 : 1, ternary, B, H, W, stream);
            if (rc0) return rc0;
            const size_t blocks = (n + 256 * 8 - 1) / (256 * 8);
            CDNET_LAUNCH(k_t_inside_plain, (unsigned)(blocks > 65535 ? 65535 : blocks), 256, 0, st, ids, inside, n,
                         instance_level == 4 ? 0 : 127);
        }
    } else {
        CDNET_LAUNCH(k_t_ternary<IdT>, px_grid(B, H, W), px_block(), 0, st, ids, fg, instance_level, ternary, inside, interior, H, W);
    }
    // 2. instances: process(interior*255, min_size=5) (:759) or measure.label (:773), then dilation disk(1)
    nvtx_mark("targets: instances (process / label + dilation)");
    int rc;
    if (instance_level == 1) {
        rc = ws_process_launch(interior, inst_raw, status, B, H, W, 5, 1, sub_ws, sub_bytes, st);
    } else {
        Arena sub(sub_ws, sub_bytes);
        int32_t* Lp = sub.take<int32_t>(n);
        int32_t* idmap = sub.take<int32_t>(n);
        if (!sub.ok) return CDNET_E_WORKSPACE;
        // out_c != 3 (:723-725, :734): the labelling itself is the instance map, nothing is dilated
        int32_t* dst = instance_level >= 2 ? inst : inst_raw;
        rc = CDNET_E_BADARG;
        if (instance_level == 2 || instance_level >= 4) {
            if constexpr (!kWide) rc = ccl_label_values_launch(ids, dst, nullptr, Lp, idmap, rowcnt, B, H, W, st);
        } else {
            rc = ccl_label_launch(interior, dst, nullptr, Lp, idmap, rowcnt, B, H, W, 8, st);
        }
    }
    if (rc) return rc;
    if (instance_level < 2) {
        rc = label_dilate_launch(inst_raw, inst, 4, B, H, W, 1, st);
        if (rc) return rc;
    } else {
        const size_t plane = (size_t)H * W;
        const size_t gx = (plane + 256 * 8 - 1) / (256 * 8);
        const dim3 grid((unsigned)(gx > 1024 ? 1024 : gx), B);
        CDNET_CUDA_OK(cudaMemsetAsync(fg, 0x7f, sizeof(int32_t) * (size_t)B, st));  // fg is free again: tile minima
        CDNET_LAUNCH(k_t_inst_min, grid, 256, 0, st, inst, fg, plane);
        CDNET_LAUNCH(k_t_drop_first, grid, 256, 0, st, inst, fg, plane);
    }
    // 3. centres, support maxima, direction classes, point map
    nvtx_mark("targets: centres, direction classes, point map");
    if (instance_level >= 4) rc = centres_edt_launch(inst, inst_raw, cness, best, centre, maxd2, tab, B, H, W, st);
    else rc = centres_launch(inst, cness, best, centre, maxd2, tab, B, H, W, st);
    if (rc) return rc;
    CDNET_LAUNCH(k_t_support_max, px_grid(B, H, W), px_block(), 0, st, inst, centre, maxd2, cflag, tab, H, W);
    {
        // label list in the (now dead) centerness plane; [0] of rowcnt = count, [1] = work-stealing cursor
        unsigned long long* lablist = (unsigned long long*)cness;
        CDNET_CUDA_OK(cudaMemsetAsync(rowcnt, 0, 2 * sizeof(int32_t), st));
        const size_t nblk = (nt + 255) / 256;
        CDNET_LAUNCH(k_t_label_list, (unsigned)(nblk > 65535 ? 65535 : nblk), 256, 0, st, centre, lablist, rowcnt, tab, nt);
        CDNET_LAUNCH(k_t_dir_default, px_grid(B, H, W), px_block(), 0, st, inside, (long long*)direction, dir_out,
                     num_classes, H, W);
        CDNET_LAUNCH(k_t_direction_lab, n_sm * 8, 128, 0, st, inst, centre, maxd2, inside, (long long*)direction, dir_out,
                     lablist, rowcnt, rowcnt + 1, tab, num_classes, H, W);
    }
    CDNET_LAUNCH(k_t_gauss, dim3(ceil_div(W, kGX), ceil_div(H, kGY), B), dim3(kGX, kGY), 0, st, cflag, (__half*)point, H, W);
    return last_error();
}

extern "C" int cdnet_encode_targets(const uint8_t* ids, int instance_level, uint8_t* ternary, uint16_t* point,
                                    int64_t* direction, int32_t* inst_out, float* dir_out, int32_t* status, int B, int H,
                                    int W, int num_classes, const double* gauss_w, void* ws, size_t ws_bytes,
                                    void* stream) {
    return encode_targets_impl<uint8_t>(ids, instance_level, ternary, point, direction, inst_out, dir_out, status, B, H, W,
                                        num_classes, gauss_w, ws, ws_bytes, stream);
}

// int32 instance ids (no uint8 wrap at 256): the out_c = 3 transform, instance_level 0 / 1
extern "C" int cdnet_encode_targets_i32(const int32_t* ids, int instance_level, uint8_t* ternary, uint16_t* point,
                                        int64_t* direction, int32_t* inst_out, float* dir_out, int32_t* status, int B,
                                        int H, int W, int num_classes, const double* gauss_w, void* ws, size_t ws_bytes,
                                        void* stream) {
    return encode_targets_impl<int32_t>(ids, instance_level, ternary, point, direction, inst_out, dir_out, status, B, H, W,
                                        num_classes, gauss_w, ws, ws_bytes, stream);
}

// int32 ids: n_distinct[b] = min(number of distinct values, 3) (the reference only asks `> 2`), fg_count[b] = non-zero
// pixels.  scratch: int32 [B,5]
extern "C" int cdnet_label_stats_i32(const int32_t* ids, int32_t* n_distinct, int32_t* fg_count, int32_t* scratch, int B,
                                     int H, int W, void* stream) {
    if (!ids || !n_distinct || !fg_count || !scratch || bad_dims(B, H, W)) return CDNET_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t plane = (size_t)H * W;
    int gx = (int)((plane + 256 * 16 - 1) / (256 * 16));
    int* stats = scratch;
    int* third = scratch + 4 * (size_t)B;
    CDNET_LAUNCH(k_stats_init, ceil_div(B, 256), 256, 0, st, stats, B);
    CDNET_CUDA_OK(cudaMemsetAsync(third, 0, sizeof(int) * (size_t)B, st));
    CDNET_LAUNCH(k_label_stats_wide, dim3(gx, B), 256, 0, st, ids, stats, plane);
    CDNET_LAUNCH(k_label_third_wide, dim3(gx, B), 256, 0, st, ids, stats, third, plane);
    CDNET_LAUNCH(k_stats_distinct, ceil_div(B, 256), 256, 0, st, stats, third, n_distinct, fg_count, B);
    return last_error();
}
