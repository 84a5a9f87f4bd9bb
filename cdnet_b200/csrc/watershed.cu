// watershed.cu -- postproc_other.process (K9-K12): instance distance map, markers, marker-controlled
// watershed, small-object removal.
//
// Replaces postproc_other.py:15-54 (the `ws` branch :36-49 and the unet/micronet head :50-54).
//
// O(H*W) reformulation of the reference's O(N_inst*H*W) loop (`gen_inst_dst_map`, :16-27): the EDT
// of one 4-connected component equals the global EDT of the binary mask restricted to it
// (SURVEY.md Appendix B.2), so ONE exact EDT plus a per-component maximum (atomicMax on the
// component root) gives uint8(255 * d / max d) for every instance at once.
//
// Watershed (skimage.segmentation.watershed, connectivity 1, compactness 0): a priority flood in
// (value, age) order that labels a pixel when it is PUSHED.  Values are uint8 (the reference negates
// a uint8 array, :47), so a 256-bucket FIFO per foreground component reproduces (value, age)
// exactly; ties between the initial age-0 marker pixels are resolved by raster index (this build's
// canonical order, SURVEY.md section 7 hard-part 1).  Floods of different components of `pred`
// never interact, so each component is flooded by one warp: lanes 0-3 fetch the four neighbours
// (-W, -1, +1, +W) in parallel, lane 0 owns the bucket heads/tails (shared memory) and the `next`
// links (global).
#include "internal.h"

namespace cdnet {

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

// L <- roots of pred's 4-conn components; maxd2[root] = max d2 over the component.  A squared distance of
// kEdtInf means the tile has no background pixel: the reference's gen_inst_dst_map raises there
// (postproc_other.py:18-19), reported as CDNET_S_NO_BACKGROUND.
__global__ void __launch_bounds__(kBX* kBY) k_comp_stats(const uint8_t* __restrict__ pred, int* __restrict__ L,
                                                         const int* __restrict__ d2, int* __restrict__ maxd2,
                                                         int32_t* __restrict__ status, int H, int W) {
    PX_COORDS
    int r = -1, d = 0;
    if (inb && pred[tile + p]) {
        int* Lt = L + tile;
        r = uf_find(Lt, p);
        Lt[p] = r;
        d = d2[tile + p];
    }
    // a warp covers 32 pixels of one row: mostly background, or the inside of ONE nucleus -- then one atomic serves all
    const unsigned act = __ballot_sync(0xffffffffu, r >= 0);
    if (!act) return;
    const int first = __ffs(act) - 1;
    const int r0 = __shfl_sync(0xffffffffu, r, first);
    if (__all_sync(0xffffffffu, r < 0 || r == r0)) {
        const int m = __reduce_max_sync(0xffffffffu, r >= 0 ? d : 0);
        if (lane == first) atomicMax(maxd2 + tile + r0, m);
    } else if (r >= 0) {
        atomicMax(maxd2 + tile + r, d);
    }
    if (r >= 0 && d >= kEdtInf && status && !(__ldcg(status + b) & CDNET_S_NO_BACKGROUND))
        atomicOr(status + b, CDNET_S_NO_BACKGROUND);
}

// four pixels per thread (W % 4 == 0, aligned planes); pixels of one quad that share a root share the atomic
__global__ void __launch_bounds__(256) k_comp_stats4(const uint8_t* __restrict__ pred, int* __restrict__ L,
                                                     const int* __restrict__ d2, int* __restrict__ maxd2,
                                                     int32_t* __restrict__ status, size_t plane, size_t nquads) {
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < nquads; q += (size_t)gridDim.x * blockDim.x) {
        const size_t i = 4 * q;
        const uint32_t pw = *(const uint32_t*)(pred + i);
        if (!pw) continue;
        const size_t b = i / plane, tile = b * plane;
        int* Lt = L + tile;
        const int p0 = (int)(i - tile);
        const int4 dv = *(const int4*)(d2 + i);
        const int ds[4] = {dv.x, dv.y, dv.z, dv.w};
        int4 lv = *(const int4*)(L + i);
        int rs[4] = {lv.x, lv.y, lv.z, lv.w};
        int cur = -1, m = 0;
        bool inf = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (!((pw >> (8 * k)) & 0xffu)) continue;
            int r = rs[k];                       // the pixel's parent: walk on from there
            if (r != p0 + k) r = uf_find(Lt, r);
            rs[k] = r;
            inf = inf || ds[k] >= kEdtInf;
            if (r != cur) {
                if (cur >= 0) atomicMax(maxd2 + tile + cur, m);
                cur = r;
                m = ds[k];
            } else {
                m = max(m, ds[k]);
            }
        }
        if (cur >= 0) atomicMax(maxd2 + tile + cur, m);
        *(int4*)(L + i) = make_int4(rs[0], rs[1], rs[2], rs[3]);  // background entries are written back unchanged
        if (inf && status && !(__ldcg(status + b) & CDNET_S_NO_BACKGROUND)) atomicOr(status + b, CDNET_S_NO_BACKGROUND);
    }
}

// dist = uint8(255 * (sqrt(d2) / sqrt(max d2)))  (postproc_other.py:24-26, f64, truncating);
// val = (uint8)(-dist) (:47); marker0 = dist > 125 (:39-41)
__global__ void __launch_bounds__(kBX* kBY) k_dist_marker(const uint8_t* __restrict__ pred, const int* __restrict__ L,
                                                          const int* __restrict__ d2, const int* __restrict__ maxd2,
                                                          uint8_t* __restrict__ val, uint8_t* __restrict__ marker0,
                                                          int H, int W) {
    PX_COORDS
    if (!inb) return;
    uint8_t dist = 0;
    if (pred[tile + p]) {
        const double d = __dsqrt_rn((double)d2[tile + p]);
        const double dm = __dsqrt_rn((double)maxd2[tile + L[tile + p]]);
        const double s = __dmul_rn(255.0, __ddiv_rn(d, dm));
        dist = (uint8_t)(int)s;
    }
    val[tile + p] = (uint8_t)(0u - (unsigned)dist);
    marker0[tile + p] = dist > 125;
}

// the same, four pixels per thread (W % 4 == 0, 16-byte aligned planes): 32- and 128-bit accesses, and a thread that sees
// four background pixels -- most of them -- touches neither the label nor the distance plane
__global__ void __launch_bounds__(256) k_dist_marker4(const uint8_t* __restrict__ pred, const int* __restrict__ L,
                                                      const int* __restrict__ d2, const int* __restrict__ maxd2,
                                                      uint8_t* __restrict__ val, uint8_t* __restrict__ marker0, size_t plane,
                                                      size_t nquads) {
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < nquads; q += (size_t)gridDim.x * blockDim.x) {
        const size_t i = 4 * q;
        const size_t tile = (i / plane) * plane;
        const uint32_t pw = *(const uint32_t*)(pred + i);
        uint32_t vw = 0, mw = 0;
        if (pw) {
            const int4 lv = *(const int4*)(L + i);
            const int4 dv = *(const int4*)(d2 + i);
            const int ls[4] = {lv.x, lv.y, lv.z, lv.w}, ds[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if ((pw >> (8 * k)) & 0xffu) {
                    const double d = __dsqrt_rn((double)ds[k]);
                    const double dm = __dsqrt_rn((double)maxd2[tile + ls[k]]);
                    const uint32_t dist = (uint32_t)(uint8_t)(int)__dmul_rn(255.0, __ddiv_rn(d, dm));
                    vw |= ((0u - dist) & 0xffu) << (8 * k);
                    mw |= (uint32_t)(dist > 125u) << (8 * k);
                }
            }
        }
        *(uint32_t*)(val + i) = vw;
        *(uint32_t*)(marker0 + i) = mw;
    }
}

// binary_erosion(iterations=1): cross structure, border_value 0 (postproc_other.py:43)
__global__ void __launch_bounds__(kBX* kBY) k_erode_cross(const uint8_t* __restrict__ state, uint8_t* __restrict__ out,
                                                          int H, int W) {
    PX_COORDS
    if (!inb) return;
    const uint8_t* S = state + tile;
    bool k = false;
    if (x > 0 && y > 0 && x + 1 < W && y + 1 < H) k = S[p] && S[p - 1] && S[p + 1] && S[p - W] && S[p + W];
    out[tile + p] = k;
}

// four pixels per thread (W % 4 == 0, aligned planes): the cross is an AND of five byte-words
__global__ void __launch_bounds__(256) k_erode_cross4(const uint8_t* __restrict__ state, uint8_t* __restrict__ out, int H,
                                                      int W, size_t nquads) {
    const int WQ = W >> 2;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < nquads; q += (size_t)gridDim.x * blockDim.x) {
        const int xq = (int)(q % WQ);
        const int y = (int)((q / WQ) % H);
        const size_t i = 4 * q;
        auto nz = [](uint32_t w) { return ((((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) >> 7) & 0x01010101u; };
        uint32_t r = 0;
        const uint32_t c = nz(*(const uint32_t*)(state + i));
        if (c && y > 0 && y + 1 < H) {
            const uint32_t up = nz(*(const uint32_t*)(state + i - W)), dn = nz(*(const uint32_t*)(state + i + W));
            const uint32_t lf = xq > 0 ? (uint32_t)(state[i - 1] != 0) : 0u;           // border_value 0
            const uint32_t rt = xq + 1 < WQ ? (uint32_t)(state[i + 4] != 0) : 0u;
            r = c & up & dn & ((c << 8) | lf) & ((c >> 8) | (rt << 24));
        }
        *(uint32_t*)(out + i) = r;
    }
}

// out <- markers * mask (skimage watershed's input validation); bounding box of every component of
// pred, keyed by its root
__global__ void __launch_bounds__(kBX* kBY) k_flood_prep(const uint8_t* __restrict__ pred, const int* __restrict__ L,
                                                         int* __restrict__ out, int* __restrict__ ymax,
                                                         int* __restrict__ xmin, int* __restrict__ xmax,
                                                         unsigned int* __restrict__ rootlist, int* __restrict__ nroots,
                                                         int* __restrict__ flag, const uint8_t* __restrict__ val,
                                                         int32_t* __restrict__ status, int H, int W) {
    PX_COORDS
    int r = -1;
    bool contested = false;
    if (inb) {
        if (pred[tile + p]) {
            r = L[tile + p];
            // Diagnostic for the one place where parity with scikit-image is unpinned: pixels that
            // two age-0 markers of EQUAL priority compete for.  skimage's heap pops such markers in an order that depends
            // on its internal heap mechanics; this build pops them in raster order.  A mask pixel without a marker is
            // counted when the smallest priority among its 4-neighbouring marker pixels is held by markers of two
            // different labels (first-order exposure: whichever of them is popped first labels the pixel).  The count
            // lands in status bits 8..31.  Marker pixels outside the mask are being zeroed by their own threads right
            // now, so a neighbour counts only where the mask is set (markers * mask).
            if (status && out[tile + p] == 0) {
                int best = 256, lab = 0;
                const int dy[4] = {-1, 0, 0, 1}, dx[4] = {0, -1, 1, 0};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int yy = y + dy[k], xx = x + dx[k];
                    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                    const int q = yy * W + xx;
                    if (!pred[tile + q]) continue;
                    const int l = out[tile + q];
                    if (l <= 0) continue;
                    const int v = val[tile + q];
                    if (v < best) { best = v; lab = l; contested = false; }
                    else if (v == best && l != lab) contested = true;
                }
            }
        } else {
            out[tile + p] = 0;
        }
    }
    {
        const unsigned mc = __ballot_sync(0xffffffffu, contested);
        if (lane == 0 && mc) atomicAdd(status + b, __popc(mc) << 8);
    }
    {
        // compact list of the components that hold a marker (the others stay unlabelled: nothing to flood) -- the flood
        // kernel hands them out to persistent warps.  The first marker pixel to raise its component's flag enters it.
        bool rep = false;
        if (r >= 0 && out[tile + p] > 0 && __ldcg(flag + tile + r) == 0) rep = atomicExch(flag + tile + r, 1) == 0;
        const unsigned m = __ballot_sync(0xffffffffu, rep);
        int basei = 0;
        if (lane == 0 && m) basei = atomicAdd(nroots, __popc(m));
        basei = __shfl_sync(0xffffffffu, basei, 0);
        if (rep) rootlist[basei + __popc(m & ((1u << lane) - 1))] = (unsigned int)(tile + r);
    }
    // bounding boxes: one set of atomics per (warp, component).  A warp is 32 pixels of one row -- mostly no component
    // at all, or one; only a mixed warp pays for the match
    const unsigned act = __ballot_sync(0xffffffffu, r >= 0);
    if (!act) return;
    const int r0 = __shfl_sync(0xffffffffu, r, __ffs(act) - 1);
    const unsigned peers = __all_sync(0xffffffffu, r < 0 || r == r0) ? act : __match_any_sync(0xffffffffu, r);
    if (r >= 0) {
        const int first = __ffs(peers) - 1, last = 31 - __clz(peers);
        if (lane == first) {
            atomicMin(xmin + tile + r, x);
            atomicMax(ymax + tile + r, y);
        }
        if (lane == last) atomicMax(xmax + tile + r, x);
    }
}

// rowmax[b][y] = largest marker id on row y.  Ids are handed out in raster order of the components' first pixels, so
// max(rowmax[0..y]) = number of marker components that start on rows 0..y
__global__ void __launch_bounds__(kBX* kBY) k_label_rowmax(const int* __restrict__ labels, int* __restrict__ rowmax,
                                                           int H, int W) {
    PX_COORDS
    const int l = inb ? labels[tile + p] : 0;
    const int m = __reduce_max_sync(0xffffffffu, l);  // a warp covers 32 pixels of one row (kBX = 128)
    if (lane == 0 && m > 0) atomicMax(rowmax + (size_t)b * H + y, m);
}

// Row-sharded callers hand in an extended tile and use the rows [own_lo, own_hi) only.  A 4-connected component of
// the mask that touches those rows must lie inside the tile completely, or its distance normalisation and markers
// are not the slide's: a component that reaches an outer row of the tile (0 when own_lo > 0, H-1 when own_hi < H)
// and the own rows sets CDNET_S_SHARD_OVERFLOW.  L holds flat roots = first raster pixel of the component.
// step 0: clear the marks of the components on row own_hi-1; 1: mark those on row H-1; 2: test (both seams).
__global__ void __launch_bounds__(256) k_shard_overflow(const uint8_t* __restrict__ pred, const int* __restrict__ L,
                                                        int* __restrict__ mark, int32_t* __restrict__ status, int H,
                                                        int W, int own_lo, int own_hi, int step) {
    const int x = blockIdx.x * 256 + threadIdx.x;
    const int b = blockIdx.y;
    if (x >= W) return;
    const size_t tile = (size_t)b * H * W;
    const int pb = (own_hi - 1) * W + x, pl = (H - 1) * W + x, pt = own_lo * W + x;
    bool bad = false;
    if (own_hi < H) {
        if (step == 0 && pred[tile + pb]) mark[tile + L[tile + pb]] = 0;
        if (step == 1 && pred[tile + pl]) mark[tile + L[tile + pl]] = 1;
        if (step == 2 && pred[tile + pb]) bad = mark[tile + L[tile + pb]] == 1;
    }
    if (step == 2 && own_lo > 0 && pred[tile + pt]) bad = bad || L[tile + pt] < W;  // root on row 0
    if (bad) atomicOr(status + b, CDNET_S_SHARD_OVERFLOW);
}

// one quad of k_flood_prep4
__device__ __forceinline__ void flood_prep_quad(const uint8_t* __restrict__ pred, const int* __restrict__ L,
                                                int* __restrict__ out, int* __restrict__ ymax, int* __restrict__ xmin,
                                                int* __restrict__ xmax, unsigned int* s_roots, int* s_n,
                                                int* __restrict__ flag, const uint8_t* __restrict__ val, int32_t* __restrict__ status, int H, int W,
                                                size_t plane, size_t q) {
    {
        const size_t i = 4 * q;
        const uint32_t pw = *(const uint32_t*)(pred + i);
        if (!pw) {
            *(int4*)(out + i) = make_int4(0, 0, 0, 0);
            return;
        }
        const size_t b = i / plane, tile = b * plane;
        const int p0 = (int)(i - tile), y = p0 / W, x0 = p0 - y * W;
        const int4 lv = *(const int4*)(L + i);
        const int4 ov4 = *(const int4*)(out + i);
        const int ls[4] = {lv.x, lv.y, lv.z, lv.w};
        int ov[4] = {ov4.x, ov4.y, ov4.z, ov4.w};
        const uint32_t left = x0 > 0 ? pred[i - 1] : 0u, right = x0 + 4 < W ? pred[i + 4] : 0u;
        const uint32_t below = y + 1 < H ? *(const uint32_t*)(pred + i + W) : 0u;
        int ncont = 0;
        bool need = false;  // a mask pixel without a marker
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (!((pw >> (8 * k)) & 0xffu)) {
                ov[k] = 0;  // markers * mask
                continue;
            }
            const int r = ls[k], x = x0 + k;
            // (a plain look first: after the component's first marker pixel the flag is up and no atomic is needed)
            if (ov[k] > 0 && __ldcg(flag + tile + r) == 0 && atomicExch(flag + tile + r, 1) == 0)
                s_roots[atomicAdd(s_n, 1)] = (unsigned int)(tile + r);
            const uint32_t lf = k ? ((pw >> (8 * (k - 1))) & 0xffu) : left;
            const uint32_t rt = k < 3 ? ((pw >> (8 * (k + 1))) & 0xffu) : right;
            if (!lf) atomicMin(xmin + tile + r, x);
            if (!rt) atomicMax(xmax + tile + r, x);
            if (!((below >> (8 * k)) & 0xffu)) atomicMax(ymax + tile + r, y);
            need = need || ov[k] == 0;
        }
        if (status && need) {
            // tie exposure (see k_flood_prep): the markers around the quad arrive as quads too, gated by the mask
            // (markers * mask); values are fetched only next to a marker
            const uint32_t above = y > 0 ? *(const uint32_t*)(pred + i - W) : 0u;
            const int4 zero4 = make_int4(0, 0, 0, 0);
            const int4 upo = above ? *(const int4*)(out + i - W) : zero4;
            const int4 dno = below ? *(const int4*)(out + i + W) : zero4;
            const int ups[4] = {upo.x, upo.y, upo.z, upo.w}, dns[4] = {dno.x, dno.y, dno.z, dno.w};
            const int os[4] = {ov4.x, ov4.y, ov4.z, ov4.w};
            const int lfo = left ? out[i - 1] : 0, rto = right ? out[i + 4] : 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!((pw >> (8 * k)) & 0xffu) || os[k] != 0) continue;
                int nl[4];
                nl[0] = ((above >> (8 * k)) & 0xffu) ? ups[k] : 0;
                nl[1] = k ? (((pw >> (8 * (k - 1))) & 0xffu) ? os[k > 0 ? k - 1 : 0] : 0) : lfo;
                nl[2] = k < 3 ? (((pw >> (8 * (k + 1))) & 0xffu) ? os[k < 3 ? k + 1 : 3] : 0) : rto;
                nl[3] = ((below >> (8 * k)) & 0xffu) ? dns[k] : 0;
                if (nl[0] <= 0 && nl[1] <= 0 && nl[2] <= 0 && nl[3] <= 0) continue;
                const ptrdiff_t off[4] = {-(ptrdiff_t)W, -1, 1, (ptrdiff_t)W};
                int best = 256, lab = 0;
                bool two = false;
#pragma unroll
                for (int d = 0; d < 4; ++d) {
                    if (nl[d] <= 0) continue;
                    const int v = val[(ptrdiff_t)(i + k) + off[d]];
                    if (v < best) { best = v; lab = nl[d]; two = false; }
                    else if (v == best && nl[d] != lab) two = true;
                }
                ncont += two;
            }
        }
        *(int4*)(out + i) = make_int4(ov[0], ov[1], ov[2], ov[3]);
        if (ncont) atomicAdd(status + b, ncont << 8);
    }
}

// k_flood_prep, four pixels per thread (W % 4 == 0, aligned planes).  Quads without a mask pixel -- most -- are one
// 128-bit store.  The bounding box needs atomics only where the component ends: x_min where the left neighbour is not
// mask (in a 4-connected labelling a mask neighbour is the same component), x_max / y_max likewise.
__global__ void __launch_bounds__(256) k_flood_prep4(const uint8_t* __restrict__ pred, const int* __restrict__ L,
                                                     int* __restrict__ out, int* __restrict__ ymax,
                                                     int* __restrict__ xmin, int* __restrict__ xmax,
                                                     unsigned int* __restrict__ rootlist, int* __restrict__ nroots,
                                                     int* __restrict__ flag, const uint8_t* __restrict__ val,
                                                     int32_t* __restrict__ status, int H, int W, size_t nquads) {
    // The compact list of the components with a marker (work items of k_flood) is filled through ONE counter: thousands of returning
    // atomics on one address serialise (~17 ns each: they WERE this kernel's run time), so a block collects its roots in
    // shared memory and reserves list space once per ~1000 roots.
    __shared__ unsigned int s_roots[2048];
    __shared__ int s_n, s_base;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const size_t plane = (size_t)H * W;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t iters = (nquads + stride - 1) / stride;  // the same for every thread: the flush below has barriers
    for (size_t it = 0; it < iters; ++it) {
        const size_t q = it * stride + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (q < nquads) flood_prep_quad(pred, L, out, ymax, xmin, xmax, s_roots, &s_n, flag, val, status, H, W, plane, q);
        __syncthreads();
        if (s_n > 1024 || it + 1 == iters) {  // at most 512 roots join per round (neighbours share a root)
            const int cnt = s_n;
            if (threadIdx.x == 0 && cnt) s_base = atomicAdd(nroots, cnt);
            __syncthreads();
            for (int j = threadIdx.x; j < cnt; j += blockDim.x) rootlist[s_base + j] = s_roots[j];
            __syncthreads();
            if (threadIdx.x == 0) s_n = 0;
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(kBX* kBY) k_bbox_init(int* __restrict__ ymax, int* __restrict__ xmin,
                                                        int* __restrict__ xmax, int H, int W) {
    PX_COORDS
    if (!inb) return;
    ymax[tile + p] = -1;
    xmin[tile + p] = 0x7fffffff;
    xmax[tile + p] = -1;
}

struct Buckets {
    int head[256];
    int tail[256];
};

// Components whose bounding box has at most kCap pixels (every single nucleus) are flooded entirely in
// shared memory: labels, values, FIFO links and bucket heads/tails of the box are staged per warp, so a
// queue step costs shared-memory latency instead of several dependent L2 round trips.  Larger components
// take the global-memory path with the same queue discipline.
constexpr int kCap = 1536;
struct WarpBox {
    int lab[kCap];              // -1 = not this component, 0 = unlabelled, > 0 = label
    unsigned short nxt[kCap];   // FIFO link (0xffff = end)
    unsigned char val[kCap];
    unsigned short head[256], tail[256];
};

// The box is staged with a one-pixel border of "not this component" cells, so the four neighbours of cell e are
// e - pw, e - 1, e + 1, e + pw without any bounds test or division (pw = box width + 2).
__device__ __forceinline__ void flood_box(WarpBox& S, const uint8_t* __restrict__ P, const int* __restrict__ Lt,
                                          const uint8_t* __restrict__ V, volatile int* O, int root, int W, int x0, int y0,
                                          int bw, int bh, int lane) {
    const int pw = bw + 2, ph = bh + 2;
    const int n = pw * ph;
    for (int i = lane; i < 256; i += 32) { S.head[i] = 0xffff; S.tail[i] = 0xffff; }
    // stage: all four planes are requested independently (no load waits on another load's result)
    int lmn = 0x7fffffff, lmx = 0;
    for (int ly = 0; ly < ph; ++ly) {
        const int gy = y0 + ly - 1;
        for (int lx = lane; lx < pw; lx += 32) {
            const int idx = ly * pw + lx;
            int lab = -1, v = 0;
            if (ly > 0 && ly <= bh && lx > 0 && lx <= bw) {
                const int g = gy * W + x0 + lx - 1;
                const int pm = P[g], lr = Lt[g], ov = O[g];
                v = V[g];
                if (pm && lr == root) lab = ov;
            }
            if (lab > 0) { lmn = min(lmn, lab); lmx = max(lmx, lab); }
            S.lab[idx] = lab;
            S.val[idx] = (unsigned char)v;
        }
    }
    // a component that holds a single marker label is simply filled with it (every pixel of a 4-connected
    // component is reached, and nobody competes); one without markers stays unlabelled.  Only components
    // with two or more markers -- touching nuclei -- need the ordered flood.
    __syncwarp();
    lmn = __reduce_min_sync(0xffffffffu, lmn);
    lmx = __reduce_max_sync(0xffffffffu, lmx);
    if (lmx == 0) return;
    if (lmn == lmx) {
        for (int ly = 1; ly <= bh; ++ly) {
            const int gy = y0 + ly - 1;
            for (int lx = 1 + lane; lx <= bw; lx += 32)
                if (S.lab[ly * pw + lx] == 0) O[gy * W + x0 + lx - 1] = lmn;
        }
        __syncwarp();
        return;
    }
    __syncwarp();
    int cur = 256;
    // age-0 elements: marker pixels in raster order
    for (int base = 0; base < n; base += 32) {
        const int idx = base + lane;
        const bool mk = idx < n && S.lab[idx] > 0;
        const int v = mk ? S.val[idx] : 0;
        unsigned m = __ballot_sync(0xffffffffu, mk);
        while (m) {
            const int l = __ffs(m) - 1;
            m &= m - 1;
            const int vv = __shfl_sync(0xffffffffu, v, l);
            if (lane == 0) {
                const int qq = base + l;
                S.nxt[qq] = 0xffff;
                if (S.tail[vv] == 0xffff) S.head[vv] = qq; else S.nxt[S.tail[vv]] = qq;
                S.tail[vv] = qq;
            }
            cur = min(cur, vv);
        }
    }
    __syncwarp();
    for (;;) {
        // next non-empty bucket: usually the current one
        int e = cur < 256 ? S.head[cur] : 0xffff;
        if (e == 0xffff) {
            int found = -1;
            for (int base = cur & ~31; base < 256; base += 32) {
                const int b = base + lane;
                const unsigned m = __ballot_sync(0xffffffffu, b >= cur && S.head[b] != 0xffff);
                if (m) { found = base + __ffs(m) - 1; break; }
            }
            if (found < 0) break;
            cur = found;
            e = S.head[cur];
        }
        const int lbl = S.lab[e];
        const int nx = S.nxt[e];
        // lanes 0..3 probe the neighbours in skimage's order (-W, -1, +1, +W)
        int q = -1, v = 0;
        if (lane < 4) {
            const int qq = e + (lane == 0 ? -pw : (lane == 1 ? -1 : (lane == 2 ? 1 : pw)));
            if (S.lab[qq] == 0) { q = qq; v = S.val[qq]; }
        }
        unsigned m = __ballot_sync(0xffffffffu, q >= 0);
        __syncwarp();  // every lane has read head[cur] / lab / nxt before lane 0 rewrites the queue
        if (lane == 0) {
            S.head[cur] = nx;
            if (nx == 0xffff) S.tail[cur] = 0xffff;
        }
        __syncwarp();
        while (m) {
            const int l = __ffs(m) - 1;
            m &= m - 1;
            const int qq = __shfl_sync(0xffffffffu, q, l);
            const int vv = __shfl_sync(0xffffffffu, v, l);
            if (lane == 0) {
                S.lab[qq] = lbl;
                S.nxt[qq] = 0xffff;
                if (S.tail[vv] == 0xffff) S.head[vv] = qq; else S.nxt[S.tail[vv]] = qq;
                S.tail[vv] = qq;
            }
            cur = min(cur, vv);
        }
        __syncwarp();
    }
    __syncwarp();
    for (int ly = 1; ly <= bh; ++ly) {
        const int gy = y0 + ly - 1;
        for (int lx = 1 + lane; lx <= bw; lx += 32) {
            const int l = S.lab[ly * pw + lx];
            if (l >= 0) O[gy * W + x0 + lx - 1] = l;
        }
    }
    __syncwarp();
}

// persistent warps: every warp pulls component roots from the compact list until it is exhausted
__global__ void __launch_bounds__(128) k_flood(const uint8_t* __restrict__ pred, const int* __restrict__ L,
                                               const uint8_t* __restrict__ val, volatile int* out, volatile int* next,
                                               const int* __restrict__ ymax, const int* __restrict__ xmin,
                                               const int* __restrict__ xmax, const unsigned int* __restrict__ rootlist,
                                               const int* __restrict__ nroots, int* __restrict__ cursor, int H, int W) {
    __shared__ WarpBox s_box[4];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int total = *nroots;
    const unsigned int plane = (unsigned int)H * (unsigned int)W;
    for (;;) {
        int wi = 0;
        if (lane == 0) wi = atomicAdd(cursor, 1);
        wi = __shfl_sync(0xffffffffu, wi, 0);
        if (wi >= total) break;
        const unsigned int g = rootlist[wi];
        const int b = (int)(g / plane);
        const int root = (int)(g - (unsigned int)b * plane);
        const size_t tile = (size_t)b * plane;
        const uint8_t* P = pred + tile;
        const int* Lt = L + tile;
        const uint8_t* V = val + tile;
        volatile int* O = out + tile;
        volatile int* N = next + tile;
        const int y0 = root / W, y1 = ymax[tile + root], x0 = xmin[tile + root], x1 = xmax[tile + root];
        if ((x1 - x0 + 3) * (y1 - y0 + 3) <= kCap) {
            flood_box(s_box[wid], P, Lt, V, O, root, W, x0, y0, x1 - x0 + 1, y1 - y0 + 1, lane);
            continue;
        }
        // ---- large component: same queue discipline on global memory (bucket heads/tails reuse the box)
        int* head = s_box[wid].lab;
        int* tail = s_box[wid].lab + 256;
        for (int i = lane; i < 256; i += 32) { head[i] = -1; tail[i] = -1; }
        __syncwarp();
        int cur = 256;
        for (int yy = y0; yy <= y1; ++yy) {
            for (int xb = x0; xb <= x1; xb += 32) {
                const int xx = xb + lane;
                const int q = yy * W + xx;
                const bool mk = xx <= x1 && P[q] && Lt[q] == root && O[q] != 0;
                const int v = mk ? V[q] : 0;
                unsigned m = __ballot_sync(0xffffffffu, mk);
                while (m) {
                    const int l = __ffs(m) - 1;
                    m &= m - 1;
                    const int qq = __shfl_sync(0xffffffffu, q, l);
                    const int vv = __shfl_sync(0xffffffffu, v, l);
                    if (lane == 0) {
                        N[qq] = -1;
                        if (tail[vv] < 0) head[vv] = qq; else N[tail[vv]] = qq;
                        tail[vv] = qq;
                    }
                    cur = min(cur, vv);
                }
            }
        }
        __syncwarp();
        for (;;) {
            int found = -1;
            for (int base = cur & ~31; base < 256; base += 32) {
                const int idx = base + lane;
                const unsigned m = __ballot_sync(0xffffffffu, idx >= cur && head[idx] >= 0);
                if (m) { found = base + __ffs(m) - 1; break; }
            }
            if (found < 0) break;
            cur = found;
            const int e = head[cur];
            // independent loads first: the label of e, its FIFO link and the three planes of its neighbour
            const int lbl = O[e];
            const int nx = N[e];
            const int ey = e / W, ex = e - ey * W;
            int q = -1, v = 0;
            if (lane < 4) {
                const int qy = ey + (lane == 0 ? -1 : (lane == 3 ? 1 : 0));
                const int qx = ex + (lane == 1 ? -1 : (lane == 2 ? 1 : 0));
                if (qy >= 0 && qy < H && qx >= 0 && qx < W) {
                    const int qq = qy * W + qx;
                    const int pm = P[qq], om = O[qq], vm = V[qq];
                    if (pm && om == 0) { q = qq; v = vm; }
                }
            }
            __syncwarp();
            if (lane == 0) {
                head[cur] = nx;
                if (nx < 0) tail[cur] = -1;
            }
            unsigned m = __ballot_sync(0xffffffffu, q >= 0);
            __syncwarp();
            while (m) {
                const int l = __ffs(m) - 1;
                m &= m - 1;
                const int qq = __shfl_sync(0xffffffffu, q, l);
                const int vv = __shfl_sync(0xffffffffu, v, l);
                if (lane == 0) {
                    O[qq] = lbl;
                    N[qq] = -1;
                    if (tail[vv] < 0) head[vv] = qq; else N[tail[vv]] = qq;
                    tail[vv] = qq;
                }
                cur = min(cur, vv);
            }
            __syncwarp();
        }
        __syncwarp();
    }
}

size_t ws_process_workspace(int B, int H, int W) {
    const size_t n = (size_t)B * H * W;
    return 5 * pad256(n * 4) + pad256((size_t)B * ((size_t)H * W + 1) * 4) + 3 * pad256(n) + pad256((size_t)B * H * 4);
}

// kernels from ccl.cu used here
int ccl_forest_launch(const uint8_t* mask, int32_t* L, int B, int H, int W, int conn, cudaStream_t st);

__global__ void k_state_mask(const uint8_t* __restrict__ state, uint8_t* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = state[i] ? 1 : 0;
}

int ws_process_launch(const uint8_t* pred01, int32_t* labels, int32_t* status, int B, int H, int W, int min_size,
                      int ws_flag, void* ws, size_t ws_bytes, cudaStream_t st, int32_t* marker_rowmax, int own_lo,
                      int own_hi) {
    const size_t n = (size_t)B * H * W;
    if (n >= 4294967296ull) return CDNET_E_BADARG;  // root list holds 32-bit batch-global pixel indices
    CDNET_RANGE("process(): EDT, markers, watershed, remove small");
    Arena ar(ws, ws_bytes);
    int32_t* A = ar.take<int32_t>(n);    // forest of pred
    int32_t* Bp = ar.take<int32_t>(n);   // g2 -> touch / idmap -> ymax
    int32_t* C = ar.take<int32_t>(n);    // d2 -> marker forest -> xmin
    int32_t* D = ar.take<int32_t>(n);    // max d2 per root -> xmax
    int32_t* E = ar.take<int32_t>(n);    // next links
    int32_t* counts = ar.take<int32_t>((size_t)B * ((size_t)H * W + 1));
    uint8_t* val = ar.take<uint8_t>(n);
    uint8_t* mk = ar.take<uint8_t>(n);
    uint8_t* state = ar.take<uint8_t>(n);
    int32_t* rowcnt = ar.take<int32_t>((size_t)B * H < 2 ? 2 : (size_t)B * H);  // [0], [1] double as list count / cursor
    if (!ar.ok) return CDNET_E_WORKSPACE;
    int rc;
    if (!ws_flag) {
        // postproc_other.py:50-54: fill holes -> label -> remove small
        rc = fill_holes_state_launch(pred01, state, A, Bp, B, H, W, st);
        if (rc) return rc;
        const size_t blocks = (n + 1023) / 1024;
        CDNET_LAUNCH(k_state_mask, (unsigned)(blocks > (1u << 20) ? (1u << 20) : blocks), 256, 0, st, state, mk, n);
        rc = ccl_label_launch(mk, labels, nullptr, A, Bp, rowcnt, B, H, W, 4, st);
        if (rc) return rc;
        return remove_small_labels_launch(labels, counts, B, H, W, min_size, st);
    }
    // 1. components of pred (:37) and the exact EDT of the mask (:24 for all instances at once)
    rc = ccl_forest_launch(pred01, A, B, H, W, 4, st);
    if (rc) return rc;
    rc = edt_launch(pred01, C, Bp, rowcnt, B, H, W, st);  // rowcnt is free until the markers are labelled
    if (rc) return rc;
    CDNET_CUDA_OK(cudaMemsetAsync(D, 0, n * 4, st));
    if (W % 4 == 0 && (((uintptr_t)pred01) & 3) == 0) {
        const size_t nq = n / 4, blocks = (nq + 255) / 256;
        CDNET_LAUNCH(k_comp_stats4, (unsigned)(blocks > (1u << 20) ? (1u << 20) : blocks), 256, 0, st, pred01, A, C, D, status,
                     (size_t)H * W, nq);
    } else {
        CDNET_LAUNCH(k_comp_stats, px_grid(B, H, W), px_block(), 0, st, pred01, A, C, D, status, H, W);
    }
    if (marker_rowmax && status && (own_lo > 0 || own_hi < H))
        for (int step = 0; step < 3; ++step)  // E is free until the flood
            CDNET_LAUNCH(k_shard_overflow, dim3(ceil_div(W, 256), B), 256, 0, st, pred01, A, E, status, H, W, own_lo,
                         own_hi, step);
    // 2. uint8 distance, its negation, markers (:25-26, :39-41, :47)
    if (W % 4 == 0 && (((uintptr_t)pred01 | (uintptr_t)val | (uintptr_t)mk) & 3) == 0) {
        const size_t nq = n / 4, blocks = (nq + 255) / 256;
        CDNET_LAUNCH(k_dist_marker4, (unsigned)(blocks > (1u << 20) ? (1u << 20) : blocks), 256, 0, st, pred01, A, C, D, val, mk,
                     (size_t)H * W, nq);
    } else {
        CDNET_LAUNCH(k_dist_marker, px_grid(B, H, W), px_block(), 0, st, pred01, A, C, D, val, mk, H, W);
    }
    // 3. fill holes, cross erosion, label, remove small (:42-46).  Bp .. E are free here and lie back to back in the
    // workspace: room for the run-based kernels (node plane + bit-planes).  Tiles up to 1024 columns do all three steps
    // there -- the filled bit-plane of one chain is eroded as the next chain packs it; wider tiles fill and erode on the
    // pixel-parent kernels and only label on runs; tiny tiles, where the 256-byte padding of the run kernels' slices does
    // not fit, and CDNET_NO_RLE=1 stay on the pixel-parent kernels altogether
    {
        const size_t span = (size_t)((char*)E - (char*)Bp) + pad256(n * 4);
        const bool room = span >= rle_tail_workspace(B, H, W);
        if (room && rle_markers_supported(W)) {
            rc = rle_markers_launch(mk, labels, B, H, W, Bp, span, st);
        } else {
            rc = fill_holes_state_launch(mk, state, C, Bp, B, H, W, st);
            if (rc) return rc;
            if (W % 4 == 0 && (((uintptr_t)state | (uintptr_t)mk) & 3) == 0) {
                const size_t nq = n / 4, blocks = (nq + 255) / 256;
                CDNET_LAUNCH(k_erode_cross4, (unsigned)(blocks > (1u << 20) ? (1u << 20) : blocks), 256, 0, st, state, mk, H, W, nq);
            } else {
                CDNET_LAUNCH(k_erode_cross, px_grid(B, H, W), px_block(), 0, st, state, mk, H, W);
            }
            if (room && rle_tail_supported(0)) rc = rle_label4_launch(mk, labels, B, H, W, Bp, span, st);
            else rc = ccl_label_launch(mk, labels, nullptr, C, Bp, rowcnt, B, H, W, 4, st);
        }
    }
    if (rc) return rc;
    if (marker_rowmax) {
        // markers of all sizes: the small ones dropped next keep their ids reserved (:46 leaves gaps)
        CDNET_CUDA_OK(cudaMemsetAsync(marker_rowmax, 0, sizeof(int32_t) * (size_t)B * H, st));
        CDNET_LAUNCH(k_label_rowmax, px_grid(B, H, W), px_block(), 0, st, labels, marker_rowmax, H, W);
    }
    rc = remove_small_labels_launch(labels, counts, B, H, W, min_size, st);
    if (rc) return rc;
    // 4. flood (:47)
    CDNET_LAUNCH(k_bbox_init, px_grid(B, H, W), px_block(), 0, st, Bp, C, D, H, W);
    // root list lives in the `counts` buffer (free between the two remove-small passes); rowcnt[0] = number
    // of roots, rowcnt[1] = work-stealing cursor
    unsigned int* rootlist = (unsigned int*)counts;
    CDNET_CUDA_OK(cudaMemsetAsync(rowcnt, 0, 2 * sizeof(int32_t), st));
    CDNET_CUDA_OK(cudaMemsetAsync(E, 0, n * 4, st));  // "component is on the list" flags, keyed by root (free until the flood)
    if (W % 4 == 0 && (((uintptr_t)pred01) & 3) == 0) {
        const size_t nq = n / 4, blocks = (nq + 255) / 256;
        // a small grid: every block flushes its roots with one atomic at the end (plus one per ~1000 roots)
        CDNET_LAUNCH(k_flood_prep4, (unsigned)(blocks > 148u * 8 ? 148u * 8 : blocks), 256, 0, st, pred01, A, labels, Bp, C, D,
                     rootlist, rowcnt, E, val, status, H, W, nq);
    } else {
        CDNET_LAUNCH(k_flood_prep, px_grid(B, H, W), px_block(), 0, st, pred01, A, labels, Bp, C, D, rootlist, rowcnt, E, val,
                     status, H, W);
    }
    static int n_sm = 0;
    if (!n_sm) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        if (n_sm <= 0) n_sm = 148;
    }
    CDNET_LAUNCH(k_flood, dim3(n_sm * 4), 128, 0, st, pred01, A, val, labels, E, Bp, C, D, rootlist, rowcnt, rowcnt + 1, H,
                 W);
    rc = last_error();
    if (rc) return rc;
    // 5. remove small (:48)
    return remove_small_labels_launch(labels, counts, B, H, W, min_size, st);
}

}  // namespace cdnet
