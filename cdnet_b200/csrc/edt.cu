// edt.cu -- exact Euclidean distance transform (K9).
//
// Replaces scipy.ndimage.distance_transform_edt (postproc_other.py:24,
// my_transforms_direction.py:802,822): distance of every non-zero pixel to the nearest zero pixel of
// the same image (no implicit background outside the frame), sqrt(float64(dy^2 + dx^2)).
//
// Separable and exact in int32: pass 1 finds for each pixel the nearest zero of its column (g, stored squared);
// pass 2 takes min over x' of (x-x')^2 + g2(y,x').  Both passes are written for what the path feeds them -- nuclei,
// a few tens of pixels across -- and stay linear on anything else:
//   * FAST: one thread per pixel scans outwards and stops after O(d) steps (pass 1: first zero above / below; pass 2:
//     as soon as (x-x')^2 >= best).  Every access of a warp is to 32 consecutive elements of one row.  The scans are
//     cut at kFar steps; a pixel that is not finished by then is marked.
//   * FAR (tissue-scale blobs, all-foreground tiles -- marked pixels only): pass 1 re-does a marked column with two
//     sequential sweeps (O(H) for the whole column instead of O(H) per pixel); pass 2 gives every marked pixel a whole
//     warp whose lanes take the offsets k, k+32, .. of the same early-out scan (the work per pixel is still O(d), but
//     32-wide, and there are few such pixels per lane).
// An image (column) without any zero pixel yields kInf (scipy's result is implementation-defined
// there; it cannot happen on the reference's path except for an all-foreground tile).
#include "internal.h"

namespace cdnet {

constexpr int kInf = kEdtInf;
constexpr int kFar = 64;  // scan cut-off of the fast kernels

// g2 < 0 marks an unfinished pixel; colflag[b * W + x] != 0 a column that holds one
__global__ void __launch_bounds__(256) k_edt_cols(const uint8_t* __restrict__ mask, int* __restrict__ g2,
                                                  int* __restrict__ colflag, int H, int W) {
    const int x = blockIdx.x * 64 + threadIdx.x;
    const int y = blockIdx.y * 4 + threadIdx.y;
    const int b = blockIdx.z;
    if (x >= W || y >= H) return;
    const size_t tile = (size_t)b * H * W;
    const uint8_t* M = mask + tile;
    int r = 0;
    if (M[y * W + x]) {
        r = kInf;
        const int kmax = max(y, H - 1 - y);
        const int kcut = min(kmax, kFar);
        for (int k = 1; k <= kcut; ++k) {
            const bool up = (y - k >= 0) && M[(y - k) * W + x] == 0;
            const bool dn = (y + k < H) && M[(y + k) * W + x] == 0;
            if (up || dn) { r = k * k; break; }
        }
        if (r == kInf && kcut < kmax) {
            r = -1;
            colflag[(size_t)b * W + x] = 1;
        }
    }
    g2[tile + (size_t)y * W + x] = r;
}

// k_edt_cols, four pixels per thread (W % 4 == 0, 4-byte aligned mask): the four columns of a quad share every row load of
// the outward scan, and a quad without a foreground pixel is one 128-bit store
__global__ void __launch_bounds__(256) k_edt_cols4(const uint8_t* __restrict__ mask, int* __restrict__ g2,
                                                   int* __restrict__ colflag, int H, int W) {
    const int xq = blockIdx.x * 64 + threadIdx.x;
    const int y = blockIdx.y * 4 + threadIdx.y;
    const int b = blockIdx.z;
    const int x = 4 * xq;
    if (x >= W || y >= H) return;
    const size_t tile = (size_t)b * H * W;
    const uint8_t* M = mask + tile;
    auto nzb = [](uint32_t w) { return (((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) & 0x80808080u; };  // 0x80 per non-zero byte
    uint32_t todo = nzb(*(const uint32_t*)(M + (size_t)y * W + x));
    int r[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if ((todo >> (8 * k)) & 0x80u) r[k] = kInf;
    if (todo) {
        const int kmax = max(y, H - 1 - y);
        const int kcut = min(kmax, kFar);
        for (int k = 1; k <= kcut && todo; ++k) {
            uint32_t zero = 0;  // 0x80 per byte that is a zero pixel above or below
            if (y - k >= 0) zero |= ~nzb(*(const uint32_t*)(M + (size_t)(y - k) * W + x)) & 0x80808080u;
            if (y + k < H) zero |= ~nzb(*(const uint32_t*)(M + (size_t)(y + k) * W + x)) & 0x80808080u;
            const uint32_t hit = zero & todo;
            if (hit) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if ((hit >> (8 * j)) & 0x80u) r[j] = k * k;
                todo &= ~hit;
            }
        }
        if (todo && kcut < kmax) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if ((todo >> (8 * j)) & 0x80u) {
                    r[j] = -1;
                    colflag[(size_t)b * W + x + j] = 1;
                }
        }
    }
    *(int4*)(g2 + tile + (size_t)y * W + x) = make_int4(r[0], r[1], r[2], r[3]);
}

// one thread per marked column: distance to the last zero above (downward sweep), then to the next zero below
__global__ void __launch_bounds__(128) k_edt_cols_far(const uint8_t* __restrict__ mask, int* __restrict__ g2,
                                                      const int* __restrict__ colflag, int H, int W) {
    const int x = blockIdx.x * 128 + threadIdx.x;
    const int b = blockIdx.y;
    if (x >= W || !colflag[(size_t)b * W + x]) return;
    const size_t tile = (size_t)b * H * W;
    const uint8_t* M = mask + tile + x;
    int* G = g2 + tile + x;
    int last = -kInf;  // row of the last zero seen
    for (int y = 0; y < H; ++y) {
        if (M[(size_t)y * W] == 0) last = y;
        const long long d = (long long)y - last;
        G[(size_t)y * W] = d >= 32768 ? kInf : (int)(d * d);
    }
    last = kInf;
    for (int y = H - 1; y >= 0; --y) {
        if (M[(size_t)y * W] == 0) last = y;
        const long long d = (long long)last - y;
        const int up = G[(size_t)y * W];
        G[(size_t)y * W] = d >= 32768 ? up : min(up, (int)(d * d));
    }
}

// rowflag[b * H + y] != 0: row y holds a pixel that k_edt_rows_far has to finish
__global__ void __launch_bounds__(256) k_edt_rows(const int* __restrict__ g2, int* __restrict__ d2, int* __restrict__ rowflag,
                                                  int H, int W) {
    const int x = blockIdx.x * 64 + threadIdx.x;
    const int y = blockIdx.y * 4 + threadIdx.y;
    const int b = blockIdx.z;
    if (x >= W || y >= H) return;
    const size_t tile = (size_t)b * H * W;
    const int* G = g2 + tile + (size_t)y * W;
    int best = G[x];
    if (best > 1) {  // 0 (background) and 1 cannot improve
        const int kmax = max(x, W - 1 - x);
        const int kcut = min(kmax, kFar);
        int k = 1;
        for (; k <= kcut; ++k) {
            const int kk = k * k;
            if (kk >= best) break;
            if (x - k >= 0) best = min(best, kk + min(G[x - k], kInf - kk));
            if (x + k < W) best = min(best, kk + min(G[x + k], kInf - kk));
        }
        if (k > kcut && kcut < kmax && k * k < best) {  // unfinished: k_edt_rows_far completes it
            best = -1;
            rowflag[(size_t)b * H + y] = 1;
        }
    }
    d2[tile + (size_t)y * W + x] = best;
}

// one warp per 32 consecutive pixels of a row; the (rare) marked pixels are finished one after the other, 32 offsets
// at a time
__global__ void __launch_bounds__(256) k_edt_rows_far(const int* __restrict__ g2, int* __restrict__ d2,
                                                      const int* __restrict__ rowflag, int H, int W) {
    const int lane = threadIdx.x & 31;
    const int x0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * 32;
    const int b = blockIdx.z;
    if (x0 >= W) return;
    const size_t tile = (size_t)b * H * W;
    for (int y = blockIdx.y; y < H; y += gridDim.y) {
    if (!rowflag[(size_t)b * H + y]) continue;  // nothing left in this row (every row of a tile of nuclei)
    const int* G = g2 + tile + (size_t)y * W;
    int* D = d2 + tile + (size_t)y * W;
    const int mine = (x0 + lane < W) ? D[x0 + lane] : 0;
    unsigned todo = __ballot_sync(0xffffffffu, mine < 0);
    while (todo) {
        const int i = __ffs(todo) - 1;
        todo &= todo - 1;
        const int x = x0 + i;
        int best = G[x];
        const int kmax = max(x, W - 1 - x);
        for (int k0 = 1; k0 <= kmax; k0 += 32) {
            if ((long long)k0 * k0 >= best) break;  // no offset of this or any later round can improve (warp-uniform)
            const int k = k0 + lane;
            int cand = kInf;
            if (k <= kmax && k < 32768) {
                const int kk = k * k;
                if (x - k >= 0) cand = min(cand, kk + min(G[x - k], kInf - kk));
                if (x + k < W) cand = min(cand, kk + min(G[x + k], kInf - kk));
            }
            best = min(best, __reduce_min_sync(0xffffffffu, cand));
        }
        if (lane == 0) D[x] = best;
    }
    }
}

__global__ void k_sqrt_f64(const int* __restrict__ d2, double* __restrict__ dist, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dist[i] = __dsqrt_rn((double)d2[i]);
}

int edt_launch(const uint8_t* mask, int32_t* d2, int32_t* g2, int32_t* rowflag, int B, int H, int W, cudaStream_t st) {
    dim3 block(64, 4), grid(ceil_div(W, 64), ceil_div(H, 4), B);
    // the column flags borrow the head of d2, which pass 2 overwrites afterwards
    int* colflag = d2;
    CDNET_CUDA_OK(cudaMemsetAsync(colflag, 0, sizeof(int) * (size_t)B * W, st));
    if (W % 4 == 0 && (((uintptr_t)mask) & 3) == 0 && (((uintptr_t)g2) & 15) == 0)
        CDNET_LAUNCH(k_edt_cols4, dim3(ceil_div(W / 4, 64), ceil_div(H, 4), B), block, 0, st, mask, g2, colflag, H, W);
    else
        CDNET_LAUNCH(k_edt_cols, grid, block, 0, st, mask, g2, colflag, H, W);
    if (H > kFar + 1) CDNET_LAUNCH(k_edt_cols_far, dim3(ceil_div(W, 128), B), 128, 0, st, mask, g2, colflag, H, W);
    if (W > kFar + 1) CDNET_CUDA_OK(cudaMemsetAsync(rowflag, 0, sizeof(int) * (size_t)B * H, st));
    CDNET_LAUNCH(k_edt_rows, grid, block, 0, st, g2, d2, rowflag, H, W);
    if (W > kFar + 1) {
        // ~4 000 blocks that loop over the rows: most rows are skipped by their flag, and a block per row would spend the
        // launch on nothing (56 000 blocks for 14 tiles of 1000^2: 38 us)
        const int per_tile = ceil_div(W, 256) * B;
        int gy = 4096 / per_tile;
        gy = gy < 1 ? 1 : (gy > H ? H : gy);
        CDNET_LAUNCH(k_edt_rows_far, dim3(ceil_div(W, 256), gy, B), 256, 0, st, g2, d2, rowflag, H, W);
    }
    return last_error();
}

}  // namespace cdnet

using namespace cdnet;

extern "C" size_t cdnet_edt_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0 || (double)H * W >= 2147483648.0) return 0;
    return pad256((size_t)B * H * W * 4) + pad256((size_t)B * H * 4);
}

extern "C" int cdnet_edt(const uint8_t* mask, int32_t* d2, double* dist, int B, int H, int W, void* ws, size_t ws_bytes,
                         void* stream) {
    if (!mask || !d2 || B <= 0 || H <= 0 || W <= 0 || (double)H * W >= 2147483648.0) return CDNET_E_BADARG;
    Arena ar(ws, ws_bytes);
    const size_t n = (size_t)B * H * W;
    int32_t* g2 = ar.take<int32_t>(n);
    int32_t* rowflag = ar.take<int32_t>((size_t)B * H);
    if (!ar.ok) return CDNET_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = edt_launch(mask, d2, g2, rowflag, B, H, W, st);
    if (rc) return rc;
    if (dist) {
        const size_t blocks = (n + 1023) / 1024;
        CDNET_LAUNCH(k_sqrt_f64, (unsigned)(blocks > (1u << 20) ? (1u << 20) : blocks), 256, 0, st, d2, dist, n);
    }
    return last_error();
}
