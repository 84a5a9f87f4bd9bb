// edt.cu -- exact Euclidean distance transform (K9).
//
// Replaces scipy.ndimage.distance_transform_edt (postproc_other.py:24,
// my_transforms_direction.py:802,822): distance of every non-zero pixel to the nearest zero pixel of
// the same image (no implicit background outside the frame), sqrt(float64(dy^2 + dx^2)).
//
// Separable and exact in int32: pass 1 walks each column for the nearest zero above/below
// (g, stored squared); pass 2 takes min over x' of (x-x')^2 + g2(y,x') and stops as soon as
// (x-x')^2 >= best, i.e. after O(d) steps -- nuclei are small, so this beats a full lower-envelope
// scan and every access of a warp is to 32 consecutive elements of one row.
// An image (column) without any zero pixel yields kInf (scipy's result is implementation-defined
// there; it cannot happen on the reference's path except for an all-foreground tile).
#include "internal.h"

namespace cdnet {

constexpr int kInf = kEdtInf;

__global__ void __launch_bounds__(256) k_edt_cols(const uint8_t* __restrict__ mask, int* __restrict__ g2, int H, int W) {
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
        for (int k = 1; k <= kmax; ++k) {
            const bool up = (y - k >= 0) && M[(y - k) * W + x] == 0;
            const bool dn = (y + k < H) && M[(y + k) * W + x] == 0;
            if (up || dn) { r = k * k; break; }
        }
    }
    g2[tile + (size_t)y * W + x] = r;
}

__global__ void __launch_bounds__(256) k_edt_rows(const int* __restrict__ g2, int* __restrict__ d2, int H, int W) {
    const int x = blockIdx.x * 64 + threadIdx.x;
    const int y = blockIdx.y * 4 + threadIdx.y;
    const int b = blockIdx.z;
    if (x >= W || y >= H) return;
    const size_t tile = (size_t)b * H * W;
    const int* G = g2 + tile + (size_t)y * W;
    int best = G[x];
    if (best > 1) {  // 0 (background) and 1 cannot improve
        const int kmax = max(x, W - 1 - x);
        for (int k = 1; k <= kmax; ++k) {
            const int kk = k * k;
            if (kk >= best) break;
            if (x - k >= 0) best = min(best, kk + min(G[x - k], kInf - kk));
            if (x + k < W) best = min(best, kk + min(G[x + k], kInf - kk));
        }
    }
    d2[tile + (size_t)y * W + x] = best;
}

__global__ void k_sqrt_f64(const int* __restrict__ d2, double* __restrict__ dist, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dist[i] = __dsqrt_rn((double)d2[i]);
}

int edt_launch(const uint8_t* mask, int32_t* d2, int32_t* g2, int B, int H, int W, cudaStream_t st) {
    dim3 block(64, 4), grid(ceil_div(W, 64), ceil_div(H, 4), B);
    CDNET_LAUNCH(k_edt_cols, grid, block, 0, st, mask, g2, H, W);
    CDNET_LAUNCH(k_edt_rows, grid, block, 0, st, g2, d2, H, W);
    return last_error();
}

}  // namespace cdnet

using namespace cdnet;

extern "C" size_t cdnet_edt_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0 || (double)H * W >= 2147483648.0) return 0;
    return pad256((size_t)B * H * W * 4);
}

extern "C" int cdnet_edt(const uint8_t* mask, int32_t* d2, double* dist, int B, int H, int W, void* ws, size_t ws_bytes,
                         void* stream) {
    if (!mask || !d2 || B <= 0 || H <= 0 || W <= 0 || (double)H * W >= 2147483648.0) return CDNET_E_BADARG;
    Arena ar(ws, ws_bytes);
    const size_t n = (size_t)B * H * W;
    int32_t* g2 = ar.take<int32_t>(n);
    if (!ar.ok) return CDNET_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = edt_launch(mask, d2, g2, B, H, W, st);
    if (rc) return rc;
    if (dist) {
        const size_t blocks = (n + 1023) / 1024;
        CDNET_LAUNCH(k_sqrt_f64, (unsigned)(blocks > (1u << 20) ? (1u << 20) : blocks), 256, 0, st, d2, dist, n);
    }
    return last_error();
}
