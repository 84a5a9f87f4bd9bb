// morph.cu -- grey dilation of a label image by disk(r) (K13).
//
// Replaces skimage.morphology.dilation(labels, selem=disk(radius)) at test_dam.py:563, test.py:295,
// my_transforms_direction.py:760,774: maximum over the footprint x^2 + y^2 <= r^2 (r = 1: 5-pixel
// cross, r = 2: 13 pixels); scipy's `reflect` border is equivalent to ignoring out-of-image taps for
// these symmetric footprints.
#include "internal.h"

namespace cdnet {

template <int R, typename OUT>
__global__ void __launch_bounds__(256) k_label_dilate(const int* __restrict__ labels, OUT* __restrict__ out, int H, int W) {
    const int x = blockIdx.x * 64 + threadIdx.x;
    const int y = blockIdx.y * 4 + threadIdx.y;
    const int b = blockIdx.z;
    if (x >= W || y >= H) return;
    const size_t tile = (size_t)b * H * W;
    const int* Lb = labels + tile;
    int m = 0;  // labels are >= 0
#pragma unroll
    for (int dy = -R; dy <= R; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int dx = -R; dx <= R; ++dx) {
            if (dx * dx + dy * dy > R * R) continue;
            const int xx = x + dx;
            if (xx < 0 || xx >= W) continue;
            m = max(m, __ldg(Lb + yy * W + xx));
        }
    }
    out[tile + (size_t)y * W + x] = (OUT)m;
}

template <typename OUT>
static int dilate_dispatch(const int32_t* labels, OUT* out, int B, int H, int W, int radius, cudaStream_t st) {
    dim3 block(64, 4), grid(ceil_div(W, 64), ceil_div(H, 4), B);
    switch (radius) {
        case 0: CDNET_LAUNCH((k_label_dilate<0, OUT>), grid, block, 0, st, labels, out, H, W); break;
        case 1: CDNET_LAUNCH((k_label_dilate<1, OUT>), grid, block, 0, st, labels, out, H, W); break;
        case 2: CDNET_LAUNCH((k_label_dilate<2, OUT>), grid, block, 0, st, labels, out, H, W); break;
        case 3: CDNET_LAUNCH((k_label_dilate<3, OUT>), grid, block, 0, st, labels, out, H, W); break;
        case 4: CDNET_LAUNCH((k_label_dilate<4, OUT>), grid, block, 0, st, labels, out, H, W); break;
        default: return CDNET_E_BADARG;
    }
    return last_error();
}

int label_dilate_launch(const int32_t* labels, void* out, int out_elem_bytes, int B, int H, int W, int radius,
                        cudaStream_t st) {
    if (out_elem_bytes == 4) return dilate_dispatch<int32_t>(labels, (int32_t*)out, B, H, W, radius, st);
    if (out_elem_bytes == 8) return dilate_dispatch<long long>(labels, (long long*)out, B, H, W, radius, st);
    return CDNET_E_BADARG;
}

}  // namespace cdnet

extern "C" int cdnet_label_dilate(const int32_t* labels, void* out, int out_elem_bytes, int B, int H, int W, int radius,
                                  void* stream) {
    if (!labels || !out || B <= 0 || H <= 0 || W <= 0 || (const void*)labels == out) return CDNET_E_BADARG;
    return cdnet::label_dilate_launch(labels, out, out_elem_bytes, B, H, W, radius, (cudaStream_t)stream);
}
