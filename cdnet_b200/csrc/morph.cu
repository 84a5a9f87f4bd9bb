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

// disk(1) / disk(2), 4 pixels per thread (W % 4 == 0): five 128-bit row loads + a few edge scalars instead
// of 13 scalar loads per pixel; int64 output as two 128-bit stores.
template <int R, typename OUT>
__global__ void __launch_bounds__(256) k_label_dilate4(const int* __restrict__ labels, OUT* __restrict__ out, int H, int W) {
    const int x4 = (blockIdx.x * 64 + threadIdx.x) * 4;
    const int y = blockIdx.y * 4 + threadIdx.y;
    const int b = blockIdx.z;
    if (x4 >= W || y >= H) return;
    const size_t tile = (size_t)b * H * W;
    const int* Lb = labels + tile;
    int m[4] = {0, 0, 0, 0};
#pragma unroll
    for (int dy = -R; dy <= R; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= H) continue;
        const int* row = Lb + (size_t)yy * W + x4;
        const int4 c = __ldg((const int4*)row);
        const int reach = (R == 1) ? (dy == 0 ? 1 : 0) : (dy == 0 ? 2 : (dy == 1 || dy == -1 ? 1 : 0));
        int v[8] = {0, 0, c.x, c.y, c.z, c.w, 0, 0};  // columns x4-2 .. x4+5
        if (reach >= 1) {
            if (x4 > 0) v[1] = __ldg(row - 1);
            if (x4 + 4 < W) v[6] = __ldg(row + 4);
        }
        if (reach >= 2) {
            if (x4 > 1) v[0] = __ldg(row - 2);
            if (x4 + 5 < W) v[7] = __ldg(row + 5);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int r = v[i + 2];
            if (reach >= 1) r = max(r, max(v[i + 1], v[i + 3]));
            if (reach >= 2) r = max(r, max(v[i], v[i + 4]));
            m[i] = max(m[i], r);
        }
    }
    OUT* o = out + tile + (size_t)y * W + x4;
    if (sizeof(OUT) == 4) {
        *(int4*)o = make_int4(m[0], m[1], m[2], m[3]);
    } else {
        *(longlong2*)o = make_longlong2((long long)m[0], (long long)m[1]);
        *(longlong2*)(o + 2) = make_longlong2((long long)m[2], (long long)m[3]);
    }
}

template <typename OUT>
static int dilate_dispatch(const int32_t* labels, OUT* out, int B, int H, int W, int radius, cudaStream_t st) {
    dim3 block(64, 4), grid(ceil_div(W, 64), ceil_div(H, 4), B);
    if (W % 4 == 0 && ((uintptr_t)labels & 15) == 0 && ((uintptr_t)out & 15) == 0 && (radius == 1 || radius == 2)) {
        dim3 grid4(ceil_div(W, 256), ceil_div(H, 4), B);
        if (radius == 1) CDNET_LAUNCH((k_label_dilate4<1, OUT>), grid4, block, 0, st, labels, out, H, W);
        else CDNET_LAUNCH((k_label_dilate4<2, OUT>), grid4, block, 0, st, labels, out, H, W);
        return last_error();
    }
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
