// handoff.cu -- device-resident hand-off from the CNN to the post-processing.
//
// Replaces, in ONE kernel over the raw network outputs of the 8 test-time-augmentation variants,
//   get_probmaps            test_dam.py:983-1013   softmax of the 3-class head, softmax of the direction head,
//                                                  direction[0] *= mask[0], first-maximum argmax
//   un-flip / un-rotate     test_dam.py:357-367, 426-441   np.flip / np.rot90(k=3) back to the original frame
//   averaging               test_dam.py:445-450    ((((((p0 + hf) + vf) + hvf) + r90) + r90_hf) + r90_vf) + r90_hvf) / 8
// which the reference runs as torch softmax on the GPU, .cpu().numpy(), eight numpy flips / rotations and seven
// array additions per map.  Output is exactly what cdnet_dam_postproc consumes: prob float32 [B,3,H,W],
// point float32 [B,1,H,W], dcm uint8 [B,8,H,W] -- nothing leaves the device in between.
//
// Roofline: HBM.  Algorithmic bytes per output pixel = 8 variants x (3 + 1 + C) float32 read + 16 + 8 written
// (C = 9: 440 B/px).  Each logit is read once, coalesced: the four rotated variants are read along THEIR rows
// (the original frame's columns) and transposed through a padded shared-memory tile.
#include <math.h>

#include "internal.h"

namespace cdnet {

struct TtaPtrs {
    const float* mask[8];   // [B,3,h_v,w_v]
    const float* point[8];  // [B,1,h_v,w_v]
    const float* dir[8];    // [B,C,h_v,w_v]
};

constexpr int kTile = 32, kRowsPerPass = 8, kPasses = kTile / kRowsPerPass;

struct PixelOut {
    float p0, p1, p2, pt;
    int cls;
};

// one pixel of one variant from its logits: the 3-class soft-max as torch's spatial soft-max kernel does it in
// float32 (channel maximum, channels summed in order, exp(x - max) / sum), then the reference's numpy steps.
//
// Direction head (:1008-1011): arg max_c q_c, q = softmax(dir), q_0 scaled by p0.  The soft-max's common positive
// factor 1/sum cannot change the winner, so no sum and no division: among the classes >= 1 the largest logit wins
// (the first one on ties, like np.argmax), and class 0 keeps the pixel unless exp(l* - max) > exp(l0 - max) * p0
// -- two exponentials instead of C exponentials and C IEEE divisions (406 -> ~200 thread instructions per
// (pixel, variant), profiles/r01_widening.md).  Results can differ from the reference's only where two scaled
// probabilities agree to float rounding.
template <int C>
__device__ __forceinline__ PixelOut eval_logits(float m0, float m1, float m2, float pt, const float (&l)[C]) {
    PixelOut o;
    const float mx = fmaxf(fmaxf(m0, m1), m2);
    const float e0 = expf(m0 - mx), e1 = expf(m1 - mx), e2 = expf(m2 - mx);
    const float s = __fadd_rn(__fadd_rn(e0, e1), e2);
    o.p0 = __fdiv_rn(e0, s);
    o.p1 = __fdiv_rn(e1, s);
    o.p2 = __fdiv_rn(e2, s);
    o.pt = pt;
    const float l0 = l[0];
    float ls = l[1];
    int arg = 1;
#pragma unroll
    for (int c = 2; c < C; ++c)
        if (l[c] > ls) { ls = l[c]; arg = c; }
    const float dmx = fmaxf(l0, ls);
    const float w0 = __fmul_rn(expf(l0 - dmx), o.p0);
    if (!(expf(ls - dmx) > w0)) arg = 0;
    o.cls = arg;
    return o;
}

// scalar path: channel c of the pixel lives at base + off + c * plane
template <int C>
__device__ __forceinline__ PixelOut eval_pixel(const float* __restrict__ mask, const float* __restrict__ point,
                                               const float* __restrict__ dir, size_t plane, size_t off) {
    const float* pm = mask + off;
    const float m0 = __ldg(pm);
    pm += plane;
    const float m1 = __ldg(pm);
    pm += plane;
    const float m2 = __ldg(pm);
    const float pt = __ldg(point + off);
    const float* pd = dir + off;
    float l[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {  // all loads first, then the compare chain
        l[c] = __ldg(pd);
        pd += plane;
    }
    return eval_logits<C>(m0, m1, m2, pt, l);
}

// vector path: four consecutive pixels of one source row, 128-bit loads; `rev` = the four pixels are wanted in
// reversed order (a flipped axis): pixel j is component 3 - j
template <int C>
__device__ __forceinline__ void eval_pixels4(const float* __restrict__ mask, const float* __restrict__ point,
                                             const float* __restrict__ dir, size_t plane, size_t off, bool rev,
                                             PixelOut (&o)[4]) {
    const float* pm = mask + off;
    const float4 m0 = __ldg((const float4*)pm);
    pm += plane;
    const float4 m1 = __ldg((const float4*)pm);
    pm += plane;
    const float4 m2 = __ldg((const float4*)pm);
    const float4 pt = __ldg((const float4*)(point + off));
    const float* pd = dir + off;
    float4 l4[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        l4[c] = __ldg((const float4*)pd);
        pd += plane;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float l[C];
#pragma unroll
        for (int c = 0; c < C; ++c) l[c] = j == 0 ? l4[c].x : (j == 1 ? l4[c].y : (j == 2 ? l4[c].z : l4[c].w));
        const float a0 = j == 0 ? m0.x : (j == 1 ? m0.y : (j == 2 ? m0.z : m0.w));
        const float a1 = j == 0 ? m1.x : (j == 1 ? m1.y : (j == 2 ? m1.z : m1.w));
        const float a2 = j == 0 ? m2.x : (j == 1 ? m2.y : (j == 2 ? m2.z : m2.w));
        const float ap = j == 0 ? pt.x : (j == 1 ? pt.y : (j == 2 ? pt.z : pt.w));
        const PixelOut r = eval_logits<C>(a0, a1, a2, ap, l);
        // component j of the loaded quad is pixel (rev ? 3 - j : j) of the caller's quad
        if (rev) o[3 - j] = r;
        else o[j] = r;
    }
}

template <int C>
__global__ void __launch_bounds__(kTile* kRowsPerPass) k_tta_merge(TtaPtrs P, float* __restrict__ prob_out,
                                                                    float* __restrict__ point_out,
                                                                    uint8_t* __restrict__ dcm_out, int H, int W,
                                                                    int n_var) {
    // staging tile of the rotated variants, [local y][local x]: filled by lanes that walk y (row stride
    // kTile + 1 words: conflict-free), consumed by lanes that walk x
    __shared__ float s_val[4][kTile][kTile + 1];
    __shared__ int s_cls[kTile][kTile + 1];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
    const size_t plane = (size_t)H * W;
    float acc[kPasses][4];
    const int x = x0 + tx;

    // ---- variants 0..3: flips only, read in place (a reversed row is still one coalesced segment) ----
#pragma unroll 1
    for (int v = 0; v < (n_var < 4 ? n_var : 4); ++v) {
        const float* mask = P.mask[v] + (size_t)b * 3 * plane;
        const float* point = P.point[v] + (size_t)b * plane;
        const float* dir = P.dir[v] + (size_t)b * C * plane;
#pragma unroll
        for (int r = 0; r < kPasses; ++r) {
            const int y = y0 + ty + r * kRowsPerPass;
            if (x < W && y < H) {
                const int sy = (v & 2) ? H - 1 - y : y, sx = (v & 1) ? W - 1 - x : x;
                const PixelOut o = eval_pixel<C>(mask, point, dir, plane, (size_t)sy * W + sx);
                if (v == 0) { acc[r][0] = o.p0; acc[r][1] = o.p1; acc[r][2] = o.p2; acc[r][3] = o.pt; }
                else {
                    acc[r][0] = __fadd_rn(acc[r][0], o.p0);
                    acc[r][1] = __fadd_rn(acc[r][1], o.p1);
                    acc[r][2] = __fadd_rn(acc[r][2], o.p2);
                    acc[r][3] = __fadd_rn(acc[r][3], o.pt);
                }
                dcm_out[((size_t)b * n_var + v) * plane + (size_t)y * W + x] = (uint8_t)o.cls;
            }
        }
    }

    // ---- variants 4..7: rotated by 90 degrees; their frame is [W rows, H cols] --------------------------
    // original (y, x) <- variant (row, col) = (v&2 ? x : W-1-x,  v&1 ? H-1-y : y): consecutive lanes walk the
    // variant's columns, i.e. the original frame's rows y; results are transposed through shared memory
#pragma unroll 1
    for (int v = 4; v < n_var; ++v) {
        const float* mask = P.mask[v] + (size_t)b * 3 * plane;
        const float* point = P.point[v] + (size_t)b * plane;
        const float* dir = P.dir[v] + (size_t)b * C * plane;
        __syncthreads();  // the previous variant's tile has been consumed
#pragma unroll
        for (int r = 0; r < kPasses; ++r) {
            const int lx = ty + r * kRowsPerPass;  // local x of the original frame
            const int oy = y0 + tx, ox = x0 + lx;  // lanes walk y
            if (oy < H && ox < W) {
                const int row = (v & 2) ? ox : W - 1 - ox, col = (v & 1) ? H - 1 - oy : oy;
                const PixelOut o = eval_pixel<C>(mask, point, dir, plane, (size_t)row * H + col);
                s_val[0][tx][lx] = o.p0;
                s_val[1][tx][lx] = o.p1;
                s_val[2][tx][lx] = o.p2;
                s_val[3][tx][lx] = o.pt;
                s_cls[tx][lx] = o.cls;
            }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kPasses; ++r) {
            const int ly = ty + r * kRowsPerPass;
            const int y = y0 + ly;
            if (x < W && y < H) {
                acc[r][0] = __fadd_rn(acc[r][0], s_val[0][ly][tx]);
                acc[r][1] = __fadd_rn(acc[r][1], s_val[1][ly][tx]);
                acc[r][2] = __fadd_rn(acc[r][2], s_val[2][ly][tx]);
                acc[r][3] = __fadd_rn(acc[r][3], s_val[3][ly][tx]);
                dcm_out[((size_t)b * n_var + v) * plane + (size_t)y * W + x] = (uint8_t)s_cls[ly][tx];
            }
        }
    }

    const float scale = n_var == 8 ? 0.125f : 1.0f;  // "/ 8" (:446,450) is exact; without TTA nothing is averaged
#pragma unroll
    for (int r = 0; r < kPasses; ++r) {
        const int y = y0 + ty + r * kRowsPerPass;
        if (x < W && y < H) {
            const size_t p = (size_t)y * W + x;
            prob_out[((size_t)b * 3 + 0) * plane + p] = __fmul_rn(acc[r][0], scale);
            prob_out[((size_t)b * 3 + 1) * plane + p] = __fmul_rn(acc[r][1], scale);
            prob_out[((size_t)b * 3 + 2) * plane + p] = __fmul_rn(acc[r][2], scale);
            point_out[(size_t)b * plane + p] = __fmul_rn(acc[r][3], scale);
        }
    }
}

// ---- 4 pixels per thread (H % 4 == 0, W % 4 == 0, 16-byte aligned tensors) ---------------------------------------
// 256 threads own a 32 x 32 output tile: thread (ty = t / 8, tx = t % 8) produces pixels (y0 + ty, x0 + 4 tx .. + 3).
// The flipped variants are read as one float4 of the (possibly reversed) source row; for the rotated variants the
// same thread grid walks the variant's rows instead: thread (r = t / 8, cg = t % 8) reads four consecutive source
// columns (= four consecutive original rows y0 + 4 cg .. + 3) of the source row that is original column x0 + r, and
// the quad is transposed through the shared-memory tile.
template <int C>
__global__ void __launch_bounds__(256) k_tta_merge4(TtaPtrs P, float* __restrict__ prob_out, float* __restrict__ point_out,
                                                    uint8_t* __restrict__ dcm_out, int H, int W, int n_var) {
    __shared__ float s_val[4][kTile][kTile + 1];
    __shared__ unsigned char s_cls[kTile][kTile + 4];
    const int t = threadIdx.x;
    const int tx = t & 7, ty = t >> 3;
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
    const size_t plane = (size_t)H * W;
    const int x = x0 + 4 * tx, y = y0 + ty;
    const bool mine = x < W && y < H;  // W % 4 == 0: the whole quad is inside or outside
    float acc[4][4];                   // [pixel of the quad][p0, p1, p2, point]
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f;

#pragma unroll 1
    for (int v = 0; v < (n_var < 4 ? n_var : 4); ++v) {
        if (mine) {
            const float* mask = P.mask[v] + (size_t)b * 3 * plane;
            const float* point = P.point[v] + (size_t)b * plane;
            const float* dir = P.dir[v] + (size_t)b * C * plane;
            const int sy = (v & 2) ? H - 1 - y : y;
            const bool rev = (v & 1) != 0;
            const int sx = rev ? W - 4 - x : x;  // the quad [W-4-x, W-1-x] holds pixels x+3 .. x in this order
            PixelOut o[4];
            eval_pixels4<C>(mask, point, dir, plane, (size_t)sy * W + sx, rev, o);
            unsigned cls = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // the first variant initialises the sums (an explicit 0 + p would turn a point value of -0.0 into +0.0)
                acc[j][0] = v == 0 ? o[j].p0 : __fadd_rn(acc[j][0], o[j].p0);
                acc[j][1] = v == 0 ? o[j].p1 : __fadd_rn(acc[j][1], o[j].p1);
                acc[j][2] = v == 0 ? o[j].p2 : __fadd_rn(acc[j][2], o[j].p2);
                acc[j][3] = v == 0 ? o[j].pt : __fadd_rn(acc[j][3], o[j].pt);
                cls |= (unsigned)o[j].cls << (8 * j);
            }
            *(unsigned*)(dcm_out + ((size_t)b * n_var + v) * plane + (size_t)y * W + x) = cls;
        }
    }

#pragma unroll 1
    for (int v = 4; v < n_var; ++v) {
        __syncthreads();  // the previous variant's tile has been consumed
        {
            const int cg = tx, r = ty;
            const int oy = y0 + 4 * cg, ox = x0 + r;  // original rows oy .. oy + 3, original column ox
            if (oy < H && ox < W) {                    // H % 4 == 0: the whole quad is inside or outside
                const float* mask = P.mask[v] + (size_t)b * 3 * plane;
                const float* point = P.point[v] + (size_t)b * plane;
                const float* dir = P.dir[v] + (size_t)b * C * plane;
                const int row = (v & 2) ? ox : W - 1 - ox;
                const bool rev = (v & 1) != 0;         // source column = H - 1 - y
                const int col = rev ? H - 4 - oy : oy;
                PixelOut o[4];
                eval_pixels4<C>(mask, point, dir, plane, (size_t)row * H + col, rev, o);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    s_val[0][4 * cg + j][r] = o[j].p0;
                    s_val[1][4 * cg + j][r] = o[j].p1;
                    s_val[2][4 * cg + j][r] = o[j].p2;
                    s_val[3][4 * cg + j][r] = o[j].pt;
                    s_cls[4 * cg + j][r] = (unsigned char)o[j].cls;
                }
            }
        }
        __syncthreads();
        if (mine) {
            unsigned cls = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc[j][0] = __fadd_rn(acc[j][0], s_val[0][ty][4 * tx + j]);
                acc[j][1] = __fadd_rn(acc[j][1], s_val[1][ty][4 * tx + j]);
                acc[j][2] = __fadd_rn(acc[j][2], s_val[2][ty][4 * tx + j]);
                acc[j][3] = __fadd_rn(acc[j][3], s_val[3][ty][4 * tx + j]);
                cls |= (unsigned)s_cls[ty][4 * tx + j] << (8 * j);
            }
            *(unsigned*)(dcm_out + ((size_t)b * n_var + v) * plane + (size_t)y * W + x) = cls;
        }
    }

    if (mine) {
        const float scale = n_var == 8 ? 0.125f : 1.0f;
        const size_t p = (size_t)y * W + x;
#pragma unroll
        for (int k = 0; k < 3; ++k)
            *(float4*)(prob_out + ((size_t)b * 3 + k) * plane + p) =
                make_float4(__fmul_rn(acc[0][k], scale), __fmul_rn(acc[1][k], scale), __fmul_rn(acc[2][k], scale),
                            __fmul_rn(acc[3][k], scale));
        *(float4*)(point_out + (size_t)b * plane + p) =
            make_float4(__fmul_rn(acc[0][3], scale), __fmul_rn(acc[1][3], scale), __fmul_rn(acc[2][3], scale),
                        __fmul_rn(acc[3][3], scale));
    }
}

}  // namespace cdnet

using namespace cdnet;

extern "C" int cdnet_tta_merge(const float* const* mask_logits, const float* const* point, const float* const* dir_logits,
                               int n_variants, float* prob_out, float* point_out, uint8_t* dcm_out, int B, int H, int W,
                               int dir_classes, void* stream) {
    if (!mask_logits || !point || !dir_logits || !prob_out || !point_out || !dcm_out) return CDNET_E_BADARG;
    if (B <= 0 || B > 65535 || H <= 0 || W <= 0 || (double)H * W >= 2147483648.0) return CDNET_E_BADARG;
    if (n_variants != 1 && n_variants != 8) return CDNET_E_BADARG;
    TtaPtrs P;
    for (int v = 0; v < 8; ++v) {
        const int u = v < n_variants ? v : 0;
        if (!mask_logits[u] || !point[u] || !dir_logits[u]) return CDNET_E_BADARG;
        P.mask[v] = mask_logits[u];
        P.point[v] = point[u];
        P.dir[v] = dir_logits[u];
    }
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid(ceil_div(W, kTile), ceil_div(H, kTile), B), block(kTile, kRowsPerPass);
    if (ceil_div(H, kTile) > 65535) return CDNET_E_BADARG;
    if (dir_classes != 5 && dir_classes != 9 && dir_classes != 17) return CDNET_E_BADARG;
    // 128-bit path: every row of every tensor must start on a 16-byte boundary in both frames
    bool vec = H % 4 == 0 && W % 4 == 0 && ((uintptr_t)prob_out | (uintptr_t)point_out) % 16 == 0 && (uintptr_t)dcm_out % 4 == 0;
    for (int v = 0; v < n_variants; ++v)
        vec = vec && ((uintptr_t)P.mask[v] | (uintptr_t)P.point[v] | (uintptr_t)P.dir[v]) % 16 == 0;
    if (vec) {
        if (dir_classes == 5) CDNET_LAUNCH(k_tta_merge4<5>, grid, 256, 0, st, P, prob_out, point_out, dcm_out, H, W, n_variants);
        else if (dir_classes == 9) CDNET_LAUNCH(k_tta_merge4<9>, grid, 256, 0, st, P, prob_out, point_out, dcm_out, H, W, n_variants);
        else CDNET_LAUNCH(k_tta_merge4<17>, grid, 256, 0, st, P, prob_out, point_out, dcm_out, H, W, n_variants);
    } else {
        if (dir_classes == 5) CDNET_LAUNCH(k_tta_merge<5>, grid, block, 0, st, P, prob_out, point_out, dcm_out, H, W, n_variants);
        else if (dir_classes == 9) CDNET_LAUNCH(k_tta_merge<9>, grid, block, 0, st, P, prob_out, point_out, dcm_out, H, W, n_variants);
        else CDNET_LAUNCH(k_tta_merge<17>, grid, block, 0, st, P, prob_out, point_out, dcm_out, H, W, n_variants);
    }
    return last_error();
}
