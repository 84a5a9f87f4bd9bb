// postproc.cu -- inference post-processing pipelines (K3-K5 + orchestration).
//
// cdnet_dam_postproc replaces the inline block test_dam.py:455-563 (with its hard-wired switches
// dcm_combined = 1, voting_firt = 0, DDM_switch = 100, mseloss = 1, direction = 1);
// cdnet_plain_postproc replaces test.py:270-295.
//
// DAM chain per batch of tiles (every kernel is batched over B):
//   k_ddm_codes<8>    8 TTA class maps -> 2-bit DDM codes (ddm.cu)                 test_dam.py:479-487
//   k_point_max       per-tile max of the point map                                test_dam.py:530
//   k_boost_inside    mean of the 8 normalised DDMs (f64), point gate + cross dilation,
//                     boundary boost of prob[2] (f64 math stored to f32), 3-way argmax,
//                     inside = (argmax == 1)                                        test_dam.py:489,530-539
//   fill holes -> remove small (4-conn) -> 8-conn label (ccl.cu, one forest)        test_dam.py:546-561
//     or process() = watershed chain (watershed.cu) when postproc == 1             test_dam.py:559
//   k_label_dilate    disk(radius)                                                 test_dam.py:563
#include <math.h>
#include <stdlib.h>

#include "internal.h"

namespace cdnet {

__device__ __forceinline__ float ddm_value_f(uint32_t d, uint32_t f) {
    const int mn = (f & 1) ? 0 : ((f & 2) ? 1 : 2);
    const int mx = (f & 4) ? 2 : ((f & 2) ? 1 : 0);
    return __fdiv_rn((float)((int)d - mn), (float)(mx - mn));
}

__device__ __forceinline__ void block_max_to(unsigned int m, unsigned int* dst) {
    __shared__ unsigned int s_max[8];
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 8) {
        m = s_max[threadIdx.x];
        m = __reduce_max_sync(0xffu, m);
        if (threadIdx.x == 0) atomicMax(dst, m);
    }
}

__global__ void __launch_bounds__(256) k_point_max(const float* __restrict__ point, unsigned int* __restrict__ pmax,
                                                   size_t plane) {
    const int b = blockIdx.y;
    const float* P = point + (size_t)b * plane;
    unsigned int m = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (size_t)gridDim.x * blockDim.x)
        m = max(m, f32_to_ordered(P[i]));
    block_max_to(m, pmax + b);
}

// plane % 4 == 0 and 16-byte aligned planes: four independent 16-byte loads in flight per thread
__global__ void __launch_bounds__(256) k_point_max4(const float4* __restrict__ point, unsigned int* __restrict__ pmax,
                                                    size_t plane4) {
    const int b = blockIdx.y;
    const float4* P = point + (size_t)b * plane4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned int m = 0;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < plane4; i += 4 * stride) {
        const float4 a = __ldg(P + i), c = __ldg(P + i + stride), d = __ldg(P + i + 2 * stride), e = __ldg(P + i + 3 * stride);
        const unsigned int m0 = max(max(f32_to_ordered(a.x), f32_to_ordered(a.y)), max(f32_to_ordered(a.z), f32_to_ordered(a.w)));
        const unsigned int m1 = max(max(f32_to_ordered(c.x), f32_to_ordered(c.y)), max(f32_to_ordered(c.z), f32_to_ordered(c.w)));
        const unsigned int m2 = max(max(f32_to_ordered(d.x), f32_to_ordered(d.y)), max(f32_to_ordered(d.z), f32_to_ordered(d.w)));
        const unsigned int m3 = max(max(f32_to_ordered(e.x), f32_to_ordered(e.y)), max(f32_to_ordered(e.z), f32_to_ordered(e.w)));
        m = max(m, max(max(m0, m1), max(m2, m3)));
    }
    for (; i < plane4; i += stride) {
        const float4 a = __ldg(P + i);
        m = max(m, max(max(f32_to_ordered(a.x), f32_to_ordered(a.y)), max(f32_to_ordered(a.z), f32_to_ordered(a.w))));
    }
    block_max_to(m, pmax + b);
}

static void point_max_launch(const float* point, unsigned int* pmax, int B, size_t plane, cudaStream_t st) {
    if (plane % 4 == 0 && ((uintptr_t)point & 15) == 0) {
        const size_t p4 = plane / 4;
        size_t gx = (p4 + 256 * 4 - 1) / (256 * 4);
        if (gx > 65535) gx = 65535;
        CDNET_LAUNCH(k_point_max4, dim3((unsigned)gx, B), 256, 0, st, (const float4*)point, pmax, p4);
    } else {
        size_t gx = (plane + 256 * 16 - 1) / (256 * 16);
        if (gx > 65535) gx = 65535;
        CDNET_LAUNCH(k_point_max, dim3((unsigned)gx, B), 256, 0, st, point, pmax, plane);
    }
}

// one thread per pixel.  prob: [B,3,H,W]; writes inside u8; optionally prob[2] <- boosted (test_dam.py:536)
__global__ void __launch_bounds__(256) k_boost_inside(const uint16_t* __restrict__ codes, const uint32_t* __restrict__ flags,
                                                      const float* __restrict__ point, const unsigned int* __restrict__ pmax,
                                                      float* __restrict__ prob, uint8_t* __restrict__ inside,
                                                      int32_t* __restrict__ status, int H, int W, int write_prob,
                                                      int n_maps) {
    __shared__ float s_val[8][4];
    __shared__ int s_const;
    const int b = blockIdx.z;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if (tid == 0) s_const = 0;
    __syncthreads();
    if (tid < 32) {
        const int t = tid >> 2, d = tid & 3;
        const uint32_t f = (flags[b] >> (3 * t)) & 7u;
        s_val[t][d] = (d < 3 && t < n_maps) ? ddm_value_f(d, f) : 0.f;
        if (t < n_maps && d == 0 && (f == 0u || f == 1u || f == 2u || f == 4u)) s_const = 1;
    }
    __syncthreads();
    if (s_const && status && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) atomicOr(status + b, CDNET_S_DDM_CONSTANT);
    const int x = blockIdx.x * 64 + threadIdx.x;
    const int y = blockIdx.y * 4 + threadIdx.y;
    if (x >= W || y >= H) return;
    const size_t plane = (size_t)H * W;
    const int p = y * W + x;
    const uint32_t c = codes[(size_t)b * plane + p];
    double sum = 0.0;
#pragma unroll
    for (int t = 0; t < 8; ++t) sum = __dadd_rn(sum, (double)s_val[t][(c >> (2 * t)) & 3u]);
    const double ddm = n_maps == 8 ? __dmul_rn(sum, 0.125) : sum;  // np.mean over the 8 maps (exact: dyadic values)
    // point gate (test_dam.py:530-531): f32 divide by the global max, > 0.2 (f32), cross dilation
    const float* P = point + (size_t)b * plane;
    const float mx = ordered_to_f32(pmax[b]);
    bool g = __fdiv_rn(P[p], mx) > 0.2f;
    if (y > 0) g = g || (__fdiv_rn(P[p - W], mx) > 0.2f);
    if (y + 1 < H) g = g || (__fdiv_rn(P[p + W], mx) > 0.2f);
    if (x > 0) g = g || (__fdiv_rn(P[p - 1], mx) > 0.2f);
    if (x + 1 < W) g = g || (__fdiv_rn(P[p + 1], mx) > 0.2f);
    // enhanced_boundary = 2 * (ddm - ddm * gate)   (test_dam.py:532-534), f64
    const double eb = __dmul_rn(2.0, __dadd_rn(ddm, -__dmul_rn(ddm, g ? 1.0 : 0.0)));
    float* PR = prob + (size_t)b * 3 * plane;
    const float p0 = PR[p], p1 = PR[plane + p], p2 = PR[2 * plane + p];
    // prob[2] = (prob[2] + 0.5*eb) * (1 + eb): f64 arithmetic stored into the f32 array (test_dam.py:536)
    const float p2n = __double2float_rn(__dmul_rn(__dadd_rn((double)p2, __dmul_rn(0.5, eb)), __dadd_rn(1.0, eb)));
    if (write_prob) PR[2 * plane + p] = p2n;
    // np.argmax: first maximum wins, NaN counts as maximum
    int am = 0;
    float best = p0;
    if (p1 > best || (p1 != p1 && best == best)) { am = 1; best = p1; }
    if (p2n > best || (p2n != p2n && best == best)) { am = 2; }
    inside[(size_t)b * plane + p] = (am == 1) ? 1 : 0;
}

// ---- vectorised variant (W % 4 == 0): 4 pixels per thread, 128-bit loads --------------------------
// * the gate `f32(p / max) > 0.2f` is monotone in p for max > 0, so it is evaluated as `p >= thr` with
//   thr = the smallest float whose IEEE quotient exceeds 0.2f (found per block by a few nextafter steps);
// * the mean of the 8 normalised maps is an exact multiple of 1/16: two 256-entry integer LUTs over the
//   low / high byte of the code word give 16*mean;
// * where the boost is zero (gate open or DDM 0) prob[2] is unchanged bit for bit, so the f64 path only
//   runs on boundary pixels.
#ifndef CDNET_BOOST_ROWS
#define CDNET_BOOST_ROWS 4
#endif
constexpr int kBoostRows = CDNET_BOOST_ROWS;

__device__ __forceinline__ float gate_threshold(float mx) {
    float t = __fmul_rn(0.2f, mx);
    for (int i = 0; i < 8 && __fdiv_rn(t, mx) > 0.2f; ++i) t = nextafterf(t, -INFINITY);
    for (int i = 0; i < 16 && !(__fdiv_rn(t, mx) > 0.2f); ++i) t = nextafterf(t, INFINITY);
    return t;
}

// per-tile constants of k_boost_inside4, computed once per tile by k_boost_prep instead of once per block
struct BoostPrep {
    uint8_t lo[256], hi[256];
    float thr;
    int is_const, use_div;
};

template <bool PREP>
__global__ void __launch_bounds__(256) k_boost_inside4(const uint16_t* __restrict__ codes, const uint32_t* __restrict__ flags,
                                                       const float* __restrict__ point, const unsigned int* __restrict__ pmax,
                                                       float* __restrict__ prob, uint8_t* __restrict__ inside,
                                                       int32_t* __restrict__ status, int H, int W, int write_prob,
                                                       int n_maps, const BoostPrep* __restrict__ prep) {
    __shared__ uint8_t s_lo[256], s_hi[256];
    __shared__ float s_thr;
    __shared__ int s_const, s_div;
    pdl_wait();  // no-ops unless launched as a programmatic dependent (common.cuh)
    pdl_trigger();
    const int b = blockIdx.z;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if (PREP) {
        const BoostPrep& P0 = prep[b];
        s_lo[tid] = P0.lo[tid];
        s_hi[tid] = P0.hi[tid];
        if (tid == 0) { s_thr = P0.thr; s_const = P0.is_const; s_div = P0.use_div; }
    } else {
        // 2 * normalised value of code d for each map (0, 1, 2); constant maps flagged
        const uint32_t fl = flags[b];
        int bad = 0, lo = 0, hi = 0;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            if (t >= n_maps) continue;
            const uint32_t f = (fl >> (3 * t)) & 7u;
            const int mn = (f & 1) ? 0 : ((f & 2) ? 1 : 2);
            const int mx = (f & 4) ? 2 : ((f & 2) ? 1 : 0);
            if (mx == mn) bad = 1;
            const int d = (tid >> (2 * (t & 3))) & 3;
            // exact: (d-mn)/(mx-mn) in {0,.5,1}; scaled so that lo+hi = 16 * mean over the n_maps (1 or 8) maps
            const int v2 = (mx > mn && d < 3) ? ((16 / n_maps) * (d - mn)) / (mx - mn) : 0;
            if (t < 4) lo += v2; else hi += v2;
        }
        s_lo[tid] = (uint8_t)lo;
        s_hi[tid] = (uint8_t)hi;
        if (tid == 0) {
            s_const = bad;
            const float mxv = ordered_to_f32(pmax[b]);
            const bool ok = mxv > 0.0f && mxv < INFINITY;
            s_div = ok ? 0 : 1;
            s_thr = ok ? gate_threshold(mxv) : 0.0f;
            if (bad && status && blockIdx.x == 0 && blockIdx.y == 0) atomicOr(status + b, CDNET_S_DDM_CONSTANT);
        }
    }
    __syncthreads();
    const int x4 = (blockIdx.x * 64 + threadIdx.x) * 4;
    if (x4 >= W) return;
    const size_t plane = (size_t)H * W;  // 2 * plane overflows int for whole-slide tiles
    const float* P = point + (size_t)b * plane;
    const float mx = ordered_to_f32(pmax[b]);
    const float thr = s_thr;
    const bool usediv = s_div != 0;
    const bool isconst = s_const != 0;
    auto gate = [&](float v) -> bool { return usediv ? (__fdiv_rn(v, mx) > 0.2f) : (v >= thr); };
    // kBoostRows rows per thread: the per-block LUT / threshold set-up is amortised over 16 rows
#pragma unroll 1
    for (int ry = 0; ry < kBoostRows; ++ry) {
    const int y = (blockIdx.y * kBoostRows + ry) * 4 + threadIdx.y;
    if (y >= H) break;
    const int p = y * W + x4;
    const float4 pc = *(const float4*)(P + p);
    bool g[4] = {gate(pc.x), gate(pc.y), gate(pc.z), gate(pc.w)};
    const bool gl = x4 > 0 ? gate(P[p - 1]) : false;
    const bool gr = x4 + 4 < W ? gate(P[p + 4]) : false;
    bool d[4];
    d[0] = g[0] | gl | g[1];
    d[1] = g[1] | g[0] | g[2];
    d[2] = g[2] | g[1] | g[3];
    d[3] = g[3] | g[2] | gr;
    if (y > 0) {
        const float4 u = *(const float4*)(P + p - W);
        d[0] |= gate(u.x); d[1] |= gate(u.y); d[2] |= gate(u.z); d[3] |= gate(u.w);
    }
    if (y + 1 < H) {
        const float4 u = *(const float4*)(P + p + W);
        d[0] |= gate(u.x); d[1] |= gate(u.y); d[2] |= gate(u.z); d[3] |= gate(u.w);
    }
    const uint2 cw = *(const uint2*)(codes + (size_t)b * plane + p);
    const uint32_t c[4] = {cw.x & 0xffffu, cw.x >> 16, cw.y & 0xffffu, cw.y >> 16};
    float* PR = prob + (size_t)b * 3 * plane;
    const float4 q0 = *(const float4*)(PR + p);
    const float4 q1 = *(const float4*)(PR + plane + p);
    float4 q2 = *(const float4*)(PR + 2 * plane + p);
    const float a0[4] = {q0.x, q0.y, q0.z, q0.w}, a1[4] = {q1.x, q1.y, q1.z, q1.w};
    float a2[4] = {q2.x, q2.y, q2.z, q2.w};
    uint32_t res = 0;
    bool changed = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int s16 = s_lo[c[i] & 0xff] + s_hi[c[i] >> 8];  // 16 * mean of the 8 maps
        if (isconst) {
            a2[i] = __int_as_float(0x7fc00000);  // NaN DDM (the reference asserts here)
            changed = true;
        } else if (s16 != 0 && !d[i]) {
            const double eb = (double)s16 * 0.125;  // 2 * ddm
            a2[i] = __double2float_rn(__dmul_rn(__dadd_rn((double)a2[i], __dmul_rn(0.5, eb)), __dadd_rn(1.0, eb)));
            changed = true;
        }
        int am = 0;
        float best = a0[i];
        if (a1[i] > best || (a1[i] != a1[i] && best == best)) { am = 1; best = a1[i]; }
        if (a2[i] > best || (a2[i] != a2[i] && best == best)) { am = 2; }
        res |= (am == 1 ? 1u : 0u) << (8 * i);
    }
    *(uint32_t*)(inside + (size_t)b * plane + p) = res;
    if (write_prob && changed) *(float4*)(PR + 2 * plane + p) = make_float4(a2[0], a2[1], a2[2], a2[3]);
    }
}

// one block per tile: the two 256-entry LUTs (16 * mean of the normalised maps over the low / high byte of the code
// word), the gate threshold and the constant-map flag -- what every block of k_boost_inside4<false> recomputes for itself
__global__ void __launch_bounds__(256) k_boost_prep(const uint32_t* __restrict__ flags, const unsigned int* __restrict__ pmax,
                                                    int32_t* __restrict__ status, BoostPrep* __restrict__ prep, int n_maps) {
    pdl_trigger();
    const int b = blockIdx.x, tid = threadIdx.x;
    const uint32_t fl = flags[b];
    int bad = 0, lo = 0, hi = 0;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        if (t >= n_maps) continue;
        const uint32_t f = (fl >> (3 * t)) & 7u;
        const int mn = (f & 1) ? 0 : ((f & 2) ? 1 : 2);
        const int mx = (f & 4) ? 2 : ((f & 2) ? 1 : 0);
        if (mx == mn) bad = 1;
        const int d = (tid >> (2 * (t & 3))) & 3;
        const int v2 = (mx > mn && d < 3) ? ((16 / n_maps) * (d - mn)) / (mx - mn) : 0;
        if (t < 4) lo += v2; else hi += v2;
    }
    prep[b].lo[tid] = (uint8_t)lo;
    prep[b].hi[tid] = (uint8_t)hi;
    if (tid == 0) {
        const float mxv = ordered_to_f32(pmax[b]);
        const bool ok = mxv > 0.0f && mxv < INFINITY;
        prep[b].is_const = bad;
        prep[b].use_div = ok ? 0 : 1;
        prep[b].thr = ok ? gate_threshold(mxv) : 0.0f;
        if (bad && status) atomicOr(status + b, CDNET_S_DDM_CONSTANT);
    }
}

// test.py:270-275: inside = argmax over C channels == 1, or prob[0] >= 0.5
__global__ void __launch_bounds__(256) k_plain_inside(const float* __restrict__ prob, int C, uint8_t* __restrict__ inside,
                                                      size_t plane, int multi_class) {
    const int b = blockIdx.y;
    const float* PR = prob + (size_t)b * C * plane;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (size_t)gridDim.x * blockDim.x) {
        uint8_t r;
        if (multi_class) {
            int am = 0;
            float best = PR[i];
            for (int c = 1; c < C; ++c) {
                const float v = PR[(size_t)c * plane + i];
                if (v > best || (v != v && best == best)) { am = c; best = v; }
            }
            r = (am == 1);
        } else {
            r = PR[i] >= 0.5f;
        }
        inside[(size_t)b * plane + i] = r;
    }
}

// one side stream + two events per (host thread, device), created on first use; CDNET_NO_SIDE_STREAM=1 turns the
// overlap off (everything on the caller's stream)
struct SideStream {
    cudaStream_t stream;
    cudaEvent_t fork, join;
};
static SideStream* side_stream() {
#ifdef CDNET_SIMT  // the host emulator of the test tier has no streams
    return nullptr;
#else
    static int off = -1;
    if (off < 0) off = getenv("CDNET_NO_SIDE_STREAM") ? 1 : 0;
    if (off) return nullptr;
    static thread_local SideStream cache[16];
    static thread_local bool ready[16] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    if (!ready[dev]) {
        if (cudaStreamCreateWithFlags(&cache[dev].stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&cache[dev].fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&cache[dev].join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        ready[dev] = true;
    }
    return &cache[dev];
#endif
}

static size_t tail_workspace(int B, int H, int W) {
    size_t a = fill_remove_label_workspace(B, H, W);
    const size_t c = ws_process_workspace(B, H, W), r = rle_tail_workspace(B, H, W);
    if (r > a) a = r;
    return a > c ? a : c;
}

// inside -> labels (postproc 0: fill/remove/label8; 1: process(); 2: process() as it runs for model_mode 'unet',
// i.e. without the watershed, postproc_other.py:35,50-54) -> dilation
static int tail_launch(const uint8_t* inside, int32_t* labels, void* out, int out_elem_bytes, int32_t* status, int B,
                       int H, int W, int min_area, int ws_min_size, int radius, int postproc, void* ws, size_t ws_bytes,
                       cudaStream_t st) {
    CDNET_RANGE("fill holes / remove small / label / dilate");
    int rc;
    if (postproc == 0 && rle_tail_supported(radius))  // run-based chain, labels and dilation in one pass (rle.cu)
        return rle_tail_launch(inside, out, out_elem_bytes, B, H, W, min_area, radius, ws, ws_bytes, st);
    if (postproc == 1 || postproc == 2)
        rc = ws_process_launch(inside, labels, status, B, H, W, ws_min_size, postproc == 1 ? 1 : 0, ws, ws_bytes, st);
    else rc = fill_remove_label_launch(inside, labels, nullptr, B, H, W, min_area, ws, ws_bytes, st);
    if (rc) return rc;
    return label_dilate_launch(labels, out, out_elem_bytes, B, H, W, radius, st);
}

}  // namespace cdnet

using namespace cdnet;

static bool bad_dims(int B, int H, int W) { return B <= 0 || H <= 0 || W <= 0 || (double)H * W >= 2147483648.0; }

extern "C" size_t cdnet_dam_postproc_workspace_bytes(int B, int H, int W) {
    if (bad_dims(B, H, W)) return 0;
    const size_t n = (size_t)B * H * W;
    return pad256(n * 2) + 2 * pad256((size_t)B * 4) + pad256((size_t)B * sizeof(BoostPrep)) + pad256(n) + pad256(n * 4) +
           tail_workspace(B, H, W);
}

extern "C" int cdnet_dam_postproc(const uint8_t* dcm, int n_maps, float* prob, const float* point, void* out,
                                  int out_elem_bytes, int32_t* status, int B, int H, int W, int direction_classes,
                                  int min_area, int radius, int postproc, int write_prob, void* ws, size_t ws_bytes,
                                  void* stream) {
    if (!dcm || !prob || !point || !out || bad_dims(B, H, W) || (out_elem_bytes != 4 && out_elem_bytes != 8) ||
        postproc < 0 || postproc > 2 || radius < 0 || radius > 4 || (n_maps != 1 && n_maps != 8))
        return CDNET_E_BADARG;
    if (direction_classes != 5 && direction_classes != 9 && direction_classes != 17) return CDNET_E_BADARG;
    if (ws_bytes < cdnet_dam_postproc_workspace_bytes(B, H, W)) return CDNET_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)B * H * W, plane = (size_t)H * W;
    Arena ar(ws, ws_bytes);
    uint16_t* codes = ar.take<uint16_t>(n);
    uint32_t* flags = ar.take<uint32_t>(B);
    unsigned int* pmax = ar.take<unsigned int>(B);
    BoostPrep* prep = ar.take<BoostPrep>(B);
    uint8_t* inside = ar.take<uint8_t>(n);
    int32_t* labels = ar.take<int32_t>(n);
    if (!ar.ok) return CDNET_E_WORKSPACE;
    void* tail_ws = (char*)ws + ar.off;
    const size_t tail_bytes = ws_bytes - ar.off;
    if (status) CDNET_CUDA_OK(cudaMemsetAsync(status, 0, sizeof(int32_t) * (size_t)B, st));
    CDNET_RANGE("cdnet_dam_postproc");
    int rc;
    // the point-map maximum (DRAM-bound) does not depend on the direction-difference codes (ALU-bound): it runs on a
    // side stream next to them and joins before the boost (fork / join with events, capturable in a CUDA graph)
    SideStream* side = g_prof_on ? nullptr : side_stream();  // the per-kernel profiler times kernels one at a time
    if (side) {
        CDNET_CUDA_OK(cudaEventRecord(side->fork, st));
        CDNET_CUDA_OK(cudaStreamWaitEvent(side->stream, side->fork, 0));
        CDNET_CUDA_OK(cudaMemsetAsync(pmax, 0, sizeof(unsigned int) * (size_t)B, side->stream));
        point_max_launch(point, pmax, B, plane, side->stream);
        CDNET_CUDA_OK(cudaEventRecord(side->join, side->stream));
    }
    {
        CDNET_RANGE("ddm codes");
        rc = ddm_codes_launch(dcm, codes, flags, B, n_maps, H, W, direction_classes, st);
    }
    if (side) CDNET_CUDA_OK(cudaStreamWaitEvent(st, side->join, 0));
    if (rc) return rc;
    CDNET_RANGE("point gate + boost + argmax, then labels");
    if (!side) {
        CDNET_CUDA_OK(cudaMemsetAsync(pmax, 0, sizeof(unsigned int) * (size_t)B, st));
        point_max_launch(point, pmax, B, plane, st);
    }
    if (W % 4 == 0 && ((uintptr_t)prob & 15) == 0 && ((uintptr_t)point & 15) == 0) {
        CDNET_LAUNCH(k_boost_prep, B, 256, 0, st, flags, pmax, status, prep, n_maps);
        CDNET_LAUNCH_PDL(k_boost_inside4<true>, dim3(ceil_div(W, 256), ceil_div(H, 4 * kBoostRows), B), dim3(64, 4), 0, st, codes,
                     flags, point, pmax, prob, inside, status, H, W, write_prob, n_maps, prep);
    } else {
        CDNET_LAUNCH(k_boost_inside, dim3(ceil_div(W, 64), ceil_div(H, 4), B), dim3(64, 4), 0, st, codes, flags, point,
                     pmax, prob, inside, status, H, W, write_prob, n_maps);
    }
    rc = last_error();
    if (rc) return rc;
    // test_dam.py:559 calls process() with its default min_size = 10
    return tail_launch(inside, labels, out, out_elem_bytes, status, B, H, W, min_area, 10, radius, postproc, tail_ws,
                       tail_bytes, st);
}

extern "C" size_t cdnet_plain_postproc_workspace_bytes(int B, int H, int W) {
    if (bad_dims(B, H, W)) return 0;
    const size_t n = (size_t)B * H * W;
    return pad256(n) + pad256(n * 4) + tail_workspace(B, H, W);
}

extern "C" int cdnet_plain_postproc(const float* prob, int C, void* out, int out_elem_bytes, int32_t* status, int B, int H,
                                    int W, int multi_class, int min_area, int radius, int postproc, void* ws,
                                    size_t ws_bytes, void* stream) {
    if (!prob || !out || C <= 0 || bad_dims(B, H, W) || (out_elem_bytes != 4 && out_elem_bytes != 8) ||
        postproc < 0 || postproc > 2 || radius < 0 || radius > 4)
        return CDNET_E_BADARG;
    if (ws_bytes < cdnet_plain_postproc_workspace_bytes(B, H, W)) return CDNET_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)B * H * W, plane = (size_t)H * W;
    Arena ar(ws, ws_bytes);
    uint8_t* inside = ar.take<uint8_t>(n);
    int32_t* labels = ar.take<int32_t>(n);
    if (!ar.ok) return CDNET_E_WORKSPACE;
    void* tail_ws = (char*)ws + ar.off;
    const size_t tail_bytes = ws_bytes - ar.off;
    if (status) CDNET_CUDA_OK(cudaMemsetAsync(status, 0, sizeof(int32_t) * (size_t)B, st));
    int gx = (int)((plane + 256 * 4 - 1) / (256 * 4));
    CDNET_LAUNCH(k_plain_inside, dim3(gx, B), 256, 0, st, prob, C, inside, plane, multi_class);
    // test.py:289-290 passes min_size = min_area to process()
    return tail_launch(inside, labels, out, out_elem_bytes, status, B, H, W, min_area, min_area, radius, postproc, tail_ws,
                       tail_bytes, st);
}

extern "C" size_t cdnet_ws_postproc_workspace_bytes(int B, int H, int W) {
    if (bad_dims(B, H, W)) return 0;
    return ws_process_workspace(B, H, W);
}

extern "C" int cdnet_ws_postproc(const uint8_t* pred01, int32_t* labels, int32_t* status, int B, int H, int W, int min_size,
                                 int ws_flag, void* ws, size_t ws_bytes, void* stream) {
    if (!pred01 || !labels || bad_dims(B, H, W)) return CDNET_E_BADARG;
    if (ws_bytes < ws_process_workspace(B, H, W)) return CDNET_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    if (status) CDNET_CUDA_OK(cudaMemsetAsync(status, 0, sizeof(int32_t) * (size_t)B, st));
    return ws_process_launch(pred01, labels, status, B, H, W, min_size, ws_flag, ws, ws_bytes, st);
}

// process() on the extended tile of a row-sharded slide (cdnet_b200/sharded.py, postproc = 1): rows [own_lo, own_hi)
// are the rank's own, the rest is overlap.  Also reports the largest marker id per row (before small markers are
// dropped), from which the ranks number the markers of the whole slide, and CDNET_S_SHARD_OVERFLOW.
extern "C" int cdnet_shard_ws_process(const uint8_t* pred01, int32_t* labels, int32_t* marker_rowmax, int32_t* status,
                                      int H, int W, int own_lo, int own_hi, int min_size, void* ws, size_t ws_bytes,
                                      void* stream) {
    if (!pred01 || !labels || !marker_rowmax || !status || bad_dims(1, H, W) || own_lo < 0 || own_hi > H ||
        own_lo >= own_hi)
        return CDNET_E_BADARG;
    if (ws_bytes < ws_process_workspace(1, H, W)) return CDNET_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    CDNET_CUDA_OK(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
    return ws_process_launch(pred01, labels, status, 1, H, W, min_size, 1, ws, ws_bytes, st, marker_rowmax, own_lo, own_hi);
}

// tile-local marker ids of a rank's own rows -> slide-global ids, one pass.  sc = {markers that start above the own
// rows, markers that start on them (= owned), markers owned by the lower ranks} on the device; an owned id l maps to
// sc[2] + l - sc[0]; any other id was adopted from a row neighbour through lut (0 = nobody claimed it -> err).
__global__ void __launch_bounds__(256) k_shard_ws_relabel(const int32_t* __restrict__ labels, const int32_t* __restrict__ sc,
                                                          const int32_t* __restrict__ lut, int32_t* __restrict__ out,
                                                          int32_t* __restrict__ err, size_t n) {
    const int above = sc[0], owned = sc[1], off = sc[2];
    bool bad = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int l = labels[i];
        int g = 0;
        if (l > above && l <= above + owned) g = off + l - above;
        else if (l > 0) {
            g = lut[l];
            bad = bad || g == 0;
        }
        out[i] = g;
    }
    if (bad) *err = 1;
}

extern "C" int cdnet_shard_ws_relabel(const int32_t* labels, const int32_t* scalars, const int32_t* lut, int32_t* out,
                                      int32_t* err, int rows, int W, void* stream) {
    if (!labels || !scalars || !lut || !out || !err || rows <= 0 || W <= 0) return CDNET_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)rows * W;
    const size_t blocks = (n + 1023) / 1024;
    CDNET_LAUNCH(k_shard_ws_relabel, (unsigned)(blocks > 148u * 16 ? 148u * 16 : blocks), 256, 0, st, labels, scalars, lut, out,
                 err, n);
    return last_error();
}

// ---- whole-slide shard pieces of the DAM chain (cdnet_b200/sharded.py) -----------------------------
// max of the shard's own point-map rows as an order-preserving uint32 (the host all-reduces MAX)
extern "C" int cdnet_shard_point_max(const float* point, uint32_t* pmax, size_t n, void* stream) {
    if (!point || !pmax || n == 0) return CDNET_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    CDNET_CUDA_OK(cudaMemsetAsync(pmax, 0, sizeof(uint32_t), st));
    point_max_launch(point, pmax, 1, n, st);
    return last_error();
}

// boost + argmax on one extended tile; flags / pmax hold the slide-global values; n_maps 1 or 8
extern "C" int cdnet_shard_boost(const uint16_t* codes, const uint32_t* flags, const float* point, const uint32_t* pmax,
                                 float* prob, uint8_t* inside, int32_t* status, int He, int W, int n_maps, int write_prob,
                                 void* stream) {
    if (!codes || !flags || !point || !pmax || !prob || !inside || He <= 0 || W <= 0 || (n_maps != 1 && n_maps != 8))
        return CDNET_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int H = He, B = 1;
    if (W % 4 == 0 && ((uintptr_t)prob & 15) == 0 && ((uintptr_t)point & 15) == 0)
        CDNET_LAUNCH(k_boost_inside4<false>, dim3(ceil_div(W, 256), ceil_div(H, 4 * kBoostRows), B), dim3(64, 4), 0, st, codes,
                     flags, point, pmax, prob, inside, status, H, W, write_prob, n_maps, (const BoostPrep*)nullptr);
    else
        CDNET_LAUNCH(k_boost_inside, dim3(ceil_div(W, 64), ceil_div(H, 4), B), dim3(64, 4), 0, st, codes, flags, point,
                     pmax, prob, inside, status, H, W, write_prob, n_maps);
    return last_error();
}

// ---- DcmVoting2, utils.py:1150-1159 ("next" row of SURVEY.md section 8f) ---------------------------
// The 8 TTA direction maps are brought into the un-flipped frame by fixed class permutations and voted per
// pixel over the 9 classes; np.argmax keeps the first maximum.
namespace cdnet {
__constant__ uint8_t c_vote_perm[8][9] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8}, {0, 5, 4, 3, 2, 1, 8, 7, 6}, {0, 1, 8, 7, 6, 5, 4, 3, 2}, {0, 5, 6, 7, 8, 1, 2, 3, 4},
    {0, 3, 4, 5, 6, 7, 8, 1, 2}, {0, 7, 6, 5, 4, 3, 2, 1, 8}, {0, 3, 2, 1, 8, 7, 6, 5, 4}, {0, 7, 8, 1, 2, 3, 4, 5, 6}};

__global__ void __launch_bounds__(256) k_dcm_voting2(const uint8_t* __restrict__ dcm, uint8_t* __restrict__ out, size_t plane) {
    const int b = blockIdx.y;
    const uint8_t* D = dcm + (size_t)b * 8 * plane;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (size_t)gridDim.x * blockDim.x) {
        unsigned long long votes = 0;  // nine 4-bit counters (max 8 votes)
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int c = D[(size_t)t * plane + i];
            if (c <= 8) votes += 1ull << (4 * c_vote_perm[t][c]);
        }
        int best = 0, bestv = (int)(votes & 15);
#pragma unroll
        for (int c = 1; c < 9; ++c) {
            const int v = (int)((votes >> (4 * c)) & 15);
            if (v > bestv) { bestv = v; best = c; }
        }
        out[(size_t)b * plane + i] = (uint8_t)best;
    }
}
}  // namespace cdnet

extern "C" int cdnet_dcm_voting2(const uint8_t* dcm, uint8_t* out, int B, int H, int W, void* stream) {
    if (!dcm || !out || bad_dims(B, H, W)) return CDNET_E_BADARG;
    const size_t plane = (size_t)H * W;
    int gx = (int)((plane + 256 * 4 - 1) / (256 * 4));
    if (gx > 65535) gx = 65535;
    CDNET_LAUNCH(cdnet::k_dcm_voting2, dim3(gx, B), 256, 0, (cudaStream_t)stream, dcm, out, plane);
    return cdnet::last_error();
}
