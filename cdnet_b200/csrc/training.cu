// training.cu -- the direction quantiser as stand-alone operators and the training-side consumers of the
// target transform.
//
// Replaces, one streaming kernel each (all HBM-bound, no reuse, 128-bit or 64-bit accesses where the layout
// allows):
//   DTOffsetHelper.label_to_vector      data_prepare/SegFix_offset_helper.py:246-261 (+ table :50-89)
//   DTOffsetHelper.align_angle(_c4)     :286-341
//   DTOffsetHelper.angle_to_vector      :423-450
//   DTOffsetHelper.vector_to_label      :486-506 (through angle_to_direction_label :452-484)
//   direction one-hot + foreground mask train_util_dam.py:123-142
//   LabelEncoding without direction     my_transforms.py:661-760 (ternary / binary label image only)
#include <math.h>

#include "internal.h"

namespace cdnet {

// ---- class -> (dh, dw) ------------------------------------------------------------------------------
// The tables of SegFix_offset_helper.py:50-89 are rings of radius r = n/8 walked from (0,-r) up the left
// column, right along the top row, down the right column and back along the bottom row; 5 / 9 / 17 classes
// prepend the zero vector for background; 4 (and 5) classes use the four diagonals (c4_align_axis unset).
struct VecTable {
    int dh[36], dw[36];  // int, not char: a by-value kernel parameter indexed with a register stays in the
    int n;               // constant bank only for 4-byte (or wider) elements
};

static bool build_vec_table(int num_classes, VecTable* t) {
    int k = 0;
    const bool bg = num_classes == 5 || num_classes == 9 || num_classes == 17;
    const int dirs = bg ? num_classes - 1 : num_classes;
    if (bg) { t->dh[k] = 0; t->dw[k] = 0; ++k; }
    if (dirs == 4) {
        const int d[4][2] = {{-1, -1}, {-1, 1}, {1, 1}, {1, -1}};
        for (int i = 0; i < 4; ++i, ++k) { t->dh[k] = d[i][0]; t->dw[k] = d[i][1]; }
    } else if (dirs == 8 || dirs == 16 || dirs == 32) {
        const int r = dirs / 8;
        int h = 0, w = -r;
        for (int i = 0; i < dirs; ++i, ++k) {
            t->dh[k] = h;
            t->dw[k] = w;
            if (w == -r && h > -r && i < dirs / 2) --h;      // up the left column
            else if (h == -r && w < r) ++w;                    // along the top row
            else if (w == r && h < r) ++h;                     // down the right column
            else if (h == r && w > -r) --w;                    // back along the bottom row
            else --h;                                          // up the left column to (1, -r)
        }
    } else {
        return false;
    }
    t->n = k;
    return true;
}

template <typename T>
__global__ void __launch_bounds__(256) k_label_to_vector(const T* __restrict__ labels, long long* __restrict__ out,
                                                         size_t plane, VecTable tab) {
    // the table is indexed per pixel: staged in shared memory
    __shared__ int s_h[36], s_w[36];
    if (threadIdx.x < 36) {
        s_h[threadIdx.x] = threadIdx.x < tab.n ? tab.dh[threadIdx.x] : 0;
        s_w[threadIdx.x] = threadIdx.x < tab.n ? tab.dw[threadIdx.x] : 0;
    }
    __syncthreads();
    const size_t img = blockIdx.y;
    const T* L = labels + img * plane;
    long long* oh = out + img * 2 * plane;
    long long* ow = oh + plane;
    const long long n_tab = tab.n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (size_t)gridDim.x * blockDim.x) {
        const long long v = (long long)L[i];
        const int k = (v >= 0 && v < n_tab) ? (int)v : 35;  // slot 35 is never a class: (0, 0)
        oh[i] = s_h[k];
        ow[i] = s_w[k];
    }
}

// ---- 128-bit accesses: four consecutive elements from a 16-byte aligned address -------------------------------------
template <typename T>
struct Quad {
    T v[4];
};
__device__ __forceinline__ Quad<float> ld_quad(const float* p) {
    const float4 t = __ldg((const float4*)p);
    Quad<float> q;
    q.v[0] = t.x; q.v[1] = t.y; q.v[2] = t.z; q.v[3] = t.w;
    return q;
}
__device__ __forceinline__ Quad<double> ld_quad(const double* p) {
    const double2 a = __ldg((const double2*)p), b = __ldg((const double2*)p + 1);
    Quad<double> q;
    q.v[0] = a.x; q.v[1] = a.y; q.v[2] = b.x; q.v[3] = b.y;
    return q;
}
__device__ __forceinline__ void st_quad(float* p, const Quad<float>& q) {
    *(float4*)p = make_float4(q.v[0], q.v[1], q.v[2], q.v[3]);
}
__device__ __forceinline__ void st_quad(double* p, const Quad<double>& q) {
    *(double2*)p = make_double2(q.v[0], q.v[1]);
    *((double2*)p + 1) = make_double2(q.v[2], q.v[3]);
}
__device__ __forceinline__ void st_quad(long long* p, const Quad<long long>& q) {
    *(longlong2*)p = make_longlong2(q.v[0], q.v[1]);
    *((longlong2*)p + 1) = make_longlong2(q.v[2], q.v[3]);
}

// ---- align_angle --------------------------------------------------------------------------------------
// Index of the (upper-inclusive) bin of `a`.  The bin edges T_j = -180 + step * (j + 1/2), j = 0..n-1, are exact in
// f32 and f64 for the supported n (8, 16, 32), so comparing the f64-promoted angle is what numpy / torch compute;
// the reference's mask loop (:323-339) assigns bin #{j : T_j < a}, the count n wrapping to bin 0 -- found here
// by bisection (n is a power of two).  NaN matches no bin: the reference leaves its zero-initialised outputs
// (index 0, angle 0.0).
__device__ __forceinline__ int align_bin(double a, int n, bool* matched) {
    const double step = 360.0 / (double)n, t0 = -180.0 + 0.5 * step;
    *matched = (a == a);
    int c = 0;
    for (int s = n >> 1; s >= 1; s >>= 1)
        if (t0 + step * (double)(c + s - 1) < a) c += s;
    if (t0 + step * (double)c < a) c += 1;
    return c == n ? 0 : c;
}

// align_angle_c4 (:286-309): trunc((a + 180) / 90) in the angle's own precision, clamped to 0..3.
// torch.trunc(..).long(): NaN / out-of-range conversions are implementation-defined on the host (x86: INT64_MIN);
// the clamp makes every such value 0, +inf included (it is INT64_MIN there too)
template <typename TI>
__device__ __forceinline__ int align_bin_c4(TI a) {
    const TI q = (a + (TI)180) / (TI)90;
    long long k = 0;
    if (q == q && q > (TI)-9.0e18 && q < (TI)9.0e18) k = (long long)q;  // the C cast truncates toward zero
    return (int)(k < 0 ? 0 : (k > 3 ? 3 : k));
}

// Every kernel below is a stream over n elements: quads first (`vec` = every base pointer is 16-byte aligned),
// then the ragged tail element by element.
#define CDNET_STREAM_QUADS(n, vec)                                                        \
    const size_t nq = (vec) ? (n) / 4 : 0;                                                \
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x

// TI: input angle type; TO: type of the snapped angle (numpy path f64, torch path f32)
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) k_align_angle(const TI* __restrict__ angle, TO* __restrict__ snapped,
                                                     long long* __restrict__ index, size_t n, int classes, int vec) {
    CDNET_STREAM_QUADS(n, vec);
    const double step = 360.0 / (double)classes;
    for (size_t q = tid; q < nq; q += nth) {
        const Quad<TI> a = ld_quad(angle + 4 * q);
        Quad<TO> s;
        Quad<long long> k;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            bool hit;
            const int bin = align_bin((double)a.v[j], classes, &hit);
            s.v[j] = hit ? (TO)(-180.0 + step * (double)bin) : (TO)0.0;
            k.v[j] = bin;
        }
        if (snapped) st_quad(snapped + 4 * q, s);
        if (index) st_quad(index + 4 * q, k);
    }
    for (size_t i = nq * 4 + tid; i < n; i += nth) {
        bool hit;
        const int bin = align_bin((double)angle[i], classes, &hit);
        if (snapped) snapped[i] = hit ? (TO)(-180.0 + step * (double)bin) : (TO)0.0;
        if (index) index[i] = bin;
    }
}

// the snapped angle of align_angle_c4 is float32 on both of the reference's paths
template <typename TI>
__global__ void __launch_bounds__(256) k_align_angle_c4(const TI* __restrict__ angle, float* __restrict__ snapped,
                                                        long long* __restrict__ index, size_t n, int vec) {
    CDNET_STREAM_QUADS(n, vec);
    for (size_t q = tid; q < nq; q += nth) {
        const Quad<TI> a = ld_quad(angle + 4 * q);
        Quad<float> s;
        Quad<long long> k;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int bin = align_bin_c4<TI>(a.v[j]);
            s.v[j] = (float)(bin * 90 - 135);
            k.v[j] = bin;
        }
        if (snapped) st_quad(snapped + 4 * q, s);
        if (index) st_quad(index + 4 * q, k);
    }
    for (size_t i = nq * 4 + tid; i < n; i += nth) {
        const int bin = align_bin_c4<TI>(angle[i]);
        if (snapped) snapped[i] = (float)(bin * 90 - 135);
        if (index) index[i] = bin;
    }
}

// ---- angle_to_vector: snap, then (sin, cos) of the bin centre from a host-computed table ------------------
struct SinCosTable {
    double s[33], c[33];  // [classes] = the row for angles no bin matches (NaN keeps the snapped angle 0.0)
};
template <typename TI>
__device__ __forceinline__ int sincos_row(TI a, int classes, int c4) {
    if (c4) return align_bin_c4<TI>(a);
    bool hit;
    const int bin = align_bin((double)a, classes, &hit);
    return hit ? bin : classes;
}
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) k_angle_to_vector(const TI* __restrict__ angle, TO* __restrict__ vec, size_t n,
                                                         int classes, int c4, SinCosTable tab, int vecq) {
    CDNET_STREAM_QUADS(n, vecq);
    for (size_t q = tid; q < nq; q += nth) {
        const Quad<TI> a = ld_quad(angle + 4 * q);
        Quad<TO> lo, hi;  // (s0, c0, s1, c1), (s2, c2, s3, c3)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = sincos_row<TI>(a.v[j], classes, c4);
            Quad<TO>& d = j < 2 ? lo : hi;
            d.v[2 * (j & 1)] = (TO)tab.s[r];
            d.v[2 * (j & 1) + 1] = (TO)tab.c[r];
        }
        st_quad(vec + 8 * q, lo);
        st_quad(vec + 8 * q + 4, hi);
    }
    for (size_t i = nq * 4 + tid; i < n; i += nth) {
        const int r = sincos_row<TI>(angle[i], classes, c4);
        vec[2 * i] = (TO)tab.s[r];
        vec[2 * i + 1] = (TO)tab.c[r];
    }
}

// ---- vector_to_label: atan2 -> degrees -> bin index (no ignore mask: seg_label_map is None at :503-505) ----
// np.arctan2 in the vector's precision, np.rad2deg = x * (180 / pi) in the same precision
template <typename TI>
__device__ __forceinline__ long long vector_bin(TI v0, TI v1, int classes, int c4) {
    double deg;
    if (sizeof(TI) == 4) {
        const float a = (float)atan2((double)v0, (double)v1);
        deg = (double)__fmul_rn(a, 57.295779513082320876798154814105f);
        if (c4) return align_bin_c4<float>((float)deg);
    } else {
        deg = __dmul_rn(atan2((double)v0, (double)v1), 57.295779513082320876798154814105);
        if (c4) return align_bin_c4<double>(deg);
    }
    bool hit;
    return align_bin(deg, classes, &hit);
}
template <typename TI>
__global__ void __launch_bounds__(256) k_vector_to_label(const TI* __restrict__ vec, long long* __restrict__ label,
                                                         size_t n, int classes, int c4, int vecq) {
    CDNET_STREAM_QUADS(n, vecq);
    for (size_t q = tid; q < nq; q += nth) {
        const Quad<TI> lo = ld_quad(vec + 8 * q), hi = ld_quad(vec + 8 * q + 4);
        Quad<long long> k;
        k.v[0] = vector_bin<TI>(lo.v[0], lo.v[1], classes, c4);
        k.v[1] = vector_bin<TI>(lo.v[2], lo.v[3], classes, c4);
        k.v[2] = vector_bin<TI>(hi.v[0], hi.v[1], classes, c4);
        k.v[3] = vector_bin<TI>(hi.v[2], hi.v[3], classes, c4);
        st_quad(label + 4 * q, k);
    }
    for (size_t i = nq * 4 + tid; i < n; i += nth) label[i] = vector_bin<TI>(vec[2 * i], vec[2 * i + 1], classes, c4);
}

// ---- direction one-hot + foreground mask (train_util_dam.py:123-142) ------------------------------------------
// per tile: [0] = min class id, [1] = max class id (a tile with min == max has ONE distinct value)
__global__ void k_dir_minmax_init(long long* __restrict__ mm, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) {
        mm[2 * b] = 0x7fffffffffffffffll;
        mm[2 * b + 1] = -0x7fffffffffffffffll - 1;
    }
}

__global__ void __launch_bounds__(256) k_dir_minmax(const long long* __restrict__ dir, long long* __restrict__ mm,
                                                    size_t plane) {
    const int b = blockIdx.y;
    const long long* D = dir + (size_t)b * plane;
    long long lo = 0x7fffffffffffffffll, hi = -0x7fffffffffffffffll - 1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (size_t)gridDim.x * blockDim.x) {
        const long long v = D[i];
        lo = v < lo ? v : lo;
        hi = v > hi ? v : hi;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const long long l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = l2 < lo ? l2 : lo;
        hi = h2 > hi ? h2 : hi;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(mm + 2 * b, lo);
        atomicMax(mm + 2 * b + 1, hi);
    }
}

// out[b][k][p] = 1 where direction[b][p] == k and the pixel is foreground in the ternary target OF TILE 0
// (the reference indexes `target[0, :, :]` for every j, :139); a tile with a single distinct direction value
// gets channel 0 set everywhere and no mask (:141).  Ids outside [0, C) raise IndexError in the reference:
// CDNET_S_CLASS_RANGE.  Four pixels per thread when the plane allows 128-bit stores.
template <typename TT, int VEC>
__global__ void __launch_bounds__(256) k_dir_one_hot(const long long* __restrict__ dir, const TT* __restrict__ target0,
                                                     const long long* __restrict__ mm, float* __restrict__ out,
                                                     int32_t* __restrict__ status, int C, size_t plane) {
    const int b = blockIdx.y;
    const long long lo = mm[2 * b], hi = mm[2 * b + 1];
    const bool multi = lo != hi;
    if (multi && (lo < 0 || hi >= C)) {
        if (status && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(status + b, CDNET_S_CLASS_RANGE);
        return;  // the reference raises before producing anything for this batch
    }
    const long long* D = dir + (size_t)b * plane;
    float* O = out + (size_t)b * C * plane;
    const size_t nvec = plane / VEC;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
        long long d[VEC];
        bool fg[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            d[j] = D[i * VEC + j];
            const TT t = target0[i * VEC + j];
            fg[j] = (t == (TT)1) || (t == (TT)2);
        }
        for (int k = 0; k < C; ++k) {
            float v[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) v[j] = multi ? ((d[j] == k && fg[j]) ? 1.0f : 0.0f) : (k == 0 ? 1.0f : 0.0f);
            if (VEC == 4) *(float4*)(O + (size_t)k * plane + i * 4) = make_float4(v[0], v[1], v[2], v[3]);
            else O[(size_t)k * plane + i] = v[0];
        }
    }
}

// ---- my_transforms.LabelEncoding without direction (my_transforms.py:661-760) -----------------------------------
// mode 0: out_c == 3, instance ids   (:713-727)  1 inside, 2 where cross-max != cross-min of the ids
//         (dilation(L) & ~erosion(L, disk(1)) > 0 on L = measure.label(ids): two pixels of one cross differ in
//         L exactly when they differ in ids, because equal ids inside one cross are 8-connected)
// mode 1: out_c == 3, {0,255} label  (:728-742)  1 where > 127.5, 2 on the same boundary of the binary image
// mode 2: out_c != 3, instance ids   (:690-699)  2 where id > 0
// mode 3: out_c != 3, {0,255} label  (:700-708)  erosion(disk(1)) of 2 * (ch0 > 127.5 or ch1 > 127.5)
// output image = uint8(new_label / 2 * 255) in {0, 127, 255} (:761)
__global__ void __launch_bounds__(256) k_ternary_label(const uint8_t* __restrict__ ch0, const uint8_t* __restrict__ ch1,
                                                       int mode, uint8_t* __restrict__ out, int H, int W) {
    const int x = blockIdx.x * 64 + threadIdx.x;
    const int y = blockIdx.y * 4 + threadIdx.y;
    const int b = blockIdx.z;
    if (x >= W || y >= H) return;
    const size_t tile = (size_t)b * H * W;
    const uint8_t* A = ch0 + tile;
    const uint8_t* Bc = ch1 ? ch1 + tile : nullptr;
    const int p = y * W + x;
    auto val = [&](int q) -> int {
        const int v = A[q];
        if (mode == 0) return v;
        if (mode == 1) return v > 127 ? 1 : 0;
        if (mode == 2) return v > 0 ? 2 : 0;
        return (v > 127 || (Bc && Bc[q] > 127)) ? 2 : 0;
    };
    const int v = val(p);
    int mx = v, mn = v;
    if (mode != 2) {
        if (y > 0) { const int u = val(p - W); mx = max(mx, u); mn = min(mn, u); }
        if (y + 1 < H) { const int u = val(p + W); mx = max(mx, u); mn = min(mn, u); }
        if (x > 0) { const int u = val(p - 1); mx = max(mx, u); mn = min(mn, u); }
        if (x + 1 < W) { const int u = val(p + 1); mx = max(mx, u); mn = min(mn, u); }
    }
    int nl;
    if (mode == 0 || mode == 1) nl = (mx != mn) ? 2 : (v > 0 ? 1 : 0);
    else if (mode == 2) nl = v;
    else nl = mn;  // erosion: minimum over the cross, taps outside the image ignored
    out[tile + p] = nl == 0 ? 0 : (nl == 1 ? 127 : 255);
}

// four pixels per thread for modes 0-2 (W % 4 == 0, 4-byte aligned planes): the rows above / below and the centre
// row arrive as 32-bit words, the two horizontal neighbours outside the word as single bytes
__global__ void __launch_bounds__(256) k_ternary_label4(const uint8_t* __restrict__ ch0, int mode,
                                                        uint8_t* __restrict__ out, int H, int W) {
    const int x = (blockIdx.x * 64 + threadIdx.x) * 4;
    const int y = blockIdx.y * 4 + threadIdx.y;
    const int b = blockIdx.z;
    if (x >= W || y >= H) return;
    const size_t tile = (size_t)b * H * W;
    const uint8_t* A = ch0 + tile + (size_t)y * W + x;
    auto val = [&](unsigned v) -> int { return mode == 0 ? (int)v : (mode == 1 ? (v > 127u ? 1 : 0) : (v > 0u ? 2 : 0)); };
    const unsigned wc = __ldg((const unsigned*)A);
    unsigned res = 0;
    if (mode == 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j) res |= (((wc >> (8 * j)) & 0xffu) ? 255u : 0u) << (8 * j);
    } else {
        // a missing neighbour (outside the image) is replaced by the centre pixel: it changes neither max nor min
        const unsigned wu = y > 0 ? __ldg((const unsigned*)(A - W)) : wc;
        const unsigned wd = y + 1 < H ? __ldg((const unsigned*)(A + W)) : wc;
        const unsigned left = x > 0 ? (unsigned)__ldg(A - 1) : (wc & 0xffu);
        const unsigned right = x + 4 < W ? (unsigned)__ldg(A + 4) : (wc >> 24);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int v = val((wc >> (8 * j)) & 0xffu);
            const int u = val((wu >> (8 * j)) & 0xffu), d = val((wd >> (8 * j)) & 0xffu);
            const int l = val(j == 0 ? left : (wc >> (8 * (j - 1))) & 0xffu);
            const int r = val(j == 3 ? right : (wc >> (8 * (j + 1))) & 0xffu);
            const int mx = max(max(max(v, u), max(d, l)), r), mn = min(min(min(v, u), min(d, l)), r);
            const unsigned nl = (mx != mn) ? 255u : (v > 0 ? 127u : 0u);
            res |= nl << (8 * j);
        }
    }
    *(unsigned*)(out + tile + (size_t)y * W + x) = res;
}

static inline unsigned stream_grid(size_t n, int per_block) {
    size_t g = (n + per_block - 1) / per_block;
    if (g < 1) g = 1;
    return (unsigned)(g > 148u * 32u ? 148u * 32u : g);
}

}  // namespace cdnet

using namespace cdnet;

extern "C" int cdnet_label_to_vector(const void* labels, int elem_bytes, int64_t* out, int N, size_t plane,
                                     int num_classes, void* stream) {
    VecTable tab;
    if (!labels || !out || N <= 0 || N > 65535 || plane == 0 || !build_vec_table(num_classes, &tab)) return CDNET_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid(stream_grid(plane, 256 * 4), N);
    if (elem_bytes == 1) CDNET_LAUNCH(k_label_to_vector<uint8_t>, grid, 256, 0, st, (const uint8_t*)labels, (long long*)out, plane, tab);
    else if (elem_bytes == 4) CDNET_LAUNCH(k_label_to_vector<int32_t>, grid, 256, 0, st, (const int32_t*)labels, (long long*)out, plane, tab);
    else if (elem_bytes == 8) CDNET_LAUNCH(k_label_to_vector<long long>, grid, 256, 0, st, (const long long*)labels, (long long*)out, plane, tab);
    else return CDNET_E_BADARG;
    return last_error();
}

static bool align_classes_ok(int n) { return n == 4 || n == 8 || n == 16 || n == 32; }
static int aligned16(const void* a, const void* b, const void* c) {
    return (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15u) == 0;  // NULL counts as aligned
}

extern "C" int cdnet_align_angle(const void* angle, int in_elem_bytes, void* snapped, int out_elem_bytes, int64_t* index,
                                 size_t n, int num_classes, void* stream) {
    if (!angle || n == 0 || !align_classes_ok(num_classes) || (in_elem_bytes != 4 && in_elem_bytes != 8)) return CDNET_E_BADARG;
    if (snapped && out_elem_bytes != 4 && out_elem_bytes != 8) return CDNET_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = stream_grid(n, 256 * 4);
    long long* idx = (long long*)index;
    const int vec = aligned16(angle, snapped, index);
    if (num_classes == 4) {
        if (snapped && out_elem_bytes != 4) return CDNET_E_BADARG;  // align_angle_c4 returns float32
        if (in_elem_bytes == 4) CDNET_LAUNCH(k_align_angle_c4<float>, grid, 256, 0, st, (const float*)angle, (float*)snapped, idx, n, vec);
        else CDNET_LAUNCH(k_align_angle_c4<double>, grid, 256, 0, st, (const double*)angle, (float*)snapped, idx, n, vec);
        return last_error();
    }
    if (in_elem_bytes == 4) {
        if (!snapped || out_elem_bytes == 8) CDNET_LAUNCH((k_align_angle<float, double>), grid, 256, 0, st, (const float*)angle, (double*)snapped, idx, n, num_classes, vec);
        else CDNET_LAUNCH((k_align_angle<float, float>), grid, 256, 0, st, (const float*)angle, (float*)snapped, idx, n, num_classes, vec);
    } else {
        if (!snapped || out_elem_bytes == 8) CDNET_LAUNCH((k_align_angle<double, double>), grid, 256, 0, st, (const double*)angle, (double*)snapped, idx, n, num_classes, vec);
        else CDNET_LAUNCH((k_align_angle<double, float>), grid, 256, 0, st, (const double*)angle, (float*)snapped, idx, n, num_classes, vec);
    }
    return last_error();
}

extern "C" int cdnet_angle_to_vector(const void* angle, int in_elem_bytes, void* vec, int out_elem_bytes,
                                     const double* table, size_t n, int num_classes, void* stream) {
    if (!angle || !vec || !table || n == 0 || !align_classes_ok(num_classes)) return CDNET_E_BADARG;
    if ((in_elem_bytes != 4 && in_elem_bytes != 8) || (out_elem_bytes != 4 && out_elem_bytes != 8)) return CDNET_E_BADARG;
    SinCosTable tab;
    for (int i = 0; i < 33; ++i) { tab.s[i] = 0.0; tab.c[i] = 0.0; }
    for (int i = 0; i <= num_classes; ++i) { tab.s[i] = table[2 * i]; tab.c[i] = table[2 * i + 1]; }
    const int c4 = num_classes == 4;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = stream_grid(n, 256 * 4);
    const int vq = aligned16(angle, vec, nullptr);
    if (in_elem_bytes == 4 && out_elem_bytes == 8) CDNET_LAUNCH((k_angle_to_vector<float, double>), grid, 256, 0, st, (const float*)angle, (double*)vec, n, num_classes, c4, tab, vq);
    else if (in_elem_bytes == 4) CDNET_LAUNCH((k_angle_to_vector<float, float>), grid, 256, 0, st, (const float*)angle, (float*)vec, n, num_classes, c4, tab, vq);
    else if (out_elem_bytes == 8) CDNET_LAUNCH((k_angle_to_vector<double, double>), grid, 256, 0, st, (const double*)angle, (double*)vec, n, num_classes, c4, tab, vq);
    else CDNET_LAUNCH((k_angle_to_vector<double, float>), grid, 256, 0, st, (const double*)angle, (float*)vec, n, num_classes, c4, tab, vq);
    return last_error();
}

extern "C" int cdnet_vector_to_label(const void* vec, int elem_bytes, int64_t* label, size_t n, int num_classes,
                                     void* stream) {
    if (!vec || !label || n == 0 || !align_classes_ok(num_classes) || (elem_bytes != 4 && elem_bytes != 8)) return CDNET_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = stream_grid(n, 256 * 4);
    const int c4 = num_classes == 4;
    const int vq = aligned16(vec, label, nullptr);
    if (elem_bytes == 4) CDNET_LAUNCH(k_vector_to_label<float>, grid, 256, 0, st, (const float*)vec, (long long*)label, n, num_classes, c4, vq);
    else CDNET_LAUNCH(k_vector_to_label<double>, grid, 256, 0, st, (const double*)vec, (long long*)label, n, num_classes, c4, vq);
    return last_error();
}

extern "C" size_t cdnet_direction_one_hot_workspace_bytes(int B) { return B > 0 ? pad256((size_t)B * 16) : 0; }

extern "C" int cdnet_direction_one_hot(const int64_t* direction, const void* target0, int target_elem_bytes, float* out,
                                       int32_t* status, int B, int C, size_t plane, void* ws, size_t ws_bytes,
                                       void* stream) {
    if (!direction || !target0 || !out || B <= 0 || B > 65535 || C <= 0 || plane == 0) return CDNET_E_BADARG;
    if (target_elem_bytes != 1 && target_elem_bytes != 8) return CDNET_E_BADARG;
    Arena ar(ws, ws_bytes);
    long long* mm = ar.take<long long>((size_t)B * 2);
    if (!ar.ok) return CDNET_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    if (status) CDNET_CUDA_OK(cudaMemsetAsync(status, 0, sizeof(int32_t) * (size_t)B, st));
    const long long* dir = (const long long*)direction;
    CDNET_LAUNCH(k_dir_minmax_init, dim3((B + 255) / 256), 256, 0, st, mm, B);
    CDNET_LAUNCH(k_dir_minmax, dim3(stream_grid(plane, 256 * 8), B), 256, 0, st, dir, mm, plane);
    const bool vec4 = plane % 4 == 0 && ((uintptr_t)out % 16 == 0);
    if (vec4) {
        const dim3 grid(stream_grid(plane / 4, 256), B);
        if (target_elem_bytes == 1) CDNET_LAUNCH((k_dir_one_hot<uint8_t, 4>), grid, 256, 0, st, dir, (const uint8_t*)target0, mm, out, status, C, plane);
        else CDNET_LAUNCH((k_dir_one_hot<long long, 4>), grid, 256, 0, st, dir, (const long long*)target0, mm, out, status, C, plane);
    } else {
        const dim3 grid(stream_grid(plane, 256), B);
        if (target_elem_bytes == 1) CDNET_LAUNCH((k_dir_one_hot<uint8_t, 1>), grid, 256, 0, st, dir, (const uint8_t*)target0, mm, out, status, C, plane);
        else CDNET_LAUNCH((k_dir_one_hot<long long, 1>), grid, 256, 0, st, dir, (const long long*)target0, mm, out, status, C, plane);
    }
    return last_error();
}

extern "C" int cdnet_ternary_label(const uint8_t* ch0, const uint8_t* ch1, int mode, uint8_t* out, int B, int H, int W,
                                   void* stream) {
    if (!ch0 || !out || B <= 0 || B > 65535 || H <= 0 || W <= 0 || (double)H * W >= 2147483648.0 || mode < 0 || mode > 3)
        return CDNET_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (mode != 3 && W % 4 == 0 && (((uintptr_t)ch0 | (uintptr_t)out) & 3u) == 0)
        CDNET_LAUNCH(k_ternary_label4, dim3(ceil_div(W, 256), ceil_div(H, 4), B), dim3(64, 4), 0, st, ch0, mode, out, H, W);
    else
        CDNET_LAUNCH(k_ternary_label, dim3(ceil_div(W, 64), ceil_div(H, 4), B), dim3(64, 4), 0, st, ch0, ch1, mode, out, H, W);
    return last_error();
}
