// metrics.cu -- pair table of two instance-label images: the one reduction behind the reference's instance
// metrics (stats_utils.py:7-98 get_fast_aji, :101-177 get_fast_aji_plus, :182-275 get_fast_pq,
// :279-318 get_fast_dice_2, :324-334 get_dice_1, :338-357 get_dice_2, :361-389 remap_label).
//
// The reference materialises one H x W mask per instance and loops over (true, pred) pairs: O(N * H * W) numpy work,
// seconds per image.  Everything those functions need is the sparse table
//     n[t][q] = #{pixels with true == t and pred == q}           (t, q >= 0; the counts add up to H * W)
// from which the instance areas are row / column sums (the q == 0 / t == 0 entries count the uncovered pixels) and
// the intersections are the entries with t, q > 0.  One pass over the two label images builds the table in a
// hash map in HBM: equal-key runs of a warp row are counted with one ballot, one atomicAdd per run.  The float64
// epilogue (a few thousand pairs) stays on the host, in the reference's own order of operations.
#include "internal.h"

namespace cdnet {

constexpr unsigned long long kEmpty = 0xffffffffffffffffull;

__device__ __forceinline__ uint32_t pair_hash(unsigned long long k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return (uint32_t)k;
}

template <typename T>
__global__ void __launch_bounds__(256) k_pair_count(const T* __restrict__ a, const T* __restrict__ b,
                                                    unsigned long long* __restrict__ hkeys, int* __restrict__ hcnt,
                                                    int* __restrict__ status, size_t plane, uint32_t hmask) {
    const int tile = blockIdx.y;
    const T* A = a + (size_t)tile * plane;
    const T* Bp = b + (size_t)tile * plane;
    unsigned long long* K = hkeys + (size_t)tile * ((size_t)hmask + 1);
    int* C = hcnt + (size_t)tile * ((size_t)hmask + 1);
    const int lane = threadIdx.x & 31;
    const size_t padded = (plane + 31) & ~size_t(31);
    auto insert = [&](unsigned long long key, int run) {
        uint32_t slot = pair_hash(key) & hmask;
        for (uint32_t probe = 0; probe <= hmask; ++probe) {
            const unsigned long long old = atomicCAS(K + slot, kEmpty, key);
            if (old == kEmpty || old == key) {
                atomicAdd(C + slot, run);
                return;
            }
            slot = (slot + 1) & hmask;
        }
        atomicOr(status + tile, CDNET_S_PAIR_OVERFLOW);
    };
    int joint_bg = 0;  // pixels with key (0, 0): by far the most frequent key, counted per thread and added once per warp
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < padded; i += (size_t)gridDim.x * blockDim.x) {
        unsigned long long key = 0ull;
        bool bad = false;
        if (i < plane) {
            const long long t = (long long)A[i], q = (long long)Bp[i];
            bad = t < 0 || q < 0 || t > 0x7fffffffll || q > 0x7fffffffll;
            if (!bad) key = ((unsigned long long)t << 32) | (unsigned long long)q;
            if (!bad && key == 0ull) ++joint_bg;
        }
        if (bad) atomicOr(status + tile, CDNET_S_PAIR_RANGE);
        // run heads inside the warp (32 consecutive pixels): a lane starts a run when its key differs from the
        // previous lane's; the run length is the distance to the next head
        const unsigned long long prev = __shfl_up_sync(0xffffffffu, key, 1);
        const bool head = lane == 0 || prev != key;
        const unsigned heads = __ballot_sync(0xffffffffu, head);
        if (head && key != 0ull) {
            const unsigned above = heads & ~((2u << lane) - 1u);  // heads in higher lanes
            const int run = (above ? __ffs(above) - 1 : 32) - lane;
            insert(key, run);
        }
    }
    joint_bg = __reduce_add_sync(0xffffffffu, joint_bg);
    if (lane == 0 && joint_bg) insert(0ull, joint_bg);
}

// hash slots -> compact (key, count) list per tile (order is arbitrary; the host sorts by key)
__global__ void __launch_bounds__(256) k_pair_compact(const unsigned long long* __restrict__ hkeys,
                                                      const int* __restrict__ hcnt, unsigned long long* __restrict__ keys,
                                                      int* __restrict__ counts, int* __restrict__ n_out,
                                                      int* __restrict__ status, uint32_t hsize, int cap) {
    const int tile = blockIdx.y;
    const unsigned long long* K = hkeys + (size_t)tile * hsize;
    const int* C = hcnt + (size_t)tile * hsize;
    const int lane = threadIdx.x & 31;
    const uint32_t padded = (hsize + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < padded; i += gridDim.x * blockDim.x) {
        const unsigned long long k = i < hsize ? K[i] : kEmpty;
        const bool used = k != kEmpty;
        const unsigned m = __ballot_sync(0xffffffffu, used);
        int base = 0;
        if (lane == 0 && m) base = atomicAdd(n_out + tile, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (used) {
            const int idx = base + __popc(m & ((1u << lane) - 1u));
            if (idx < cap) {
                keys[(size_t)tile * cap + idx] = k;
                counts[(size_t)tile * cap + idx] = C[i];
            } else {
                atomicOr(status + tile, CDNET_S_PAIR_OVERFLOW);
            }
        }
    }
}

// out[i] = new_ids[j] where sorted_ids[j] == in[i] (binary search), 0 where the id is absent or 0
template <typename T>
__global__ void __launch_bounds__(256) k_remap_labels(const T* __restrict__ in, int* __restrict__ out,
                                                      const int* __restrict__ sorted_ids, const int* __restrict__ new_ids,
                                                      int n_ids, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const long long v = (long long)in[i];
        int r = 0;
        if (v != 0) {
            int lo = 0, hi = n_ids - 1;
            while (lo <= hi) {
                const int mid = (lo + hi) >> 1;
                const long long s = (long long)sorted_ids[mid];
                if (s == v) { r = new_ids[mid]; break; }
                if (s < v) lo = mid + 1; else hi = mid - 1;
            }
        }
        out[i] = r;
    }
}

static uint32_t hash_size(int cap) {
    uint32_t h = 64;
    while (h < 2u * (uint32_t)cap && h < (1u << 31)) h <<= 1;
    return h;
}

}  // namespace cdnet

using namespace cdnet;

extern "C" size_t cdnet_label_pairs_workspace_bytes(int B, int cap) {
    if (B <= 0 || cap <= 0 || cap > (1 << 29)) return 0;
    const size_t hs = hash_size(cap);
    return pad256((size_t)B * hs * 8) + pad256((size_t)B * hs * 4);
}

extern "C" int cdnet_label_pairs(const void* true_lab, const void* pred_lab, int elem_bytes, uint64_t* keys,
                                 int32_t* counts, int32_t* n_out, int32_t* status, int B, int H, int W, int cap, void* ws,
                                 size_t ws_bytes, void* stream) {
    if (!true_lab || !pred_lab || !keys || !counts || !n_out || !status || B <= 0 || H <= 0 || W <= 0 || cap <= 0 ||
        cap > (1 << 29) || (elem_bytes != 4 && elem_bytes != 8) || (double)H * W >= 2147483648.0)
        return CDNET_E_BADARG;
    if (ws_bytes < cdnet_label_pairs_workspace_bytes(B, cap)) return CDNET_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t hs = hash_size(cap);
    Arena ar(ws, ws_bytes);
    unsigned long long* hkeys = ar.take<unsigned long long>((size_t)B * hs);
    int* hcnt = ar.take<int>((size_t)B * hs);
    if (!ar.ok) return CDNET_E_WORKSPACE;
    CDNET_CUDA_OK(cudaMemsetAsync(hkeys, 0xff, (size_t)B * hs * 8, st));
    CDNET_CUDA_OK(cudaMemsetAsync(hcnt, 0, (size_t)B * hs * 4, st));
    CDNET_CUDA_OK(cudaMemsetAsync(n_out, 0, (size_t)B * 4, st));
    CDNET_CUDA_OK(cudaMemsetAsync(status, 0, (size_t)B * 4, st));
    const size_t plane = (size_t)H * W;
    size_t gx = (plane + 256 * 8 - 1) / (256 * 8);
    if (gx > 65535) gx = 65535;
    if (elem_bytes == 4)
        CDNET_LAUNCH(k_pair_count<int32_t>, dim3((unsigned)gx, B), 256, 0, st, (const int32_t*)true_lab,
                     (const int32_t*)pred_lab, hkeys, hcnt, status, plane, hs - 1);
    else
        CDNET_LAUNCH(k_pair_count<long long>, dim3((unsigned)gx, B), 256, 0, st, (const long long*)true_lab,
                     (const long long*)pred_lab, hkeys, hcnt, status, plane, hs - 1);
    size_t gc = ((size_t)hs + 256 * 4 - 1) / (256 * 4);
    if (gc > 65535) gc = 65535;
    CDNET_LAUNCH(k_pair_compact, dim3((unsigned)gc, B), 256, 0, st, hkeys, hcnt, (unsigned long long*)keys, counts, n_out,
                 status, hs, cap);
    return last_error();
}

extern "C" int cdnet_remap_labels(const void* in, int elem_bytes, int32_t* out, const int32_t* sorted_ids,
                                  const int32_t* new_ids, int n_ids, size_t n, void* stream) {
    if (!in || !out || n == 0 || n_ids < 0 || (n_ids > 0 && (!sorted_ids || !new_ids)) ||
        (elem_bytes != 4 && elem_bytes != 8))
        return CDNET_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    size_t g = (n + 256 * 8 - 1) / (256 * 8);
    if (g > (1u << 20)) g = 1u << 20;
    if (elem_bytes == 4)
        CDNET_LAUNCH(k_remap_labels<int32_t>, (unsigned)g, 256, 0, st, (const int32_t*)in, out, sorted_ids, new_ids, n_ids, n);
    else
        CDNET_LAUNCH(k_remap_labels<long long>, (unsigned)g, 256, 0, st, (const long long*)in, out, sorted_ids, new_ids,
                     n_ids, n);
    return last_error();
}
