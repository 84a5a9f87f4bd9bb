// seam.cu -- device-side reconciliation of components that straddle the seams of a row-sharded slide
// (no reference counterpart: the reference post-processes a slide on one CPU).
//
// Every rank describes the two rows it shares with each neighbour as a compact TABLE of run starts
//     (gid, nb, attr, sum)      gid = slide-global pixel index of the LOCAL root of the pixel,
//                               nb  = the neighbour rank's gid of the same pixel (bottom side only, else -1),
//                               attr / sum = a per-root attribute (sum is non-zero once per local root)
// (cdnet_seam_export).  The tables of all ranks are all-gathered by the host plumbing (NCCL, one call, no host
// synchronisation) and every rank then solves the same small union on its own device (cdnet_seam_solve):
// gids are hashed to slots, the (gid, nb) pairs are united (the class root is the smallest gid = the class's
// first pixel in the slide), class attributes are reduced (OR of frame-touch flags, SUM of areas) and written
// back at the rank's local root pixels.  The numbering step hands out the ids of seam classes the same way
// (cdnet_seam_ids_export / cdnet_seam_ids_apply).  Nothing in a round waits for the host.
#include "internal.h"

namespace cdnet {

struct SeamScratch {
    int* keys;    // [HS] gid or -1
    int* parent;  // [HS]
    int* cattr;   // [HS]
    int* slot_g;  // [E]
    int* slot_nb; // [E]
    int HS;
};

__device__ __forceinline__ unsigned seam_hash(int key, int mask) {
    return ((unsigned)key * 2654435761u >> 7) & (unsigned)mask;
}

__device__ __forceinline__ int seam_insert(int* keys, int* parent, int* cattr, int mask, int key) {
    unsigned h = seam_hash(key, mask);
    for (;;) {
        const int old = atomicCAS(keys + h, -1, key);
        if (old == -1) { parent[h] = (int)h; cattr[h] = 0; return (int)h; }
        if (old == key) return (int)h;
        h = (h + 1) & (unsigned)mask;
    }
}

__device__ __forceinline__ int seam_lookup(const int* keys, int mask, int key) {
    unsigned h = seam_hash(key, mask);
    for (;;) {
        const int k = keys[h];
        if (k == key) return (int)h;
        if (k == -1) return -1;
        h = (h + 1) & (unsigned)mask;
    }
}

__device__ __forceinline__ int seam_find(const int* parent, int s) {
    int q = __ldcg(parent + s);
    while (q != s) { s = q; q = __ldcg(parent + s); }
    return s;
}

// ---- export -----------------------------------------------------------------------------------------------
// tbl: int4 [cap]; row 0 = (count, error flags, 0, 0); rows 1..count = entries in arbitrary order
__global__ void __launch_bounds__(256) k_seam_export(const int* __restrict__ L, const uint8_t* __restrict__ valid,
                                                     const int* __restrict__ attr, int* __restrict__ emitted, int round_id,
                                                     int off, int He, int W, int has_top, int has_bottom,
                                                     const int* __restrict__ nb_gid, int4* __restrict__ tbl, int cap) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool emit = false;
    int gid = -1, nb = -1, a = 0, asum = 0;
    if (idx < 4 * W) {
        const int side = idx / (2 * W), r = (idx / W) & 1, x = idx % W;
        if ((side == 0 && has_top) || (side == 1 && has_bottom)) {
            const int row = side == 0 ? r : He - 2 + r;
            const int p = row * W + x;
            auto gid_at = [&](int q) -> int { return (!valid || valid[q]) ? L[q] + off : -1; };
            auto nb_at = [&](int xx) -> int { return (side == 1 && nb_gid) ? nb_gid[r * W + xx] : -1; };
            gid = gid_at(p);
            nb = nb_at(x);
            if (gid >= 0) {
                emit = true;
                if (x > 0 && gid_at(p - 1) == gid && nb_at(x - 1) == nb) emit = false;  // same run as the left pixel
                if (side == 1 && nb_gid && nb < 0) atomicOr(&tbl[0].y, 1);  // classified differently on the two ranks
            } else if (nb >= 0) {
                atomicOr(&tbl[0].y, 1);
            }
            if (emit) {
                const int root = gid - off;
                a = attr ? attr[root] : 0;
                // the SUM column carries a root's attribute exactly once per rank and round
                asum = (atomicExch(emitted + root, round_id) != round_id) ? a : 0;
            }
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, emit);
    int base = 0;
    if (lane == 0 && m) base = atomicAdd(&tbl[0].x, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (emit) {
        const int pos = 1 + base + __popc(m & ((1u << lane) - 1));
        if (pos < cap) tbl[pos] = make_int4(gid, nb, a, asum);
        else atomicOr(&tbl[0].y, 2);
    }
}

// ---- solve ------------------------------------------------------------------------------------------------
__global__ void k_seam_insert(const int4* __restrict__ G, int nranks, int cap, SeamScratch sc) {
    const int mask = sc.HS - 1;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nranks * cap; e += gridDim.x * blockDim.x) {
        const int r = e / cap, i = e % cap;
        if (i == 0 || i > min(G[(size_t)r * cap].x, cap - 1)) continue;
        const int4 t = G[e];
        sc.slot_g[e] = seam_insert(sc.keys, sc.parent, sc.cattr, mask, t.x);
        sc.slot_nb[e] = t.y >= 0 ? seam_insert(sc.keys, sc.parent, sc.cattr, mask, t.y) : -1;
    }
}

__global__ void k_seam_union(const int4* __restrict__ G, int nranks, int cap, SeamScratch sc) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nranks * cap; e += gridDim.x * blockDim.x) {
        const int r = e / cap, i = e % cap;
        if (i == 0 || i > min(G[(size_t)r * cap].x, cap - 1)) continue;
        int a = sc.slot_g[e], b = sc.slot_nb[e];
        if (b < 0) continue;
        for (;;) {
            a = seam_find(sc.parent, a);
            b = seam_find(sc.parent, b);
            if (a == b) break;
            if (sc.keys[a] < sc.keys[b]) { const int t = a; a = b; b = t; }  // the smaller gid becomes the root
            if (atomicCAS(sc.parent + a, a, b) == a) break;
        }
    }
}

// mode 0: OR of attr, 1: SUM of the sum column
__global__ void k_seam_reduce(const int4* __restrict__ G, int nranks, int cap, SeamScratch sc, int mode) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nranks * cap; e += gridDim.x * blockDim.x) {
        const int r = e / cap, i = e % cap;
        if (i == 0 || i > min(G[(size_t)r * cap].x, cap - 1)) continue;
        const int4 t = G[e];
        const int root = seam_find(sc.parent, sc.slot_g[e]);
        if (mode == 0) { if (t.z) atomicOr(sc.cattr + root, 1); }
        else if (t.w) {
            // saturating add (areas only matter against min_area)
            const int old = atomicAdd(sc.cattr + root, t.w);
            if (old < 0 || old + t.w < 0) atomicExch(sc.cattr + root, 0x7fffffff);
        }
    }
}

// mode 0/1: plane[root pixel] = class attribute; mode 2: excluded[root pixel] = 1 unless this rank owns the class
// (the class's first pixel = smallest gid lies in the rank's own rows and is this very root)
__global__ void k_seam_scatter(const int4* __restrict__ G, int cap, int my_rank, SeamScratch sc, int mode, int off,
                               int own_lo, int own_hi, int* __restrict__ plane, uint8_t* __restrict__ excluded) {
    const int4* T = G + (size_t)my_rank * cap;
    const int n = min(T[0].x, cap - 1);
    for (int i = 1 + blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += gridDim.x * blockDim.x) {
        const int gid = T[i].x;
        const int root = seam_find(sc.parent, sc.slot_g[(size_t)my_rank * cap + i]);
        if (mode == 2) {
            if (sc.keys[root] != gid || gid < own_lo || gid >= own_hi) excluded[gid - off] = 1;
        } else {
            plane[gid - off] = sc.cattr[root];
        }
    }
}

// owners publish (class root gid, final id); ids: [1] offsets per rank on the device
__global__ void k_seam_ids_export(const int4* __restrict__ G, int cap, int my_rank, SeamScratch sc, int off, int own_lo,
                                  int own_hi, const int* __restrict__ idmap, int* __restrict__ emitted, int round_id,
                                  int4* __restrict__ tbl2) {
    const int4* T = G + (size_t)my_rank * cap;
    const int n = min(T[0].x, cap - 1);
    for (int i = 1 + blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += gridDim.x * blockDim.x) {
        const int gid = T[i].x;
        const int root = seam_find(sc.parent, sc.slot_g[(size_t)my_rank * cap + i]);
        if (sc.keys[root] == gid && gid >= own_lo && gid < own_hi) {
            if (atomicExch(emitted + (gid - off), round_id) != round_id) {
                const int pos = 1 + atomicAdd(&tbl2[0].x, 1);
                if (pos < cap) tbl2[pos] = make_int4(gid, idmap[gid - off], 0, 0);
                else atomicOr(&tbl2[0].y, 2);
            }
        }
    }
}

__global__ void k_seam_ids_store(const int4* __restrict__ G2, int nranks, int cap, SeamScratch sc) {
    const int mask = sc.HS - 1;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nranks * cap; e += gridDim.x * blockDim.x) {
        const int r = e / cap, i = e % cap;
        if (i == 0 || i > min(G2[(size_t)r * cap].x, cap - 1)) continue;
        const int4 t = G2[e];
        const int s = seam_lookup(sc.keys, mask, t.x);
        if (s >= 0) sc.cattr[s] = t.y;
    }
}

__global__ void k_seam_ids_apply(const int4* __restrict__ G, int cap, int my_rank, SeamScratch sc, int off,
                                 int* __restrict__ idmap) {
    const int4* T = G + (size_t)my_rank * cap;
    const int n = min(T[0].x, cap - 1);
    for (int i = 1 + blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += gridDim.x * blockDim.x) {
        const int gid = T[i].x;
        const int root = seam_find(sc.parent, sc.slot_g[(size_t)my_rank * cap + i]);
        idmap[gid - off] = sc.cattr[root];
    }
}

static int hash_size(int nranks, int cap) {
    size_t need = (size_t)nranks * cap * 4;  // <= 2 keys per entry, load factor <= 0.5
    size_t hs = 1024;
    while (hs < need) hs <<= 1;
    return (int)hs;
}

static bool carve(void* ws, size_t ws_bytes, int nranks, int cap, SeamScratch* sc) {
    Arena ar(ws, ws_bytes);
    sc->HS = hash_size(nranks, cap);
    sc->keys = ar.take<int>(sc->HS);
    sc->parent = ar.take<int>(sc->HS);
    sc->cattr = ar.take<int>(sc->HS);
    sc->slot_g = ar.take<int>((size_t)nranks * cap);
    sc->slot_nb = ar.take<int>((size_t)nranks * cap);
    return ar.ok;
}

}  // namespace cdnet

using namespace cdnet;

extern "C" size_t cdnet_seam_workspace_bytes(int nranks, int cap) {
    if (nranks <= 0 || cap <= 1 || (double)nranks * cap * 4 >= 1073741824.0) return 0;
    const size_t hs = (size_t)hash_size(nranks, cap);
    return 3 * pad256(hs * 4) + 2 * pad256((size_t)nranks * cap * 4);
}

// tbl must hold cap int4 rows; emitted: int32 [He,W] zeroed once per slide; round ids must be distinct and > 0
extern "C" int cdnet_seam_export(const int32_t* L, const uint8_t* valid, const int32_t* attr, int32_t* emitted,
                                 int round_id, int off, int He, int W, int has_top, int has_bottom, const int32_t* nb_gid,
                                 int32_t* tbl, int cap, void* stream) {
    if (!L || !emitted || !tbl || He < 2 || W <= 0 || cap < 2 || round_id <= 0) return CDNET_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    CDNET_CUDA_OK(cudaMemsetAsync(tbl, 0, 16, st));
    CDNET_LAUNCH(k_seam_export, ceil_div(4 * W, 256), 256, 0, st, L, valid, attr, emitted, round_id, off, He, W, has_top,
                 has_bottom, nb_gid, (int4*)tbl, cap);
    return last_error();
}

// gathered: int4 [nranks, cap].  mode 0: plane <- OR(attr) per class; 1: plane <- SUM per class; 2: excluded marks the
// seam roots this rank does not own (the hash stays valid for the ids step that follows).
extern "C" int cdnet_seam_solve(const int32_t* gathered, int nranks, int cap, int my_rank, int mode, int off, int own_lo,
                                int own_hi, int32_t* plane, uint8_t* excluded, void* ws, size_t ws_bytes, void* stream) {
    if (!gathered || nranks <= 0 || cap < 2 || my_rank < 0 || my_rank >= nranks || mode < 0 || mode > 2)
        return CDNET_E_BADARG;
    if ((mode < 2 && !plane) || (mode == 2 && !excluded)) return CDNET_E_BADARG;
    SeamScratch sc;
    if (!carve(ws, ws_bytes, nranks, cap, &sc)) return CDNET_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int4* G = (const int4*)gathered;
    CDNET_CUDA_OK(cudaMemsetAsync(sc.keys, 0xff, sizeof(int) * (size_t)sc.HS, st));
    const int nb = ceil_div(nranks * cap, 256) > 2048 ? 2048 : ceil_div(nranks * cap, 256);
    CDNET_LAUNCH(k_seam_insert, nb, 256, 0, st, G, nranks, cap, sc);
    CDNET_LAUNCH(k_seam_union, nb, 256, 0, st, G, nranks, cap, sc);
    if (mode < 2) CDNET_LAUNCH(k_seam_reduce, nb, 256, 0, st, G, nranks, cap, sc, mode);
    const int nb1 = ceil_div(cap, 256) > 1024 ? 1024 : ceil_div(cap, 256);
    CDNET_LAUNCH(k_seam_scatter, nb1, 256, 0, st, G, cap, my_rank, sc, mode, off, own_lo, own_hi, plane, excluded);
    return last_error();
}

// after cdnet_seam_solve(mode 2) and the local numbering: publish the final ids of the seam classes this rank owns
extern "C" int cdnet_seam_ids_export(const int32_t* gathered, int nranks, int cap, int my_rank, int off, int own_lo,
                                     int own_hi, const int32_t* idmap, int32_t* emitted, int round_id, int32_t* tbl2,
                                     void* ws, size_t ws_bytes, void* stream) {
    if (!gathered || !idmap || !emitted || !tbl2) return CDNET_E_BADARG;
    SeamScratch sc;
    if (!carve(ws, ws_bytes, nranks, cap, &sc)) return CDNET_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    CDNET_CUDA_OK(cudaMemsetAsync(tbl2, 0, 16, st));
    const int nb1 = ceil_div(cap, 256) > 1024 ? 1024 : ceil_div(cap, 256);
    CDNET_LAUNCH(k_seam_ids_export, nb1, 256, 0, st, (const int4*)gathered, cap, my_rank, sc, off, own_lo, own_hi, idmap,
                 emitted, round_id, (int4*)tbl2);
    return last_error();
}

extern "C" int cdnet_seam_ids_apply(const int32_t* gathered, const int32_t* gathered2, int nranks, int cap, int my_rank,
                                    int off, int32_t* idmap, void* ws, size_t ws_bytes, void* stream) {
    if (!gathered || !gathered2 || !idmap) return CDNET_E_BADARG;
    SeamScratch sc;
    if (!carve(ws, ws_bytes, nranks, cap, &sc)) return CDNET_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = ceil_div(nranks * cap, 256) > 2048 ? 2048 : ceil_div(nranks * cap, 256);
    CDNET_LAUNCH(k_seam_ids_store, nb, 256, 0, st, (const int4*)gathered2, nranks, cap, sc);
    const int nb1 = ceil_div(cap, 256) > 1024 ? 1024 : ceil_div(cap, 256);
    CDNET_LAUNCH(k_seam_ids_apply, nb1, 256, 0, st, (const int4*)gathered, cap, my_rank, sc, off, idmap);
    return last_error();
}
