// selftest.cu -- kernels that exercise the SIMT emulator itself (tests/test_simt_emulator.py).
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace cdnet { unsigned long long g_launches = 0; }

// block reduction with barriers + a warp shuffle tail: correct under every schedule
__global__ void k_block_sum(const int* in, int* out, int n) {
    __shared__ int s[256];
    const int t = threadIdx.x;
    int v = 0;
    for (int i = blockIdx.x * blockDim.x + t; i < n; i += gridDim.x * blockDim.x) v += in[i];
    s[t] = v;
    __syncthreads();
    for (int o = 128; o >= 32; o >>= 1) {
        if (t < o) s[t] += s[t + o];
        __syncthreads();
    }
    if (t < 32) {
        int w = s[t];
        for (int o = 16; o; o >>= 1) w += __shfl_down_sync(0xffffffffu, w, o);
        if (t == 0) atomicAdd(out, w);
    }
}

// the same neighbour exchange WITHOUT the barrier: the result depends on the order threads run in
__global__ void k_racy_shift(const int* in, int* out) {
    __shared__ int s[64];
    const int t = threadIdx.x;
    s[t] = 0;
    __syncthreads();
    s[t] = in[t];
    // missing __syncthreads()
    out[t] = s[(t + 1) & 63];
}

// early exits, partial masks, ballot / match / reduce
__global__ void k_warp_ops(unsigned* out) {
    const int t = threadIdx.x;
    if (t >= 40) return;                       // lanes 8..31 of warp 1 leave before any collective
    const unsigned full = t < 32 ? 0xffffffffu : 0x000000ffu;
    const unsigned b = __ballot_sync(full, (t & 1) == 0);
    const unsigned m = __match_any_sync(full, t % 3);
    const unsigned r = __reduce_max_sync(full, (unsigned)(t * 7 % 13));
    unsigned sub = 0;
    if ((t & 31) < 8) sub = __reduce_add_sync(0xffu, (unsigned)t);  // a sub-warp group with its own mask
    out[t * 4 + 0] = b; out[t * 4 + 1] = m; out[t * 4 + 2] = r; out[t * 4 + 3] = sub;
}

__global__ void k_half(const double* in, uint16_t* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __double2half(in[i]).x;
}

__global__ void k_bits(const unsigned* a, const unsigned* b, const unsigned* s, unsigned* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[4 * i + 0] = __byte_perm(a[i], b[i], s[i]);
    out[4 * i + 1] = __funnelshift_l(a[i], b[i], s[i]);
    out[4 * i + 2] = __funnelshift_r(a[i], b[i], s[i]);
    out[4 * i + 3] = (unsigned)__clz((int)a[i]) | ((unsigned)__ffs((int)b[i]) << 8) | ((unsigned)__popc(s[i]) << 16);
}

__global__ void k_deadlock() {
    if (threadIdx.x == 0) return;
    if (threadIdx.x & 1) __syncthreads();      // only the odd threads arrive: never completes
    else __syncwarp(0xffffffffu);
}

extern "C" void st_block_sum(const int* in, int* out, int n) { CDNET_LAUNCH(k_block_sum, 7, 256, 0, 0, in, out, n); }
extern "C" void st_racy_shift(const int* in, int* out) { CDNET_LAUNCH(k_racy_shift, 1, 64, 0, 0, in, out); }
extern "C" void st_warp_ops(unsigned* out) { CDNET_LAUNCH(k_warp_ops, 1, 64, 0, 0, out); }
extern "C" void st_half(const double* in, uint16_t* out, int n) { CDNET_LAUNCH(k_half, (n + 127) / 128, 128, 0, 0, in, out, n); }
extern "C" void st_bits(const unsigned* a, const unsigned* b, const unsigned* s, unsigned* out, int n) {
    CDNET_LAUNCH(k_bits, (n + 127) / 128, 128, 0, 0, a, b, s, out, n);
}
extern "C" void st_deadlock() { CDNET_LAUNCH(k_deadlock, 1, 64, 0, 0); }
