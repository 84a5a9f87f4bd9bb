// cuda_fp16.h -- SIMT emulation shim (test infrastructure only; see cuda_runtime.h in this directory).
#pragma once
#include <stdint.h>
#include <string.h>

struct __half { uint16_t x; };

// double -> binary16, round to nearest even, done directly on the double's bits (no double rounding via float)
static inline __half __double2half(double d) {
    uint64_t b;
    memcpy(&b, &d, 8);
    const uint16_t sign = (uint16_t)((b >> 48) & 0x8000u);
    const int64_t e = (int64_t)((b >> 52) & 0x7ff);
    uint64_t m = b & 0xfffffffffffffull;
    __half h;
    if (e == 0x7ff) { h.x = (uint16_t)(sign | 0x7c00u | (m ? 0x200u : 0u)); return h; }
    if (e == 0 && m == 0) { h.x = sign; return h; }
    const int64_t E = e - 1023;  // unbiased (subnormal doubles are far below half's range anyway)
    if (E > 15) { h.x = (uint16_t)(sign | 0x7c00u); return h; }
    m |= (1ull << 52);  // implicit one (e == 0 never reaches the rounding below with a meaningful value)
    int shift;           // number of low bits of the 53-bit significand to drop
    int64_t he;          // biased half exponent of the result before rounding carries
    if (E >= -14) { shift = 42; he = E + 15; }
    else { shift = 42 + (int)(-14 - E); he = 0; }
    if (shift > 63 || e == 0) { h.x = sign; return h; }
    const uint64_t kept = m >> shift;
    const uint64_t rem = m & ((1ull << shift) - 1ull);
    const uint64_t halfway = 1ull << (shift - 1);
    uint64_t r = kept;
    if (rem > halfway || (rem == halfway && (kept & 1ull))) r += 1;
    // normal: r has the implicit bit at position 10; adding (he-1)<<10 makes carries roll into the exponent
    uint64_t out = (he > 0) ? (((uint64_t)(he - 1) << 10) + r) : r;
    if (out >= 0x7c00u) out = 0x7c00u;
    h.x = (uint16_t)(sign | (uint16_t)out);
    return h;
}
static inline __half __float2half(float f) { return __double2half((double)f); }
static inline float __half2float(__half h) {
    const uint32_t s = (uint32_t)(h.x & 0x8000u) << 16;
    uint32_t e = (h.x >> 10) & 0x1fu, m = h.x & 0x3ffu, u;
    if (e == 0) {
        if (m == 0) u = s;
        else {
            int k = 0;
            while (!(m & 0x400u)) { m <<= 1; ++k; }
            u = s | ((uint32_t)(127 - 15 - k + 1) << 23) | ((m & 0x3ffu) << 13);
        }
    } else if (e == 31) u = s | 0x7f800000u | (m << 13);
    else u = s | ((e + 112u) << 23) | (m << 13);
    float f;
    memcpy(&f, &u, 4);
    return f;
}
