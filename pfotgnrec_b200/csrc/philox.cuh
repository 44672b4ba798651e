// Philox4x32-10 (Salmon et al. SC'11) -- the random stream shared with oracle/philox.py.
#pragma once
#include <stdint.h>

#define PFO_PURPOSE_NEG 1u
#define PFO_PURPOSE_NEG_REPL 2u
#define PFO_PURPOSE_NBR 3u
#define PFO_PURPOSE_DROPOUT 4u
#define PFO_PURPOSE_NEG_SEQ 5u

struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0;
        uint64_t p1 = (uint64_t)M1 * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

// floor(x * n / 2^32): a uint32 draw mapped onto [0, n)
__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t x, uint32_t n) {
    return (uint32_t)(((uint64_t)x * (uint64_t)n) >> 32);
}
