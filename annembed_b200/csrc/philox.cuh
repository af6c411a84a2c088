// Philox4x32-10 counter-based RNG (Salmon et al., SC'11), written out for host and device.
// Keyed by the user seed; the counter is (edge_lo, edge_hi, epoch, sub-stream), so any rank can
// replay the draws of any edge in any mini-epoch without communication.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace annembed {

struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ void philox_mulhilo(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo)
{
#ifdef __CUDA_ARCH__
    lo = a * b;
    hi = __umulhi(a, b);
#else
    uint64_t p = (uint64_t)a * (uint64_t)b;
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
#endif
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0, lo0, hi1, lo1;
        philox_mulhilo(M0, c0, hi0, lo0);
        philox_mulhilo(M1, c2, hi1, lo1);
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

struct Philox2 { uint32_t x, y; };

// Philox2x32-10: half the work of the 4x32 variant for 64 random bits (per-node uniforms, replayed by in-edge owners)
__host__ __device__ __forceinline__ Philox2 philox2x32_10(uint32_t c0, uint32_t c1, uint32_t k)
{
    const uint32_t M = 0xD256D193u, W = 0x9E3779B9u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi, lo;
        philox_mulhilo(M, c0, hi, lo);
        c0 = hi ^ k ^ c1;
        c1 = lo;
        k += W;
    }
    Philox2 o; o.x = c0; o.y = c1;
    return o;
}

// 24-bit uniform in [0,1), exact in fp32
__host__ __device__ __forceinline__ float u01_24(uint32_t w) { return (float)(w >> 8) * (1.0f / 16777216.0f); }

// map a 32-bit word to [0,n) (multiply-shift; bias <= n / 2^32)
__host__ __device__ __forceinline__ uint32_t below32(uint32_t w, uint32_t n)
{
#ifdef __CUDA_ARCH__
    return __umulhi(w, n);
#else
    return (uint32_t)(((uint64_t)w * (uint64_t)n) >> 32);
#endif
}

// The same with 8 more random bits below the word (a 40-bit multiply-shift): bias <= n / 2^40.  Used when n / 2^32 would be
// visible (large graphs: 10^8 nodes are 0.6 % of 2^32 by sectors, 2.3 % by nodes; the reference draws exactly uniform,
// embedder.rs:1117-1123).  floor(((w << 8 | b) * n) / 2^40) = floor((w * n + floor(b * n / 2^8)) / 2^32), no overflow for n < 2^32.
__host__ __device__ __forceinline__ uint32_t below40(uint32_t w, uint32_t low8, uint32_t n)
{
    const uint64_t p = (uint64_t)w * (uint64_t)n + (((uint64_t)(low8 & 0xFFu) * (uint64_t)n) >> 8);
    return (uint32_t)(p >> 32);
}
#define ANNEMBED_WIDE_DRAW_ABOVE (1u << 20)     // ranges above this use below40 (bias of below32 there: > 2^-12)
__host__ __device__ __forceinline__ uint32_t below_auto(uint32_t w, uint32_t low8, uint32_t n)
{
    return n > ANNEMBED_WIDE_DRAW_ABOVE ? below40(w, low8, n) : below32(w, n);
}

} // namespace annembed
