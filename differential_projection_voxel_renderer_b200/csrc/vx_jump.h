// vx_jump.h -- exact fast-forward of a serial f32 accumulation.
//
// The reference span loop advances its interpolants with one rounded f32 add per pixel
// (z += step; rasterizer.rs:1458-1461).  That recurrence is not associative, so a pixel's value depends on
// the whole chain from the span's first pixel.  vx_accum_jump(z, s, n) returns exactly the value the chain
// z <- fl(z + s) has after n steps, in O(number of binades crossed) instead of O(n):
//
//   While z stays inside one binade (same sign and exponent, normal), representable values form a uniform
//   lattice of spacing U = ulp(z) and s = q*U + r (0 <= r < U), so fl(z + s) moves a constant number of lattice
//   points per step: q (r < U/2), q+1 (r > U/2), and for the tie r = U/2 round-to-even makes the mantissa even
//   after the first step, after which the increment is constant too.  Three serial steps inside a binade
//   therefore expose the settled increment d (as a difference of raw bit patterns); the chain is then advanced
//   by k*d in one integer operation for as long as the next landing point stays strictly inside the binade
//   (fraction field in [1, 0x7FFFFF], which also keeps the exact sum inside the binade's rounding domain);
//   boundary crossings, zero, subnormals, infinities and NaN are left to real serial adds.
//
// Verified bit-for-bit against the serial loop by tests/test_jump.py (hundreds of millions of random chains
// incl. ties, sign changes, subnormals) and on the device by the frame parity tests.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define VX_HD __host__ __device__ __forceinline__
#else
#define VX_HD static inline
#endif

VX_HD uint32_t vx_f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}
VX_HD float vx_u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

// one rounded add that the compiler may not contract or re-associate
VX_HD float vx_add_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b;
    return r;
#endif
}

VX_HD float vx_accum_jump(float z, float s, uint32_t n) {
    while (n > 0) {
        const float z1 = vx_add_rn(z, s);
        if (--n == 0) return z1;
        const float z2 = vx_add_rn(z1, s);
        if (--n == 0) return z2;
        const float z3 = vx_add_rn(z2, s);
        if (--n == 0) return z3;
        const uint32_t b1 = vx_f2u(z1), b2 = vx_f2u(z2), b3 = vx_f2u(z3);
        const uint32_t e3 = (b3 >> 23) & 0xFFu;
        // z1, z2, z3 must share sign and exponent (one binade) and be normal numbers
        if ((((b1 ^ b3) | (b2 ^ b3)) >> 23) != 0 || e3 == 0 || e3 == 0xFFu) {
            z = z3;
            continue;
        }
        const int32_t d = (int32_t)(b3 - b2); // settled lattice increment (raw bits grow with magnitude)
        if (d == 0) return z3;                // s no longer moves z: the chain is constant from here on
        const uint32_t frac = b3 & 0x7FFFFFu;
        const uint32_t room = d > 0 ? 0x7FFFFFu - frac : (frac >= 1u ? frac - 1u : 0u);
        const uint32_t ad = d > 0 ? (uint32_t)d : (uint32_t)(-d);
        // common case: all n remaining steps stay inside the binade (no division needed)
        if ((unsigned long long)n * ad <= (unsigned long long)room) return vx_u2f(b3 + (uint32_t)((int32_t)n * d));
        const uint32_t kmax = room / ad;
        const uint32_t k = n < kmax ? n : kmax;
        z = vx_u2f(b3 + (uint32_t)((int32_t)k * d));
        n -= k;
    }
    return z;
}
