/* Brute-force check of vx_accum_jump against the serial chain.  Built and run by tests/test_jump.py. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "../differential_projection_voxel_renderer_b200/csrc/vx_jump.h"

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd(void) {
    rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
    return (uint32_t)(rng_state >> 16);
}
static float rnd_float(int mode) {
    uint32_t u = rnd();
    switch (mode) {
    case 0: return vx_u2f(u);                                               /* any bit pattern */
    case 1: return vx_u2f((u & 0x807FFFFFu) | ((100u + (rnd() % 56u)) << 23)); /* moderate exponents */
    case 2: return (float)((int32_t)(rnd() % 2001) - 1000) / 1024.0f;          /* small dyadic values */
    case 3: return vx_u2f((u & 0x807FFFFFu) | ((126u) << 23));                 /* [0.5, 1) like NDC depth */
    default: return vx_u2f((u & 0x80000007u) | ((90u + (rnd() % 40u)) << 23)); /* few mantissa bits: many ties */
    }
}

int main(int argc, char **argv) {
    long cases = argc > 1 ? atol(argv[1]) : 2000000;
    long bad = 0, steps_total = 0;
    for (long c = 0; c < cases; ++c) {
        float z = rnd_float((int)(rnd() % 5u));
        float s = rnd_float((int)(rnd() % 5u));
        if (rnd() % 4u == 0) s = z * vx_u2f((rnd() & 0x807FFFFFu) | ((100u + rnd() % 27u) << 23)); /* s relative to z */
        if (rnd() % 16u == 0) s = -z / (float)(1 + rnd() % 64u);                                    /* walks through zero */
        uint32_t n = rnd() % 4096u;
        volatile float ref = z;
        /* compare at several intermediate counts too */
        uint32_t checks[4] = {n / 7u, n / 3u, n - (n > 0), n};
        uint32_t done = 0;
        for (int k = 0; k < 4; ++k) {
            uint32_t target = checks[k];
            if (target < done) continue;
            for (; done < target; ++done) ref = ref + s;
            float got = vx_accum_jump(z, s, target);
            float r = ref;
            if (vx_f2u(got) != vx_f2u(r) && !(got != got && r != r)) {
                if (bad < 10) printf("MISMATCH z=%a s=%a n=%u serial=%a jump=%a\n", z, s, target, r, got);
                bad++;
            }
        }
        steps_total += n;
    }
    printf("cases=%ld serial_steps=%ld mismatches=%ld\n", cases, steps_total, bad);
    return bad ? 1 : 0;
}
