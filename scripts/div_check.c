/* Checks that q' = fma(fma(-v, q, n), y, q) with y = RN(1/v), q = RN(n*y) equals RN(n/v) for
 * float32 (Markstein's correction step), over random operands in the ranges the wall-distance
 * code sees (|v| in (1e-10, 1], n anywhere).  Usage: div_check [millions of pairs] */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t s = 88172645463325252ull;
static inline uint64_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static inline float bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

int main(int argc, char **argv)
{
    long long N = (argc > 1 ? atoll(argv[1]) : 200) * 1000000ll, bad = 0;
    for (long long i = 0; i < N; ++i) {
        uint64_t r = rnd();
        /* v: exponent in [2^-34, 2^0], random mantissa; every 16th: mantissa all ones / zeros / near */
        uint32_t ev = 127 - (uint32_t)(r % 35), mv = (uint32_t)(r >> 8) & 0x7fffff;
        if ((r >> 40) % 16 == 0) mv = 0x7fffff - ((r >> 44) & 3);
        if ((r >> 40) % 16 == 1) mv = (r >> 44) & 3;
        float v = bits((ev << 23) | mv);
        if (v > 1.f) v = 1.f;
        uint64_t r2 = rnd();
        uint32_t en = 127 - 40 + (uint32_t)(r2 % 110), mn = (uint32_t)(r2 >> 8) & 0x7fffff;   /* 2^-40 .. 2^69 */
        float n = bits((en << 23) | mn);
        if (r2 >> 63) n = -n;
        if ((r2 >> 62) & 1) v = -v;
        volatile float y = 1.0f / v;
        volatile float q = n * y;
        float rr = fmaf(-v, q, n);
        float q2 = fmaf(rr, y, q);
        volatile float ref = n / v;
        if (memcmp(&q2, (const void *)&ref, 4) != 0 && !(isinf(ref) || isinf(q2) || ref == 0.f)) {
            if (bad < 10) printf("MISMATCH n=%a v=%a q2=%a ref=%a\n", n, v, q2, ref);
            ++bad;
        }
    }
    printf("pairs %lld mismatches %lld\n", N, bad);
    return bad != 0;
}
