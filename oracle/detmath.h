/*
 * oracle/detmath.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Deterministic float32 transcendental functions for the CPU oracle.
 *
 * The reference (photon_mod.f90, vector_mod.f90) calls the Fortran intrinsics
 * log, cos, sin, acos, atan on default REAL (float32).  Their last-bit results
 * depend on the libm in use, so two builds of the reference already differ at
 * the ulp level.  To make oracle <-> CUDA-kernel parity *bit exact* both sides
 * evaluate the same explicitly specified algorithm using only IEEE-754
 * correctly-rounded operations (+ - * / sqrt, int<->float conversions) in
 * double precision, and round once to float at the end.  No FMA may be
 * contracted: compile with -ffp-contract=off (host) / -fmad=false (device).
 *
 * Specification (shared, in words, with mocassin_b200/csrc/detmath.cuh):
 *   log : x = m * 2^e with m in [sqrt(1/2), sqrt(2)); s = (m-1)/(m+1);
 *         log(x) = e*ln2 + 2*(s + s^3/3 + ... + s^15/15), Horner in s^2.
 *   sincos: q = nearest integer to x*(2/pi); r = x - q*(pi/2) (double);
 *         Taylor series for sin (to r^15) and cos (to r^16), Horner in r^2;
 *         quadrant from q mod 4.   Valid for |x| <= 1e4 (callers: |x| <= 2 pi).
 *   atan: reduce |t| > 1 by 1/t, then t > tan(pi/8) by (t-1)/(t+1);
 *         odd Taylor series to u^27.
 *   acos: acos(x) = 2*atan(sqrt((1-x)/(1+x))), acos(-1) = pi, |x|>1 clamped.
 * Each result is within 1 ulp (float) of the true value; tests/test_detmath.py
 * checks that against libm.
 */
#ifndef ORACLE_DETMATH_H
#define ORACLE_DETMATH_H

#include <stdint.h>
#include <string.h>
#include <math.h>

#define DM_PI      3.14159265358979323846
#define DM_PIO2    1.57079632679489661923
#define DM_PIO4    0.78539816339744830962
#define DM_2OPI    0.63661977236758134308
#define DM_LN2     0.69314718055994530942
#define DM_TANPIO8 0.41421356237309504880

static inline double dm_log_d(float xf)
{
    /* xf must be a positive normal float */
    uint32_t ix;
    memcpy(&ix, &xf, 4);
    int e = (int)(ix >> 23) - 127;
    uint32_t im = (ix & 0x007fffffu) | 0x3f800000u; /* m in [1,2) */
    float mf;
    memcpy(&mf, &im, 4);
    double m = (double)mf;
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    double s = (m - 1.0) / (m + 1.0);
    double z = s * s;
    double p = 1.0 / 15.0;
    p = p * z + 1.0 / 13.0;
    p = p * z + 1.0 / 11.0;
    p = p * z + 1.0 / 9.0;
    p = p * z + 1.0 / 7.0;
    p = p * z + 1.0 / 5.0;
    p = p * z + 1.0 / 3.0;
    p = p * z + 1.0;
    return (double)e * DM_LN2 + 2.0 * (s * p);
}

static inline float dm_logf(float x) { return (float)dm_log_d(x); }

static inline void dm_sincosf(float xf, float *sn, float *cs)
{
    double x = (double)xf;
    double t = x * DM_2OPI;
    /* round half away from zero; exact integer arithmetic afterwards */
    int q = (int)(t >= 0.0 ? t + 0.5 : t - 0.5);
    double r = x - (double)q * DM_PIO2;
    double z = r * r;
    /* sin(r) = r*(1 - z/3! + z^2/5! - ... ) */
    double ps = -1.0 / 1307674368000.0;      /* 1/15! */
    ps = ps * z + 1.0 / 6227020800.0;        /* 1/13! */
    ps = ps * z - 1.0 / 39916800.0;          /* 1/11! */
    ps = ps * z + 1.0 / 362880.0;            /* 1/9!  */
    ps = ps * z - 1.0 / 5040.0;              /* 1/7!  */
    ps = ps * z + 1.0 / 120.0;               /* 1/5!  */
    ps = ps * z - 1.0 / 6.0;                 /* 1/3!  */
    ps = ps * z + 1.0;
    double sr = r * ps;
    double pc = 1.0 / 20922789888000.0;      /* 1/16! */
    pc = pc * z - 1.0 / 87178291200.0;       /* 1/14! */
    pc = pc * z + 1.0 / 479001600.0;         /* 1/12! */
    pc = pc * z - 1.0 / 3628800.0;           /* 1/10! */
    pc = pc * z + 1.0 / 40320.0;             /* 1/8!  */
    pc = pc * z - 1.0 / 720.0;               /* 1/6!  */
    pc = pc * z + 1.0 / 24.0;                /* 1/4!  */
    pc = pc * z - 1.0 / 2.0;                 /* 1/2!  */
    pc = pc * z + 1.0;
    double cr = pc;
    double so, co;
    switch (q & 3) {
    case 0:  so = sr;  co = cr;  break;
    case 1:  so = cr;  co = -sr; break;
    case 2:  so = -sr; co = -cr; break;
    default: so = -cr; co = sr;  break;
    }
    *sn = (float)so;
    *cs = (float)co;
}

static inline double dm_atan_d(double t)
{
    int neg = t < 0.0;
    if (neg) t = -t;
    double base = 0.0;
    int inv = 0;
    if (t > 1.0) { t = 1.0 / t; inv = 1; }
    if (t > DM_TANPIO8) { base = DM_PIO4; t = (t - 1.0) / (t + 1.0); }
    double z = t * t;
    double p = 1.0 / 27.0;
    p = -p * z + 1.0 / 25.0;
    p = -p * z + 1.0 / 23.0;
    p = -p * z + 1.0 / 21.0;
    p = -p * z + 1.0 / 19.0;
    p = -p * z + 1.0 / 17.0;
    p = -p * z + 1.0 / 15.0;
    p = -p * z + 1.0 / 13.0;
    p = -p * z + 1.0 / 11.0;
    p = -p * z + 1.0 / 9.0;
    p = -p * z + 1.0 / 7.0;
    p = -p * z + 1.0 / 5.0;
    p = -p * z + 1.0 / 3.0;
    p = -p * z + 1.0;
    double a = base + t * p;
    if (inv) a = DM_PIO2 - a;
    return neg ? -a : a;
}

static inline float dm_atanf(float t) { return (float)dm_atan_d((double)t); }

static inline float dm_acosf(float xf)
{
    double x = (double)xf;
    if (x >= 1.0) return 0.0f;
    if (x <= -1.0) return (float)DM_PI;
    double a = 2.0 * dm_atan_d(sqrt((1.0 - x) / (1.0 + x)));
    return (float)a;
}

/* exp: x = k*ln2 + r, |r| <= ln2/2, k = nearest integer to x/ln2; exp(r) by its Taylor
 * series to r^13 (Horner); 2^k applied by exponent arithmetic.  x < -708 -> 0, x > 709 -> inf. */
static inline double dm_exp_d(double x)
{
    if (x < -708.0) return 0.0;
    if (x > 709.0) return INFINITY;
    double t = x * 1.4426950408889634074;             /* 1/ln2 */
    int k = (int)(t >= 0.0 ? t + 0.5 : t - 0.5);
    double r = x - (double)k * DM_LN2;
    double p = 1.0 / 6227020800.0;                    /* 1/13! */
    p = p * r + 1.0 / 479001600.0;
    p = p * r + 1.0 / 39916800.0;
    p = p * r + 1.0 / 3628800.0;
    p = p * r + 1.0 / 362880.0;
    p = p * r + 1.0 / 40320.0;
    p = p * r + 1.0 / 5040.0;
    p = p * r + 1.0 / 720.0;
    p = p * r + 1.0 / 120.0;
    p = p * r + 1.0 / 24.0;
    p = p * r + 1.0 / 6.0;
    p = p * r + 0.5;
    p = p * r + 1.0;
    p = p * r + 1.0;
    /* scale by 2^k in two exact steps (k in [-1022, 1024)) */
    int k1 = k / 2, k2 = k - k1;
    uint64_t b1 = (uint64_t)(k1 + 1023) << 52, b2 = (uint64_t)(k2 + 1023) << 52;
    double s1, s2;
    memcpy(&s1, &b1, 8);
    memcpy(&s2, &b2, 8);
    return p * s1 * s2;
}

static inline float dm_expf(float x) { return (float)dm_exp_d((double)x); }

#endif /* ORACLE_DETMATH_H */
