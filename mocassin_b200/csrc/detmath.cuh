// detmath.cuh -- deterministic float32 transcendentals for the transport kernels.
//
// The reference calls the Fortran intrinsics log/cos/sin/acos/atan on default REAL
// (photon_mod.f90:1184,374,393; vector_mod.f90:303-314), whose last bits depend on
// the libm in use.  To make the device results reproducible bit for bit against the
// CPU oracle the kernels evaluate a fully specified algorithm built only from IEEE
// correctly-rounded double operations (+ - * / sqrt, conversions) and round once to
// float.  Must be compiled with -fmad=false (no FMA contraction).
//
// Specification:
//   log   : x = m*2^e, m in [sqrt(1/2), sqrt(2)); s=(m-1)/(m+1);
//           log x = e*ln2 + 2*(s + s^3/3 + ... + s^15/15)   (Horner in s^2)
//   sincos: q = round-half-away(x*2/pi); r = x - q*pi/2; Taylor sin to r^15, cos to
//           r^16 (Horner in r^2); quadrant q mod 4
//   atan  : |t|>1 -> 1/t;  t>tan(pi/8) -> (t-1)/(t+1);  odd Taylor series to u^27
//   acos  : 2*atan(sqrt((1-x)/(1+x))), acos(x<=-1)=pi, acos(x>=1)=0
//   exp   : x = k*ln2 + r; Taylor to r^13; scale by 2^k
#pragma once
#include <cstdint>

namespace mcb {

#define MCB_PI      3.14159265358979323846
#define MCB_PIO2    1.57079632679489661923
#define MCB_PIO4    0.78539816339744830962
#define MCB_2OPI    0.63661977236758134308
#define MCB_LN2     0.69314718055994530942
#define MCB_TANPIO8 0.41421356237309504880

__device__ __forceinline__ float dm_logf(float xf)
{
    uint32_t ix = __float_as_uint(xf);
    int e = (int)(ix >> 23) - 127;
    double m = (double)__uint_as_float((ix & 0x007fffffu) | 0x3f800000u);
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
    return (float)((double)e * MCB_LN2 + 2.0 * (s * p));
}

__device__ __forceinline__ void dm_sincosf(float xf, float &sn, float &cs)
{
    double x = (double)xf;
    double t = x * MCB_2OPI;
    int q = (int)(t >= 0.0 ? t + 0.5 : t - 0.5);
    double r = x - (double)q * MCB_PIO2;
    double z = r * r;
    double ps = -1.0 / 1307674368000.0;
    ps = ps * z + 1.0 / 6227020800.0;
    ps = ps * z - 1.0 / 39916800.0;
    ps = ps * z + 1.0 / 362880.0;
    ps = ps * z - 1.0 / 5040.0;
    ps = ps * z + 1.0 / 120.0;
    ps = ps * z - 1.0 / 6.0;
    ps = ps * z + 1.0;
    double sr = r * ps;
    double pc = 1.0 / 20922789888000.0;
    pc = pc * z - 1.0 / 87178291200.0;
    pc = pc * z + 1.0 / 479001600.0;
    pc = pc * z - 1.0 / 3628800.0;
    pc = pc * z + 1.0 / 40320.0;
    pc = pc * z - 1.0 / 720.0;
    pc = pc * z + 1.0 / 24.0;
    pc = pc * z - 1.0 / 2.0;
    pc = pc * z + 1.0;
    double cr = pc;
    double so, co;
    switch (q & 3) {
    case 0:  so = sr;  co = cr;  break;
    case 1:  so = cr;  co = -sr; break;
    case 2:  so = -sr; co = -cr; break;
    default: so = -cr; co = sr;  break;
    }
    sn = (float)so;
    cs = (float)co;
}

__device__ __forceinline__ double dm_atan_d(double t)
{
    bool neg = t < 0.0;
    if (neg) t = -t;
    double base = 0.0;
    bool inv = false;
    if (t > 1.0) { t = 1.0 / t; inv = true; }
    if (t > MCB_TANPIO8) { base = MCB_PIO4; t = (t - 1.0) / (t + 1.0); }
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
    if (inv) a = MCB_PIO2 - a;
    return neg ? -a : a;
}

__device__ __forceinline__ float dm_atanf(float t) { return (float)dm_atan_d((double)t); }

__device__ __forceinline__ float dm_acosf(float xf)
{
    double x = (double)xf;
    if (x >= 1.0) return 0.0f;
    if (x <= -1.0) return (float)MCB_PI;
    return (float)(2.0 * dm_atan_d(sqrt((1.0 - x) / (1.0 + x))));
}

// exp: x = k*ln2 + r, |r| <= ln2/2; Taylor series of exp(r) to r^13; 2^k by exponent
// arithmetic in two exact steps.  x < -708 -> 0, x > 709 -> inf.
__device__ __forceinline__ double dm_exp_d(double x)
{
    if (x < -708.0) return 0.0;
    if (x > 709.0) return __longlong_as_double(0x7ff0000000000000ll);
    double t = x * 1.4426950408889634074;
    int k = (int)(t >= 0.0 ? t + 0.5 : t - 0.5);
    double r = x - (double)k * MCB_LN2;
    double p = 1.0 / 6227020800.0;
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
    int k1 = k / 2, k2 = k - k1;
    double s1 = __longlong_as_double((long long)(k1 + 1023) << 52);
    double s2 = __longlong_as_double((long long)(k2 + 1023) << 52);
    return p * s1 * s2;
}

__device__ __forceinline__ float dm_expf(float x) { return (float)dm_exp_d((double)x); }

}  // namespace mcb
