// pt_detmath.h -- portable, bit-reproducible transcendentals for the shading path.
//
// The reference's device programs call sinf/cosf/acosf (pathTracerPrograms.cu:347-351,
// 375-377,459-466) and powf (cuda/helpers.h:35-42).  glibc and CUDA libdevice round these
// differently in the last place, which would make a CPU-vs-GPU comparison of a stochastic
// path tracer fuzzy (one flipped Russian-roulette decision changes a whole path).  These
// replacements use only IEEE add / mul / fma / div / sqrt / round-to-int, all of which are
// correctly rounded on both x86-64 (compiled with -ffp-contract=off) and sm_100a (compiled
// with -fmad=false), so host and device produce identical bits.  Accuracy is ~1-2 ulp over
// the argument ranges the path produces (sin/cos: |x| <= ~1e3; acos: [-1,1]; pow: x in [0,1]).
//
// tests/test_detmath.py bounds the deviation from glibc.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define PT_HD __host__ __device__ __forceinline__
#else
#define PT_HD static inline
#endif

namespace ptdet {

PT_HD float fma_(float a, float b, float c)
{
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}

PT_HD double dfma_(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}

// ---- sin / cos ------------------------------------------------------------------------------
// Cody-Waite reduction by pi/2 split into three floats, then the classic single-precision
// minimax kernels on [-pi/4, pi/4].
PT_HD void sincos_reduce(float x, float& r, int& q)
{
    const float two_over_pi = 0.636619772367581343f;
    const float pio2_hi = 1.5707962512969971e+00f;   // 0x3fc90fda
    const float pio2_mid = 7.5497894158615964e-08f;  // 0x33a22168
    const float pio2_lo = 5.3903025299577648e-15f;   // 0x27c234c4
    float k = rintf(x * two_over_pi);
    r = fma_(-k, pio2_hi, x);
    r = fma_(-k, pio2_mid, r);
    r = fma_(-k, pio2_lo, r);
    q = (int)k;
}

PT_HD float sin_kernel(float r)
{
    float z = r * r;
    float p = -1.9515295891e-4f;
    p = fma_(p, z, 8.3321608736e-3f);
    p = fma_(p, z, -1.6666654611e-1f);
    return fma_(p * z, r, r);
}

PT_HD float cos_kernel(float r)
{
    float z = r * r;
    float p = 2.443315711809948e-5f;
    p = fma_(p, z, -1.388731625493765e-3f);
    p = fma_(p, z, 4.166664568298827e-2f);
    float zz = z * z;
    return fma_(p, zz, fma_(-0.5f, z, 1.0f));
}

PT_HD void sincosf_(float x, float& s, float& c)
{
    float r;
    int q;
    sincos_reduce(x, r, q);
    float sk = sin_kernel(r);
    float ck = cos_kernel(r);
    float ss = (q & 1) ? ck : sk;
    float cc = (q & 1) ? sk : ck;
    s = (q & 2) ? -ss : ss;
    c = ((q + 1) & 2) ? -cc : cc;
}

PT_HD float sinf_(float x)
{
    float s, c;
    sincosf_(x, s, c);
    return s;
}

PT_HD float cosf_(float x)
{
    float s, c;
    sincosf_(x, s, c);
    return c;
}

// ---- acos -----------------------------------------------------------------------------------
// asin kernel on [0, 0.5]; acos(x) = pi/2 - asin(x) for |x| <= 0.5, 2*asin(sqrt((1-|x|)/2))
// otherwise (reflected for x < 0).
PT_HD float asin_kernel(float a)
{
    float z = a * a;
    float p = 4.2163199048e-2f;
    p = fma_(p, z, 2.4181311049e-2f);
    p = fma_(p, z, 4.5470025998e-2f);
    p = fma_(p, z, 7.4953002686e-2f);
    p = fma_(p, z, 1.6666752422e-1f);
    return fma_(p * z, a, a);
}

PT_HD float acosf_(float x)
{
    const float pio2_hi = 1.5707962512969971e+00f;
    const float pio2_lo = 7.5497894158615964e-08f;
    float a = fabsf(x);
    if (a > 1.0f) return NAN;
    if (a <= 0.5f) {
        float as = asin_kernel(x);
        // pi/2 - asin(x), with the low word of pi/2 folded in
        return (pio2_hi - as) + pio2_lo;
    }
    float h = (1.0f - a) * 0.5f;
    float s = sqrtf(h);
    float r = 2.0f * asin_kernel(s);
    if (x < 0.0f) {
        const float pi_hi = 3.1415925025939941e+00f;
        const float pi_lo = 1.5099578831723193e-07f;
        return (pi_hi - r) + pi_lo;
    }
    return r;
}

// ---- pow(x, y) for x in (0, +inf), used with x in [0.0031308, 1] and y = 1/2.4 ----------------
// Evaluated in double (log via atanh series on [sqrt(1/2), sqrt(2)), exp via Taylor on
// [-ln2/2, ln2/2]); the double result is accurate to ~1e-14 relative, so the final rounding to
// float matches a correctly-rounded powf except in astronomically rare near-ties.
PT_HD double log_d(double x)
{
    uint64_t bits;
#if defined(__CUDA_ARCH__)
    bits = (uint64_t)__double_as_longlong(x);
#else
    memcpy(&bits, &x, 8);
#endif
    int e = (int)((bits >> 52) & 0x7ff) - 1023;
    bits = (bits & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
    double m;
#if defined(__CUDA_ARCH__)
    m = __longlong_as_double((long long)bits);
#else
    memcpy(&m, &bits, 8);
#endif
    if (m > 1.4142135623730951) {
        m *= 0.5;
        e += 1;
    }
    double s = (m - 1.0) / (m + 1.0);
    double z = s * s;
    double p = 1.0 / 19.0;
    p = dfma_(p, z, 1.0 / 17.0);
    p = dfma_(p, z, 1.0 / 15.0);
    p = dfma_(p, z, 1.0 / 13.0);
    p = dfma_(p, z, 1.0 / 11.0);
    p = dfma_(p, z, 1.0 / 9.0);
    p = dfma_(p, z, 1.0 / 7.0);
    p = dfma_(p, z, 1.0 / 5.0);
    p = dfma_(p, z, 1.0 / 3.0);
    p = dfma_(p, z, 1.0);
    double lm = 2.0 * s * p;
    return dfma_((double)e, 0.6931471805599453, lm);
}

PT_HD double exp_d(double x)
{
    // valid for |x| < ~700
    double k = rint(x * 1.4426950408889634);
    double r = dfma_(-k, 6.93147180369123816490e-01, x);
    r = dfma_(-k, 1.90821492927058770002e-10, r);
    double p = 1.0 / 6227020800.0;  // 1/13!
    p = dfma_(p, r, 1.0 / 479001600.0);
    p = dfma_(p, r, 1.0 / 39916800.0);
    p = dfma_(p, r, 1.0 / 3628800.0);
    p = dfma_(p, r, 1.0 / 362880.0);
    p = dfma_(p, r, 1.0 / 40320.0);
    p = dfma_(p, r, 1.0 / 5040.0);
    p = dfma_(p, r, 1.0 / 720.0);
    p = dfma_(p, r, 1.0 / 120.0);
    p = dfma_(p, r, 1.0 / 24.0);
    p = dfma_(p, r, 1.0 / 6.0);
    p = dfma_(p, r, 0.5);
    p = dfma_(p, r, 1.0);
    p = dfma_(p, r, 1.0);
    int ki = (int)k;
    uint64_t sb = (uint64_t)(ki + 1023) << 52;
    double sc;
#if defined(__CUDA_ARCH__)
    sc = __longlong_as_double((long long)sb);
#else
    memcpy(&sc, &sb, 8);
#endif
    return p * sc;
}

PT_HD float powf_(float x, float y)
{
    if (x <= 0.0f) return (y > 0.0f) ? 0.0f : ((y == 0.0f) ? 1.0f : INFINITY);
    double l = log_d((double)x);
    return (float)exp_d((double)y * l);
}

}  // namespace ptdet
