// pow() for float32 operands with the reference's result, at a fraction of the cost of a full double-precision pow.
//
// The reference computes (float)pow((double)x, (double)y) with glibc (arraylib.c:17). The device used to run CUDA's
// double pow for every element (~150 FP64 instructions: 1.57 TB/s, 24 % of the HBM roofline). Only the FLOAT
// rounding of the result matters, so: compute 2^(y log2 x) in double with ~40 FP64 operations and a relative error
// of <= 2^-45 (measured; log2 through s = (m-1)/(m+1) refined twice from a float reciprocal, exp2 by a degree-13
// polynomial), and accept the result only when it lies further than 2^-37 (relative) from the nearest float rounding
// boundary -- then every value within the error bound, the exact one and glibc's included, rounds to the same float.
// The remaining ~0.012 % of the elements, special operands and results outside the normal float range take the slow
// path (the caller's full pow). y == 2 (exact in double), y == 0 and far over/underflow are answered directly.
//
// Plain C so that the SAME code is validated on the CPU against glibc (tests/csrc/fast_pow_harness.c, run by
// tests/test_oracle.py: 0 mismatches in 2.5e9 accepted results during development) and compiled into the kernels.
#ifndef ARRAYLIB_B200_FAST_POW_H
#define ARRAYLIB_B200_FAST_POW_H
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define FP_HD __host__ __device__ __forceinline__
#else
#define FP_HD static inline
#endif
#if defined(__CUDA_ARCH__)
#define FP_RCP32(x) __frcp_rn(x)
#else
#define FP_RCP32(x) (1.0f / (x))
#endif

FP_HD uint64_t fp_bits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
FP_HD double fp_from_bits(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }


// Natural log of the mantissa of a positive normal double that came from a float: x = 2^e * m with m in
// [sqrt(1/2), sqrt(2)); returns log(m) with a relative error of a few 2^-53 and stores e.
FP_HD double fp_log_mantissa(double x, int* e_out) {
    uint64_t ix = fp_bits(x);
    int e = (int)(ix >> 52) - 1023;
    uint64_t mant = ix & 0x000fffffffffffffull;
    if (mant > 0x6a09e667f3bcdull) { e += 1; ix = mant | 0x3fe0000000000000ull; }
    else ix = mant | 0x3ff0000000000000ull;
    *e_out = e;
    const double m = fp_from_bits(ix);
    const double num = m - 1.0, den = m + 1.0;  // both exact (m has <= 24 significant bits here)
    const double q0 = (double)FP_RCP32((float)den);
    const double s0 = num * q0;
    const double s1 = fma(fma(-s0, den, num), q0, s0);
    const double s = fma(fma(-s1, den, num), q0, s1);  // second refinement: rel err of s ~2^-53
    const double z = s * s;
    // log(m) = 2 s (1 + z/3 + z^2/5 + ... + z^10/21)
    double p = 1.0 / 21.0;
    p = fma(p, z, 1.0 / 19.0);
    p = fma(p, z, 1.0 / 17.0);
    p = fma(p, z, 1.0 / 15.0);
    p = fma(p, z, 1.0 / 13.0);
    p = fma(p, z, 1.0 / 11.0);
    p = fma(p, z, 1.0 / 9.0);
    p = fma(p, z, 1.0 / 7.0);
    p = fma(p, z, 1.0 / 5.0);
    p = fma(p, z, 1.0 / 3.0);
    const double two_s = s + s;
    return fma(two_s * z, p, two_s);
}

// log2 of a positive normal double that came from a float. Abs error <= ~2^-52 * max(0.5, |L|).
FP_HD double fp_log2(double x) {
    int e;
    const double lm = fp_log_mantissa(x, &e);
    const double inv_hi = 1.4426950408889634, inv_lo = 2.0355273740931033e-17;  // 1/ln2 = hi + lo
    return (double)e + fma(lm, inv_lo, lm * inv_hi);
}

// (float)log((double)x) when that rounding is certain (same acceptance rule as fp_fast_pow), else 0.
FP_HD int fp_fast_log(float xf, float* out) {
    uint32_t xb; memcpy(&xb, &xf, 4);
    if (xb == 0x3f800000u) { *out = 0.0f; return 1; }
    if (xb < 0x00800000u || xb >= 0x7f800000u) {
        if ((xb & 0x7fffffffu) == 0) { *out = -INFINITY; return 1; }                // log(+-0) = -inf
        if (xb == 0x7f800000u) { *out = INFINITY; return 1; }
        if ((xb >> 31) && (xb & 0x7fffffffu) <= 0x7f800000u) { *out = NAN; return 1; }  // negative (incl. -inf): domain error
        return 0;  // subnormal or NaN: slow path
    }
    int e;
    const double lm = fp_log_mantissa((double)xf, &e);
    const double ln2_hi = 0.6931471805599453, ln2_lo = 2.3190468138462996e-17;
    const double r = fma((double)e, ln2_hi, fma((double)e, ln2_lo, lm));
    const uint32_t low = (uint32_t)(fp_bits(r) & 0x1fffffffu);
    const uint32_t d = low > 0x10000000u ? low - 0x10000000u : 0x10000000u - low;
    if (d < (1u << 15)) return 0;
    const double ar = fabs(r);
    if (!(ar > 1.2e-38)) return 0;
    *out = (float)r;
    return 1;
}

// 2^t for |t| < 1000, relative error ~2^-50
FP_HD double fp_exp2(double t) {
    const double k = rint(t);
    const double f = t - k;  // exact, |f| <= 0.5
    const double g = f * 0.6931471805599453;  // f ln2 (rel err 2^-53, |g| <= 0.3466)
    const double g2 = fma(f, 2.3190468138462996e-17, 0.0);  // low part of ln2 times f
    double p = 1.0 / 6227020800.0;            // 1/13!
    p = fma(p, g, 1.0 / 479001600.0);
    p = fma(p, g, 1.0 / 39916800.0);
    p = fma(p, g, 1.0 / 3628800.0);
    p = fma(p, g, 1.0 / 362880.0);
    p = fma(p, g, 1.0 / 40320.0);
    p = fma(p, g, 1.0 / 5040.0);
    p = fma(p, g, 1.0 / 720.0);
    p = fma(p, g, 1.0 / 120.0);
    p = fma(p, g, 1.0 / 24.0);
    p = fma(p, g, 1.0 / 6.0);
    p = fma(p, g, 0.5);
    p = fma(p, g, 1.0);
    double r = fma(p, g, 1.0);
    r = fma(r, g2, r);
    // scale by 2^k through the exponent field (result stays normal for |t| < 1000)
    return fp_from_bits(fp_bits(r) + ((uint64_t)(int64_t)k << 52));
}

// Fast pow for float operands. Returns 1 and *out when the float rounding of the result is certain,
// 0 when the caller must take the slow path (special operands, range, or a result too close to a rounding boundary).
FP_HD int fp_fast_pow(float xf, float yf, float* out) {
    uint32_t xb; memcpy(&xb, &xf, 4);
    uint32_t yb; memcpy(&yb, &yf, 4);
    const uint32_t xa = xb & 0x7fffffffu, ya = yb & 0x7fffffffu;
    // x: normal, finite, non-zero; y: finite, non-zero, |y| >= 2^-60 is not needed but keep NaN/inf/0 out
    if (ya == 0) { *out = 1.0f; return 1; }  // pow(anything, +-0) = 1, NaN included (C99 F.9.4.4; glibc agrees)
    if (xa < 0x00800000u || xa >= 0x7f800000u || ya >= 0x7f800000u) return 0;
    int negate = 0;
    if (xb >> 31) {  // negative base: only integer exponents have a real result
        const float ay = fabsf(yf);
        if (ay != truncf(ay)) {  // domain error: a NaN (the caller gives it the reference's sign, elementwise.cuh)
            *out = NAN;
            return 1;
        }
        negate = ay < 16777216.0f && (((int)ay) & 1);
    }
    const double x = (double)fabsf(xf), y = (double)yf;
    double r;
    if (yf == 2.0f) {
        r = x * x;  // exact in double
        if (r > 3.0e38 || r < 1.2e-38) return 0;
        *out = (float)(negate ? -r : r);
        return 1;
    }
    const double t = y * fp_log2(x);
    if (!(fabs(t) < 125.0)) {
        // far beyond the float range the rounded result is certain: +-inf, or +-0 below the smallest subnormal's half
        if (t > 129.0) { *out = negate ? -INFINITY : INFINITY; return 1; }
        if (t < -151.0) { *out = negate ? -0.0f : 0.0f; return 1; }
        return 0;  // overflow threshold / subnormal results: slow path
    }
    r = fp_exp2(t);
    // distance of the double from the nearest float rounding boundary (midpoint), in units of 2^-52 relative
    const uint32_t low = (uint32_t)(fp_bits(r) & 0x1fffffffu);
    const uint32_t d = low > 0x10000000u ? low - 0x10000000u : 0x10000000u - low;
    if (d < (1u << 15)) return 0;  // within 2^-37 relative of a midpoint: let the slow path decide
    *out = (float)(negate ? -r : r);
    return 1;
}
#endif
