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
static __device__ __forceinline__ float fp_rcp_approx(float x) {  // one MUFU.RCP (operand in [1.7, 2.5]); two double-precision refinements follow
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
#define FP_RCP32(x) fp_rcp_approx(x)
#else
#define FP_RCP32(x) (1.0f / (x))
#endif

// Polynomial coefficients live in a table: in the kernels a __constant__ array, so that DFMA takes them as
// constant-bank operands (as literals the compiler materialised every 64-bit value with two UMOVs per use:
// ~20 extra instructions per element in an issue-bound kernel).
#define FP_CONSTS                                                                                                     \
    { 1.0 / 21.0, 1.0 / 19.0, 1.0 / 17.0, 1.0 / 15.0, 1.0 / 13.0, 1.0 / 11.0, 1.0 / 9.0, 1.0 / 7.0, 1.0 / 5.0, 1.0 / 3.0, \
      1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0,          \
      1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5,                                                \
      0.6931471805599453, 2.3190468138462996e-17, 1.4426950408889634, 2.0355273740931033e-17, 6755399441055744.0 }
enum { FP_LN2_HI = 22, FP_LN2_LO = 23, FP_INVLN2_HI = 24, FP_INVLN2_LO = 25, FP_MAGIC = 26 };
#if defined(__CUDACC__)
static __constant__ double fp_dev_consts[27] = FP_CONSTS;
#endif
static const double fp_host_consts[27] = FP_CONSTS;
#if defined(__CUDA_ARCH__)
#define FP_K(i) fp_dev_consts[i]
#else
#define FP_K(i) fp_host_consts[i]
#endif

#if defined(__CUDA_ARCH__)
#define FP_HI(d) ((uint32_t)__double2hiint(d))
#define FP_LO(d) ((uint32_t)__double2loint(d))
#define FP_MAKE(hi, lo) __hiloint2double((int)(hi), (int)(lo))
#else
static inline uint32_t fp_hi_(double d) { uint64_t u; memcpy(&u, &d, 8); return (uint32_t)(u >> 32); }
static inline uint32_t fp_lo_(double d) { uint64_t u; memcpy(&u, &d, 8); return (uint32_t)u; }
static inline double fp_make_(uint32_t hi, uint32_t lo) { uint64_t u = ((uint64_t)hi << 32) | lo; double d; memcpy(&d, &u, 8); return d; }
#define FP_HI(d) fp_hi_(d)
#define FP_LO(d) fp_lo_(d)
#define FP_MAKE(hi, lo) fp_make_(hi, lo)
#endif

// Everything below is STRAIGHT-LINE code: out-of-domain operands produce garbage that the `ok` flag discards. The first
// version returned early per case; in the kernels that cost ~120 non-FP64 instructions per element (ncu: pow 166
// instructions per element, 42 of them FP64, issue-bound) because every element re-materialised the 64-bit constants
// and nothing could be interleaved across the elements of a thread.

// x = 2^e * m, m in [sqrt(1/2), sqrt(2)), from the bits of a positive normal FLOAT; returns log(m) (rel. error a few 2^-53).
FP_HD double fp_log_mantissa(uint32_t xbits, int* e_out) {
    int e = (int)(xbits >> 23) - 127;
    const uint32_t mant = xbits & 0x007fffffu;
    const uint32_t up = mant > 0x003504f3u;  // mantissa of sqrt(2)
    e += (int)up;
    const uint32_t mb = mant | (up ? 0x3f000000u : 0x3f800000u);
    float mf; memcpy(&mf, &mb, 4);
    *e_out = e;
    const double m = (double)mf;
    const double num = m - 1.0, den = m + 1.0;  // both exact (m has 24 significant bits)
    const double q0 = (double)FP_RCP32(mf + 1.0f);
    const double s0 = num * q0;
    const double s1 = fma(fma(-s0, den, num), q0, s0);
    const double s = fma(fma(-s1, den, num), q0, s1);  // second refinement: rel err of s ~2^-53
    const double z = s * s;
    // log(m) = 2 s (1 + z/3 + z^2/5 + ... + z^10/21)
    double p = FP_K(0);
    p = fma(p, z, FP_K(1));
    p = fma(p, z, FP_K(2));
    p = fma(p, z, FP_K(3));
    p = fma(p, z, FP_K(4));
    p = fma(p, z, FP_K(5));
    p = fma(p, z, FP_K(6));
    p = fma(p, z, FP_K(7));
    p = fma(p, z, FP_K(8));
    p = fma(p, z, FP_K(9));
    const double two_s = s + s;
    return fma(two_s * z, p, two_s);
}

// |distance| of a double from the nearest float rounding boundary (a midpoint between two floats), in units of the
// double's last place; only the low word is involved (29 bits are dropped by the conversion)
FP_HD uint32_t fp_boundary_distance(double r) {
    const uint32_t low = FP_LO(r) & 0x1fffffffu;
    return low > 0x10000000u ? low - 0x10000000u : 0x10000000u - low;
}

// (float)log((double)x) when that rounding is certain (same acceptance rule as fp_fast_pow), else 0.
FP_HD int fp_fast_log(float xf, float* out) {
    uint32_t xb; memcpy(&xb, &xf, 4);
    int e;
    const double lm = fp_log_mantissa(xb, &e);
    const double ed = (double)e;
    const double r = fma(ed, FP_K(FP_LN2_HI), fma(ed, FP_K(FP_LN2_LO), lm));
    const int normal = xb >= 0x00800000u && xb < 0x7f800000u;                    // positive, normal, finite
    const int main_ok = normal && fp_boundary_distance(r) >= (1u << 15) && fabs(r) > 1.2e-38;
    // log(1) = 0 exactly: r is +0 there, which the `fabs(r) > 1.2e-38` test would send to the slow path
    *out = (float)r;
    return main_ok || xb == 0x3f800000u;   // zero, negative, inf, NaN and subnormal operands: slow path
}

// 2^t, relative error ~2^-50; garbage for |t| >= 1000 (the caller discards it)
FP_HD double fp_exp2(double t) {
    const double magic = FP_K(FP_MAGIC);  // 1.5 * 2^52: adding it rounds t to an integer in the low word
    const double kd = t + magic;
    const uint32_t k = FP_LO(kd);
    const double f = t - (kd - magic);          // exact, |f| <= 0.5
    const double g = f * FP_K(FP_LN2_HI);       // f ln2 (|g| <= 0.3466)
    const double g2 = f * FP_K(FP_LN2_LO);      // low part of ln2 times f
    double p = FP_K(10);                        // 1/13!, 1/12!, ..., 1/2!
    p = fma(p, g, FP_K(11));
    p = fma(p, g, FP_K(12));
    p = fma(p, g, FP_K(13));
    p = fma(p, g, FP_K(14));
    p = fma(p, g, FP_K(15));
    p = fma(p, g, FP_K(16));
    p = fma(p, g, FP_K(17));
    p = fma(p, g, FP_K(18));
    p = fma(p, g, FP_K(19));
    p = fma(p, g, FP_K(20));
    p = fma(p, g, FP_K(21));
    p = fma(p, g, 1.0);
    double r = fma(p, g, 1.0);
    r = fma(r, g2, r);
    return FP_MAKE(FP_HI(r) + (k << 20), FP_LO(r));  // times 2^k through the exponent field
}

// Fast pow for float operands. Returns 1 and *out when the float rounding of the result is certain,
// 0 when the caller must take the slow path (special operands, range, or a result too close to a rounding boundary).
FP_HD int fp_fast_pow(float xf, float yf, float* out) {
    uint32_t xb; memcpy(&xb, &xf, 4);
    uint32_t yb; memcpy(&yb, &yf, 4);
    const uint32_t xa = xb & 0x7fffffffu, ya = yb & 0x7fffffffu;
    const int dom = xa >= 0x00800000u && xa < 0x7f800000u && ya < 0x7f800000u;  // x normal finite non-zero, y finite
    const int neg = (int)(xb >> 31);
    const float ay = fabsf(yf);
    const int y_int = ay == truncf(ay);
    const int odd = ay < 16777216.0f && (((int)ay) & 1);   // (floats >= 2^24 are even integers)
    const int negate = neg && odd;
    int e;
    const double lm = fp_log_mantissa(xa, &e);
    const double L = (double)e + fma(lm, FP_K(FP_INVLN2_LO), lm * FP_K(FP_INVLN2_HI));  // 1/ln2 = hi + lo                    // log2|x|, abs error <= ~2^-52 max(0.5, |L|)
    const double y = (double)yf;
    const double t = y * L;
    const double r = fp_exp2(t);
    const double at = fabs(t);
    // the main path: result a normal float away from overflow / underflow, and further than 2^-37 (relative) from a
    // float rounding boundary (error of r: <= 2^-45 measured)
    int ok = dom && (!neg || y_int) && at < 125.0 && fp_boundary_distance(r) >= (1u << 15);
    float res = (float)(negate ? -r : r);
    // far beyond the float range the rounded result is certain: +-inf, or +-0 below half the smallest subnormal
    if (dom && (!neg || y_int) && t > 129.0) { res = negate ? -INFINITY : INFINITY; ok = 1; }
    if (dom && (!neg || y_int) && t < -151.0) { res = negate ? -0.0f : 0.0f; ok = 1; }
    if (dom && neg && !y_int) { res = NAN; ok = 1; }   // domain error (the caller gives the NaN the reference's sign, elementwise.cuh)
    if (yf == 2.0f && xf == xf) {                      // x*x is exact in double, so its float rounding is the correct one
        const double xd = (double)xf;
        res = (float)(xd * xd);
        ok = 1;
    }
    if (ya == 0) { res = 1.0f; ok = 1; }               // pow(anything, +-0) = 1, NaN included (C99 F.9.4.4; glibc agrees)
    *out = res;
    return ok;
}
#endif
