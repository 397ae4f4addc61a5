// Certified fast paths for the reference's transcendental ufuncs on float32 operands.
//
// The reference computes (float)f((double)x) with glibc (arraylib.c:31-49). CUDA's double-precision routines carry
// 53 bits through ~60-250 instructions per element (1.1-4.5 TB/s at 2^28 elements). Only the FLOAT rounding of the
// result matters, so (same idea as fast_pow.h): evaluate f in double with a short polynomial kernel whose relative
// error is <= 2^-42 (Chebyshev fits from tools/gen_ufunc_coeffs.py; divisions and square roots by one residual
// correction of a float MUFU seed, <= 2^-43), and ACCEPT the result only when it lies further than 2^-37 (relative)
// from the nearest float rounding boundary. Then every value within the error bound -- the exact one, glibc's and
// ours -- rounds to the same float. The other ~0.012 % of the elements, special operands and arguments outside the
// stated ranges return 0 and the caller runs the full double-precision routine.
//
// Plain C: the SAME code is compiled into the kernels and into tests/csrc/fast_ufunc_harness.c, which compares every
// accepted result with glibc over ALL 2^32 float bit patterns (tests/test_oracle.py runs a strided subset). On the
// host the MUFU seeds are replaced by correctly rounded ones that the harness PERTURBS by +-2^-21 to cover the
// hardware's approximation error. (Replacing the F2F / I2F conversions by integer bit surgery was tried because ncu
// shows the conversion unit as the busiest pipe; it made every function 10-25 % slower: the kernels are issue-bound.)
#ifndef ARRAYLIB_B200_FAST_UFUNC_H
#define ARRAYLIB_B200_FAST_UFUNC_H
#include "fast_pow.h"

#if defined(__CUDA_ARCH__)
static __device__ __forceinline__ float fu_rsqrt_approx(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
#define FU_RCP32(x) fp_rcp_approx(x)
#define FU_RSQ32(x) fu_rsqrt_approx(x)
#elif defined(FU_PERTURB_SEEDS)
static inline float fu_perturb_(float v, float x) {  // deterministic +-2^-21 relative wobble, sign from the operand bits
    uint32_t b; memcpy(&b, &x, 4);
    b = (b ^ (b >> 7) ^ (b >> 15)) & 3u;
    const float w = b == 0 ? 1.0f + 4.8e-7f : b == 1 ? 1.0f - 4.8e-7f : b == 2 ? 1.0f + 2.4e-7f : 1.0f - 2.4e-7f;
    return v * w;
}
#define FU_RCP32(x) fu_perturb_(1.0f / (x), (x))
#define FU_RSQ32(x) fu_perturb_(1.0f / sqrtf(x), (x))
#else
#define FU_RCP32(x) (1.0f / (x))
#define FU_RSQ32(x) (1.0f / sqrtf(x))
#endif

// Coefficients, highest power first inside each group (constant-bank operands of DFMA in the kernels).
#define FU_CONSTS                                                                                                       \
    { /* 0: (exp(r) - 1)/r on |r| <= ln2/32, degree 4 */ 0.008333450014676098, 0.04166748343759084,                      \
      0.16666666665294327, 0.49999999990393607, 1.0,                                                                    \
      /* 5: cosh(sqrt z), degree 2, same range */ 0.04166764679280493, 0.49999999982708465, 1.0000000000000044,          \
      /* 8: sinh(sqrt z)/sqrt z, degree 2 */ 0.008333473351059013, 0.16666666664196456, 1.0000000000000007,              \
      /* 11: (atanh(s)/s - 1)/z, degree 4, z = s^2 <= 0.03 */ 0.09693104947071855, 0.11095088435332272,                  \
      0.14285886999303407, 0.19999999350549125, 0.3333333333372321,                                                     \
      /* 16: (sin(r)/r - 1)/z on |r| <= 1.579, degree 6 */ -7.405589436625779e-13, 1.6050804050730897e-10,               \
      -2.5051968654113158e-08, 2.7557318006278114e-06, -0.00019841269836213635, 0.008333333333325458,                   \
      -0.16666666666666646,                                                                                             \
      /* 23: (sin(r)/r - 1)/z on |r| <= 0.7895, degree 4 */ -2.4803067490801243e-08, 2.7555963052068327e-06,             \
      -0.00019841266824366998, 0.008333333330983484, -0.16666666666663738,                                              \
      /* 28: (cos(r) - 1)/z on |r| <= 0.7895, degree 5 */ 2.066330328792122e-09, -2.7555824793677795e-07,                \
      2.4801582475187907e-05, -0.0013888888881840808, 0.04166666666662902, -0.49999999999999967,                        \
      /* 34.. */ 0.04332169878499658 /* ln(2)/16 */, 0.6931471805599453 /* ln 2 */, 6755399441055744.0 /* 1.5 * 2^52 */, \
      0.3183098861837907 /* 1/pi */, 3.141592653589793 /* pi hi */, 1.2246467991473532e-16 /* pi lo */,                 \
      0.6366197723675814 /* 2/pi */, 1.5707963267948966 /* pi/2 hi */, 6.123233995736766e-17 /* pi/2 lo */,             \
      0.34657359027997264 /* ln(2) / 2 */,                                                                              \
      /* 44: (atan(t)/t - 1)/z on |t| <= 0.225, degree 5 */ 0.06759140727306162, -0.09039176935474555,                   \
      0.11109770226062675, -0.14285698480760142, 0.19999999931636747, -0.333333333332853,                               \
      0.4124104415973873 /* atan(7/16) */, 0.7853981633974483 /* atan(1) */, 23.083120654223414 /* 16 / ln 2 */ }
enum { FU_EX = 0, FU_CH = 5, FU_SH = 8, FU_LG = 11, FU_SN = 16, FU_S4 = 23, FU_C4 = 28, FU_LN2_16 = 34, FU_LN2 = 35, FU_MAGIC = 36,
       FU_INVPI = 37, FU_PI_HI = 38, FU_PI_LO = 39, FU_2OPI = 40, FU_PIO2_HI = 41, FU_PIO2_LO = 42, FU_LN2_HALF = 43,
       FU_AT = 44, FU_ATAN_C1 = 50, FU_ATAN_1 = 51, FU_16_LN2 = 52, FU_NCONST = 53 };
// 2^(j/16): read with per-lane indices, so it lives in global memory (one 128-byte line, L1-resident) instead of the
// constant bank, where divergent indices would serialise
#define FU_T16 { 1.0, 1.0442737824274138, 1.0905077326652577, 1.1387886347566916, 1.189207115002721, 1.241857812073484,  \
                 1.2968395546510096, 1.3542555469368927, 1.4142135623730951, 1.4768261459394993, 1.5422108254079407,      \
                 1.6104903319492543, 1.681792830507429, 1.7562521603732995, 1.8340080864093424, 1.9152065613971474 }
#if defined(__CUDACC__)
static __constant__ double fu_dev_consts[FU_NCONST] = FU_CONSTS;
static __device__ __align__(128) const double fu_dev_t16[16] = FU_T16;
#endif
static const double fu_host_consts[FU_NCONST] = FU_CONSTS;
static const double fu_host_t16[16] = FU_T16;
#if defined(__CUDA_ARCH__)
#define FU_K(i) fu_dev_consts[i]
#define FU_POW2_16TH(j) __ldg(&fu_dev_t16[j])
#else
#define FU_K(i) fu_host_consts[i]
#define FU_POW2_16TH(j) fu_host_t16[j]
#endif

#define FU_ACCEPT_DIST (1u << 15)  // 2^15 units of the double's last place = 2^-37 relative


// k = round(x * 16/ln 2) (magic-number rounding in double: DFMA + DADD, one issue slot fewer than the float variant with
// its conversion), r = x - k ln2/16 (|r| <= 0.02167): exp(x) = 2^(k >> 4) * 2^((k & 15)/16) * exp(r)
FP_HD int fu_exp_reduce(float x, double* R) {
    const double d = (double)x, magic = FU_K(FU_MAGIC);
    const double kd = fma(d, FU_K(FU_16_LN2), magic);
    *R = fma(-(kd - magic), FU_K(FU_LN2_16), d);
    return (int)FP_LO(kd);
}

// cosh(r) and sinh(r)/r on the reduced range
FP_HD void fu_cosh_sinh_small(double r, double* C, double* S) {
    const double z = r * r;
    double c = FU_K(FU_CH);
    c = fma(c, z, FU_K(FU_CH + 1));
    c = fma(c, z, FU_K(FU_CH + 2));
    double s = FU_K(FU_SH);
    s = fma(s, z, FU_K(FU_SH + 1));
    s = fma(s, z, FU_K(FU_SH + 2));
    *C = c;
    *S = s;
}

// 2^(k/16 - 1) for -2000 <= k <= 2000
FP_HD double fu_pow2_16th_half(int k) {
    const double T = FU_POW2_16TH((uint32_t)k & 15u);
    return FP_MAKE(FP_HI(T) + ((uint32_t)((k >> 4) - 1) << 20), FP_LO(T));
}

FP_HD double fu_pow2(int k) { return FP_MAKE((uint32_t)(1023 + k) << 20, 0u); }  // 2^k, -1022 <= k <= 1023

// num / den from a float reciprocal seed and one residual correction (relative error <= ~2^-43)
FP_HD double fu_div(double num, double den) {
    const double q0 = (double)FU_RCP32((float)den);
    const double t0 = num * q0;
    return fma(fma(-t0, den, num), q0, t0);
}

// sqrt(t) for a positive normal-float-range t, same scheme (relative error <= ~2^-43)
FP_HD double fu_sqrt(double t) {
    const float y0 = FU_RSQ32((float)t);
    const double yd = (double)y0;
    const double u0 = t * yd;
    return fma(fma(-u0, u0, t), (double)(0.5f * y0), u0);
}

// log(num/den) / 2 = s + s z P(z), s = num/den... given s = (w-1)/(w+1) for w in [1/sqrt2, sqrt2]: returns atanh(s)
FP_HD double fu_atanh_small(double s) {
    const double z = s * s;
    double p = FU_K(FU_LG);
    p = fma(p, z, FU_K(FU_LG + 1));
    p = fma(p, z, FU_K(FU_LG + 2));
    p = fma(p, z, FU_K(FU_LG + 3));
    p = fma(p, z, FU_K(FU_LG + 4));
    return fma(s * z, p, s);
}

// 2 v through the exponent field (v = +0 becomes 2^-1022, which still rounds to +0 as a float)
FP_HD double fu_twice(double v) { return FP_MAKE(FP_HI(v) + 0x00100000u, FP_LO(v)); }

// binary exponent j such that y * 2^-j lies in [1/sqrt2, sqrt2) (up to the float's rounding), y a positive normal float
FP_HD int fu_exponent_near_sqrt2(float y) {
    uint32_t b; memcpy(&b, &y, 4);
    return (int)(b >> 23) - 127 + (int)((b & 0x007fffffu) > 0x003504f3u);
}

// r a normal-float-range double: round, and say whether that rounding is certain, i.e. whether the 29 bits that the
// conversion drops are at least 2^15 away from the tie 2^28 (three integer instructions)
FP_HD int fu_finish(double r, float* out) {
    *out = (float)r;
    return ((FP_LO(r) - (0x10000000u - FU_ACCEPT_DIST + 1u)) & 0x1fffffffu) >= 2u * FU_ACCEPT_DIST - 1u;
}

FP_HD uint32_t fu_abs_bits(float x) { uint32_t b; memcpy(&b, &x, 4); return b & 0x7fffffffu; }

// ---- exp, sinh, cosh, tanh --------------------------------------------------------------------------------------------
FP_HD int fu_fast_exp(float x, float* out) {
    double r;
    const int k = fu_exp_reduce(x, &r);
    double q = FU_K(FU_EX);
    q = fma(q, r, FU_K(FU_EX + 1));
    q = fma(q, r, FU_K(FU_EX + 2));
    q = fma(q, r, FU_K(FU_EX + 3));
    q = fma(q, r, FU_K(FU_EX + 4));
    const double T = FU_POW2_16TH((uint32_t)k & 15u);
    const double v = fma(T, r * q, T);
    const double res = FP_MAKE(FP_HI(v) + ((uint32_t)(k >> 4) << 20), FP_LO(v));
    return fu_finish(res, out) & (fabsf(x) <= 87.0f);  // |x| <= 87: the result is a normal float
}

// sinh = (P - Q) C + (P + Q) r S, cosh = (P + Q) C + (P - Q) r S with P = 2^(k/16 - 1), Q = 2^(-k/16 - 1); k == 0
// gives P = Q = 1/2 and sinh = r S exactly, so there is no cancellation for small |x|
FP_HD int fu_fast_sinh(float x, float* out) {
    double r, C, S;
    const int k = fu_exp_reduce(x, &r);
    fu_cosh_sinh_small(r, &C, &S);
    const double P = fu_pow2_16th_half(k), Q = fu_pow2_16th_half(-k);
    const double res = fma(P - Q, C, (P + Q) * (r * S));
    const float ab = fabsf(x);
    const int ok = fu_finish(res, out) & (ab <= 87.0f);
    if (ab < 0x1p-14f) { *out = x; return 1; }  // |x| < 2^-14: x (1 + x^2/6) rounds to x
    return ok;
}

FP_HD int fu_fast_cosh(float x, float* out) {
    double r, C, S;
    const int k = fu_exp_reduce(x, &r);
    fu_cosh_sinh_small(r, &C, &S);
    const double P = fu_pow2_16th_half(k), Q = fu_pow2_16th_half(-k);
    const double res = fma(P + Q, C, (P - Q) * (r * S));
    return fu_finish(res, out) & (fabsf(x) <= 87.0f);
}

FP_HD int fu_fast_tanh(float x, float* out) {
    double r, C, S;
    const int k = fu_exp_reduce(x, &r);
    fu_cosh_sinh_small(r, &C, &S);
    const double P = fu_pow2_16th_half(k), Q = fu_pow2_16th_half(-k);
    const double amb = P - Q, apb = P + Q, rS = r * S;
    const double res = fu_div(fma(amb, C, apb * rS), fma(apb, C, amb * rS));
    const float ab = fabsf(x);
    const int ok = fu_finish(res, out) & (ab < 9.5f);  // |x| < 9.5
    if (ab < 0x1p-14f) { *out = x; return 1; }            // x (1 - x^2/3) rounds to x
    if (ab >= 9.5f) {                                                    // 1 - 2 exp(-2|x|) rounds to 1 from |x| = 9.02 on
        *out = x < 0.0f ? -1.0f : 1.0f;
        return 1;
    }
    return ok;
}

// ---- log, atanh, asinh, acosh ---------------------------------------------------------------------------------------
// log(v) for a positive normal double v
FP_HD double fu_log_core(double v) {
    const uint32_t hi = FP_HI(v);
    const uint32_t mant = hi & 0x000fffffu;
    const uint32_t up = mant > 0x0006a09eu;  // leading mantissa bits of sqrt(2)
    const int e = (int)(hi >> 20) - 1023 + (int)up;
    const double m = FP_MAKE(mant | (up ? 0x3fe00000u : 0x3ff00000u), FP_LO(v));
    const double s = fu_div(m - 1.0, m + 1.0);  // m - 1 exact; m + 1 rounded (2^-53)
    return fu_twice(fma((double)e, FU_K(FU_LN2_HALF), fu_atanh_small(s)));  // atanh(s) = log(m) / 2
}

FP_HD int fu_fast_log(float x, float* out) {
    const double r = fu_log_core((double)x);
    return fu_finish(r, out) & (x >= 0x1p-126f) & (x <= 3.402823466e38f);  // positive, normal, finite (log 1 = +0 passes)
}

FP_HD int fu_fast_atanh(float x, float* out) {
    const float ab = fabsf(x);
    const double d = (double)x;
    const double N = 1.0 + d, D = 1.0 - d;                    // exact for 2^-14 <= |x| < 1
    const int j = fu_exponent_near_sqrt2((1.0f + x) * FU_RCP32(1.0f - x));
    const double Dj = FP_MAKE(FP_HI(D) + ((uint32_t)j << 20), FP_LO(D));
    const double s = fu_div(N - Dj, N + Dj);
    const double r = fma((double)j, FU_K(FU_LN2_HALF), fu_atanh_small(s));
    const int ok = fu_finish(r, out) & (ab < 1.0f);
    if (ab < 0x1p-14f) { *out = x; return 1; }             // x (1 + x^2/3) rounds to x
    return ok;
}

FP_HD int fu_fast_asinh(float x, float* out) {
    const float ab = fabsf(x);
    const double a = fabs((double)x);
    const double u = fu_sqrt(fma(a, a, 1.0));
    const double N = a + u;
    const int j = fu_exponent_near_sqrt2((float)N);
    const double P = fu_pow2(j);
    // j == 0 (a < 0.354): (N - 1)/(N + 1) = a/(1 + u) exactly, which avoids the cancellation in N - 1
    const double num = j == 0 ? a + a : (a - P) + u;
    const double den = j == 0 ? fma(2.0, u, 2.0) : (a + P) + u;
    double r = fu_twice(fma((double)j, FU_K(FU_LN2_HALF), fu_atanh_small(fu_div(num, den))));
    r = FP_MAKE(FP_HI(r) | (x < 0.0f ? 0x80000000u : 0u), FP_LO(r));
    const int ok = fu_finish(r, out) & (ab <= 0x1p60f);   // |x| <= 2^60 (a^2 + 1 stays inside the float range)
    if (ab < 0x1p-14f) { *out = x; return 1; }             // x (1 - x^2/6) rounds to x
    return ok;
}

FP_HD int fu_fast_acosh(float x, float* out) {
    const double d = (double)x;
    const double u = fu_sqrt(fma(d, d, -1.0));                // x^2 - 1 with one rounding
    const double N = d + u;
    const int j = fu_exponent_near_sqrt2((float)N);
    const double P = fu_pow2(j);
    const double r = fu_twice(fma((double)j, FU_K(FU_LN2_HALF), fu_atanh_small(fu_div((d - P) + u, (d + P) + u))));
    return fu_finish(r, out) & (x > 1.0f) & (x <= 0x1p60f);  // 1 < x <= 2^60
}

// ---- sin, cos, tan ----------------------------------------------------------------------------------------------------
// sin(r) for |r| <= pi/2 (relative error ~2^-50)
FP_HD double fu_sin_half_pi(double r) {
    const double z = r * r;
    double p = FU_K(FU_SN);
    p = fma(p, z, FU_K(FU_SN + 1));
    p = fma(p, z, FU_K(FU_SN + 2));
    p = fma(p, z, FU_K(FU_SN + 3));
    p = fma(p, z, FU_K(FU_SN + 4));
    p = fma(p, z, FU_K(FU_SN + 5));
    p = fma(p, z, FU_K(FU_SN + 6));
    return fma(r * z, p, r);
}

// k = round(x / pi) by magic-number rounding in double, two-constant Cody-Waite reduction, one polynomial on |r| <= pi/2.
// Operands up to 32768; over that range |sin x|, |cos x| >= 4.1e-9 and |tan x| <= 2.4e8 for every float (exhaustive
// scan), so the results are normal floats and need no range test of their own.
FP_HD int fu_fast_sin(float x, float* out) {
    const float ab = fabsf(x);
    const double d = (double)x, magic = FU_K(FU_MAGIC);
    const double kd = fma(d, FU_K(FU_INVPI), magic);
    const double kf = kd - magic;
    const double r = fma(-kf, FU_K(FU_PI_LO), fma(-kf, FU_K(FU_PI_HI), d));
    const double v = fu_sin_half_pi(r);
    const double res = FP_MAKE(FP_HI(v) ^ (FP_LO(kd) << 31), FP_LO(v));
    const int ok = fu_finish(res, out) & (ab <= 32768.0f);
    if (ab < 0x1p-14f) { *out = x; return 1; }  // x (1 - x^2/6) rounds to x
    return ok;
}

FP_HD int fu_fast_cos(float x, float* out) {
    const double d = (double)x, magic = FU_K(FU_MAGIC);
    const double kd = fma(d, FU_K(FU_INVPI), 0.5) + magic;   // k = round(x/pi + 1/2): cos x = (-1)^k sin(x - (k - 1/2) pi)
    const double jf = (kd - magic) - 0.5;
    const double r = fma(-jf, FU_K(FU_PI_LO), fma(-jf, FU_K(FU_PI_HI), d));
    const double v = fu_sin_half_pi(r);
    const double res = FP_MAKE(FP_HI(v) ^ (FP_LO(kd) << 31), FP_LO(v));
    return fu_finish(res, out) & (fabsf(x) <= 32768.0f);
}

FP_HD int fu_fast_tan(float x, float* out) {
    const float ab = fabsf(x);
    const double d = (double)x, magic = FU_K(FU_MAGIC);
    const double kd = fma(d, FU_K(FU_2OPI), magic);
    const uint32_t k = FP_LO(kd);
    const double kf = kd - magic;
    const double r = fma(-kf, FU_K(FU_PIO2_LO), fma(-kf, FU_K(FU_PIO2_HI), d));
    const double z = r * r;
    double s = FU_K(FU_S4);
    s = fma(s, z, FU_K(FU_S4 + 1));
    s = fma(s, z, FU_K(FU_S4 + 2));
    s = fma(s, z, FU_K(FU_S4 + 3));
    s = fma(s, z, FU_K(FU_S4 + 4));
    double c = FU_K(FU_C4);
    c = fma(c, z, FU_K(FU_C4 + 1));
    c = fma(c, z, FU_K(FU_C4 + 2));
    c = fma(c, z, FU_K(FU_C4 + 3));
    c = fma(c, z, FU_K(FU_C4 + 4));
    c = fma(c, z, FU_K(FU_C4 + 5));
    const double sn = fma(r * z, s, r), cs = fma(z, c, 1.0);
    const int odd = (int)(k & 1u);
    const double num = odd ? -cs : sn, den = odd ? sn : cs;   // tan(r + k pi/2) = -cot r for odd k
    const double res = fu_div(num, den);
    const int ok = fu_finish(res, out) & (ab <= 32768.0f);
    if (ab < 0x1p-14f) { *out = x; return 1; }  // x (1 + x^2/3) rounds to x
    return ok;
}

// ---- atan, asin, acos -------------------------------------------------------------------------------------------------
// atan(X / Y) for 0 <= X <= Y: atan(b) = atan(c) + atan((b - c) / (1 + b c)) with c in {0, 7/16, 1} keeps |t| <= 0.22,
// and multiplying through by Y leaves ONE division: t = (X - c Y) / (Y + c X). Xf, Yf: float copies for the region tests.
FP_HD double fu_atan_ratio(double X, double Y, float Xf, float Yf) {
    const int p1 = Xf > 0.2f * Yf, p2 = Xf > 0.68f * Yf;
    const double c = p2 ? 1.0 : (p1 ? 0.4375 : 0.0);
    const double A = p2 ? FU_K(FU_ATAN_1) : (p1 ? FU_K(FU_ATAN_C1) : 0.0);  // rounding of atan(c): 2^-53 absolute
    const double t = fu_div(fma(-c, Y, X), fma(c, X, Y));
    const double z = t * t;
    double p = FU_K(FU_AT);
    p = fma(p, z, FU_K(FU_AT + 1));
    p = fma(p, z, FU_K(FU_AT + 2));
    p = fma(p, z, FU_K(FU_AT + 3));
    p = fma(p, z, FU_K(FU_AT + 4));
    p = fma(p, z, FU_K(FU_AT + 5));
    return A + fma(t * z, p, t);
}

FP_HD int fu_fast_atan(float x, float* out) {
    const float ab = fabsf(x);
    const float af = fabsf(x);
    const double a = (double)af;
    const int inv = af > 1.0f;  // atan(a) = pi/2 - atan(1/a)
    const double h = fu_atan_ratio(inv ? 1.0 : a, inv ? a : 1.0, inv ? 1.0f : af, inv ? af : 1.0f);
    double r = inv ? FU_K(FU_PIO2_HI) - h : h;
    r = FP_MAKE(FP_HI(r) | (x < 0.0f ? 0x80000000u : 0u), FP_LO(r));
    const int ok = fu_finish(r, out) & (ab <= 3.402823466e38f);
    if (ab < 0x1p-14f) { *out = x; return 1; }  // x (1 - x^2/3) rounds to x
    return ok;
}

// asin(x) = 2 atan(x / (1 + sqrt(1 - x^2))), acos(x) = 2 atan(sqrt(1 - x^2) / (1 + x)) for x >= 0, pi - acos(-x) below
FP_HD int fu_fast_asin(float x, float* out) {
    const float ab = fabsf(x);
    const float af = fabsf(x);
    const double a = (double)af;
    const double u = fu_sqrt(fma(-a, a, 1.0));  // 1 - x^2 with one rounding
    const double w = 1.0 + u;
    double r = fu_twice(fu_atan_ratio(a, w, af, (float)w));
    r = FP_MAKE(FP_HI(r) | (x < 0.0f ? 0x80000000u : 0u), FP_LO(r));
    const int ok = fu_finish(r, out) & (ab < 1.0f);
    if (ab < 0x1p-14f) { *out = x; return 1; }  // x (1 + x^2/6) rounds to x
    return ok;
}

FP_HD int fu_fast_acos(float x, float* out) {
    const float ab = fabsf(x);
    const float af = fabsf(x);
    const double a = (double)af;
    const double u = fu_sqrt(fma(-a, a, 1.0));
    const double w = 1.0 + a;
    const double h = fu_atan_ratio(u, w, (float)u, (float)w);
    const double r = x < 0.0f ? fma(-2.0, h, FU_K(FU_PI_HI)) : fu_twice(h);
    return fu_finish(r, out) & (ab < 1.0f);
}
#endif
