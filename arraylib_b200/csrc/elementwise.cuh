// Elementwise engine: out(contiguous) = f(in_0[strided], in_1[strided], ...).
//
// Replaces the reference loops array_array_op (arraylib/src/arraylib.c:569-589), array_scalar_op
// (:491-501), array_op (:638-645), array_deep_copy (:248-256) and the where family (:774-838),
// all of which call offset_indexer (:116-123) per element. Here the layout is folded once on the
// host (layout.cu) and run_functor() picks one of these sm_100a kernels:
//
//   ew_nd<VEC=4>     inner axis contiguous (stride 1) or broadcast (stride 0) in every operand:
//                    128-bit loads/stores, outer axes resolved with one mul-hi divmod per axis per
//                    vector. ndim==1 is the pure streaming kernel (config C2); stride-0 outer axes
//                    give the row-broadcast case (config C3, second add).
//   ew_outer         [I,1,W] o [1,J,W]: a 4x4 register tile per thread, 2/4 loads per output instead
//                    of 2 (config C3, first add).
//   ew_tile64        an operand is unit-stride along a different axis than the output (transposed
//                    view, config C5b): 64x64 tiles, 128-bit global accesses on both sides, 4x4
//                    register transposes, XOR-swizzled shared memory. ew_tile: the 4-byte version for
//                    extents that are not multiples of 4.
//   ew_interleave /  the same with a short, densely packed axis ([N,3] <-> [3,N]); ew_tile_rect for
//   ew_deinterleave  short axes that are not packed.
//   ew_run           anything else: 8 consecutive outputs per thread, every operand addressed in the cheapest
//                    way its layout allows (modes 0..4); ew_nd<VEC=1> for flat (1-D) streams.
//   ew_period        short inner axes under an operand outside ew_run's cheap modes ([N,1,7] + [1,9,7]): values
//                    or offsets of one period of the innermost axes tabulated in shared memory, four shifted
//                    copies so that 4 consecutive entries are always one aligned 128-bit load.
//   ew_heavy_stream  contiguous streams under pow (optionally the FP64 ufuncs): persistent CTAs that keep the
//                    next batch's loads in flight while they compute.
// The libm functors (exp ... atanh, pow) answer from the certified short FP64 kernels of fast_ufunc.h / fast_pow.h.
#pragma once
#include <type_traits>

#include "al_internal.h"
#include "device_utils.cuh"
#include "fast_pow.h"
#include "fast_ufunc.h"

namespace al {


// ---- scalar functions: bit-for-bit the reference's (arraylib.c:12-49) ---------------------------
// NaN results follow the reference's x86-64 build as well (probed through the compiled oracle; only the
// rare NaN lane takes these paths): an SSE arithmetic op returns its FIRST NaN operand, quieted, and the
// negative default NaN 0xFFC00000 when it creates one (inf - inf, 0 * inf, 0 / 0); glibc's double
// functions hand a NaN argument back (quieted by the float -> double conversion) and create a negative
// NaN on a domain error, except asin / acos, whose NaN is positive. The GPU would return the canonical
// 0x7FFFFFFF everywhere.
// Sign-bit edits as opaque integer instructions: written in C++ the optimiser turns `bits ^ 0x80000000` into
// neg.f32 and `bits & 0x7fffffff` into abs.f32, and those return the canonical NaN for a NaN input.
__device__ __forceinline__ float flip_sign_bit(float x) {
    uint32_t b;
    asm("xor.b32 %0, %1, 0x80000000;" : "=r"(b) : "r"(__float_as_uint(x)));
    return __uint_as_float(b);
}
__device__ __forceinline__ float clear_sign_bit(float x) {
    uint32_t b;
    asm("and.b32 %0, %1, 0x7fffffff;" : "=r"(b) : "r"(__float_as_uint(x)));
    return __uint_as_float(b);
}
__device__ __forceinline__ float nan_quiet(float x) {
    uint32_t b;
    asm("or.b32 %0, %1, 0x00400000;" : "=r"(b) : "r"(__float_as_uint(x)));
    return __uint_as_float(b);
}
__device__ __forceinline__ float nan_default_neg() { return __uint_as_float(0xFFC00000u); }
__device__ __forceinline__ float nan_default_pos() { return __uint_as_float(0x7FC00000u); }
__device__ __forceinline__ float nan_like_sse(float res, float l, float r) {
    if (res == res) return res;
    if (l != l) return nan_quiet(l);
    if (r != r) return nan_quiet(r);
    return nan_default_neg();
}

// (float)pow((double)l, (double)r) as the reference computes it (arraylib.c:17): fast_pow.h answers the elements
// whose float rounding is certain (~40 FP64 operations), the rest go through the full double-precision pow.
__device__ __forceinline__ float pow_like_reference(float l, float r) {
    float fast;
    if (fp_fast_pow(l, r, &fast)) return fast;
    return (float)pow((double)l, (double)r);
}

// *_raw: the arithmetic alone (a NaN result is whatever the GPU makes of it: the canonical 0x7FFFFFFF);
// *_apply: the same value, and when it is a NaN, the NaN the reference's build would have produced.
template <int OP>
__device__ __forceinline__ float binop_raw(float l, float r) {
    if constexpr (OP == AL_SUM) return __fadd_rn(l, r);
    else if constexpr (OP == AL_SUB) return __fsub_rn(l, r);
    else if constexpr (OP == AL_MUL) return __fmul_rn(l, r);
    else if constexpr (OP == AL_DIV) return __fdiv_rn(l, r);
    else if constexpr (OP == AL_MOD) return fmodf(l, r);  // fmod is exact, so f32 == round(f64)
    else if constexpr (OP == AL_POW) return pow_like_reference(l, r);
    // gcc's inline fmax / fmin: a NaN on the left yields the right operand (whatever it is), a NaN on the right
    // yields the left one, and on a +-0 tie the LEFT operand is returned
    // (FMNMX ignores a NaN operand like fmax / fmin do; only the tie -- +0 against -0 -- needs the left operand put back.
    // Two NaNs give the canonical NaN here; binop_apply replaces it by the right operand, which is what the reference returns.)
    else if constexpr (OP == AL_MAX) return l == r ? l : fmaxf(l, r);
    else if constexpr (OP == AL_MIN) return l == r ? l : fminf(l, r);
    else if constexpr (OP == AL_EQ) return l == r ? 1.0f : 0.0f;
    else if constexpr (OP == AL_NE) return l != r ? 1.0f : 0.0f;
    else if constexpr (OP == AL_GT) return l > r ? 1.0f : 0.0f;
    else if constexpr (OP == AL_GE) return l >= r ? 1.0f : 0.0f;
    else if constexpr (OP == AL_LT) return l < r ? 1.0f : 0.0f;
    else return l <= r ? 1.0f : 0.0f;
}

template <int OP>
__device__ __forceinline__ float binop_apply(float l, float r) {
    const float res = binop_raw<OP>(l, r);
    if (res == res) return res;
    if constexpr (OP == AL_MOD) {
        if (l != l && r != r) {  // x87 fprem rule: the NaN with the larger (quieted) significand; on a tie the positive one
            const uint32_t fl = (__float_as_uint(l) & 0x007fffffu) | 0x00400000u, fr = (__float_as_uint(r) & 0x007fffffu) | 0x00400000u;
            const bool take_r = fr > fl || (fr == fl && (__float_as_uint(l) >> 31));
            return nan_quiet(take_r ? r : l);
        }
        return nan_like_sse(res, l, r);
    } else if constexpr (OP == AL_POW) {
        // glibc pow: a NaN base comes back as x*x (sign kept) and is NEGATED when the base is "negative" and the
        // exponent an odd integer; a NaN exponent comes back as x + y; a domain error creates the negative NaN
        if (l != l) {
            const float q = nan_quiet(l);
            const float ar = fabsf(r);
            const bool odd = r == r && ar < 16777216.0f && ar == truncf(ar) && (((int)ar) & 1);
            return ((__float_as_uint(l) >> 31) && odd) ? flip_sign_bit(q) : q;
        }
        return r != r ? nan_quiet(r) : nan_default_neg();
    } else if constexpr (OP == AL_MAX || OP == AL_MIN) {
        return nan_quiet(r);  // both operands are NaN: the right one, quieted by fmax's float -> double conversion
    } else {
        return nan_like_sse(res, l, r);
    }
}

// The reference promotes to double, calls libm and rounds back (arraylib.c:31-49). sqrt/ceil/
// floor/abs/neg are exact under that double rounding, so they stay in f32; the transcendentals
// run in FP64 on the device and round once.
// The certified fast path of a libm ufunc (fast_ufunc.h): true and *out when the float rounding of the result is certain
// (bit-identical to the reference's glibc result, validated over all 2^32 operands on the CPU). Accepted results are
// finite, so the NaN rules below never apply to them.
template <int OP>
__device__ __forceinline__ bool ufunc_fast(float v, float* out) {
    if constexpr (OP == AL_EXP) return fu_fast_exp(v, out);
    else if constexpr (OP == AL_LOG) return fu_fast_log(v, out);
    else if constexpr (OP == AL_SIN) return fu_fast_sin(v, out);
    else if constexpr (OP == AL_COS) return fu_fast_cos(v, out);
    else if constexpr (OP == AL_TAN) return fu_fast_tan(v, out);
    else if constexpr (OP == AL_ASIN) return fu_fast_asin(v, out);
    else if constexpr (OP == AL_ACOS) return fu_fast_acos(v, out);
    else if constexpr (OP == AL_ATAN) return fu_fast_atan(v, out);
    else if constexpr (OP == AL_SINH) return fu_fast_sinh(v, out);
    else if constexpr (OP == AL_COSH) return fu_fast_cosh(v, out);
    else if constexpr (OP == AL_TANH) return fu_fast_tanh(v, out);
    else if constexpr (OP == AL_ASINH) return fu_fast_asinh(v, out);
    else if constexpr (OP == AL_ACOSH) return fu_fast_acosh(v, out);
    else if constexpr (OP == AL_ATANH) return fu_fast_atanh(v, out);
    else return false;
}

template <int OP>
__device__ __forceinline__ float ufunc_raw(float v) {
    if constexpr (OP == AL_NEG) return flip_sign_bit(v);  // bitwise in the reference too (a NaN keeps its payload, unquieted)
    else if constexpr (OP == AL_ABS) return clear_sign_bit(v);
    else if constexpr (OP == AL_SQRT) return __fsqrt_rn(v);
    else if constexpr (OP == AL_CEIL) return ceilf(v);
    else if constexpr (OP == AL_FLOOR) return floorf(v);
    else {
        // the ~0.012 % of the elements the fast path declines, and out-of-range operands: CUDA's full double routine
        double x = (double)v;
        if constexpr (OP == AL_EXP) x = exp(x);
        else if constexpr (OP == AL_LOG) x = log(x);
        else if constexpr (OP == AL_SIN) x = sin(x);
        else if constexpr (OP == AL_COS) x = cos(x);
        else if constexpr (OP == AL_TAN) x = tan(x);
        else if constexpr (OP == AL_ASIN) x = asin(x);
        else if constexpr (OP == AL_ACOS) x = acos(x);
        else if constexpr (OP == AL_ATAN) x = atan(x);
        else if constexpr (OP == AL_SINH) x = sinh(x);
        else if constexpr (OP == AL_COSH) x = cosh(x);
        else if constexpr (OP == AL_TANH) x = tanh(x);
        else if constexpr (OP == AL_ASINH) x = asinh(x);
        else if constexpr (OP == AL_ACOSH) x = acosh(x);
        else x = atanh(x);
        return (float)x;
    }
}

// the value as the fused chains compute it (fast path first, same as the eager kernels; a NaN is whatever the GPU makes of it)
template <int OP>
__device__ __forceinline__ float ufunc_value(float v) {
    if constexpr (OP >= AL_EXP && OP <= AL_ATANH) {
        float fast = 0.0f;
        if (ufunc_fast<OP>(v, &fast)) return fast;
    }
    return ufunc_raw<OP>(v);
}

template <int OP>
__device__ __forceinline__ float ufunc_apply(float v) {
    if constexpr (OP >= AL_EXP && OP <= AL_ATANH) {
        float fast = 0.0f;
        if (ufunc_fast<OP>(v, &fast)) return fast;
    }
    const float res = ufunc_raw<OP>(v);
    if constexpr (OP == AL_NEG) return res;
    if (res == res) return res;
    if (v != v) return nan_quiet(OP == AL_ABS ? res : v);  // fabs((double)x): the conversion quiets a signalling NaN
    return (OP == AL_ASIN || OP == AL_ACOS) ? nan_default_pos() : nan_default_neg();
}

// Functors: NIN inputs in, one float out.
template <int OP>
struct BinF {
    static constexpr int NIN = 2;
    static constexpr int kHeavy = OP == AL_POW ? 1 : 0;  // FP64 work per element: contiguous streams take ew_heavy_stream_kernel (level: see heavy_stream)
    __device__ __forceinline__ float operator()(const float* v) const { return binop_apply<OP>(v[0], v[1]); }
    __device__ __forceinline__ float raw(const float* v) const { return binop_raw<OP>(v[0], v[1]); }  // the value; a NaN not yet the reference's
};
template <int OP>
struct ScalarF {
    static constexpr int NIN = 1;
    static constexpr int kHeavy = OP == AL_POW ? 1 : 0;
    float s;
    __device__ __forceinline__ float operator()(const float* v) const { return binop_apply<OP>(v[0], s); }
    __device__ __forceinline__ float raw(const float* v) const { return binop_raw<OP>(v[0], s); }
};
template <int OP>
struct UnaryF {
    static constexpr int NIN = 1;
    static constexpr int kHeavy = OP >= AL_EXP && OP <= AL_ATANH ? 2 : 0;
    __device__ __forceinline__ float operator()(const float* v) const { return ufunc_apply<OP>(v[0]); }
    __device__ __forceinline__ float raw(const float* v) const { return ufunc_apply<OP>(v[0]); }  // (neg / abs must keep NaN payloads: no shortcut)
};
template <typename F, typename = void>
struct is_heavy : std::false_type {};
template <typename F>
struct is_heavy<F, std::void_t<decltype(F::kHeavy)>> : std::bool_constant<(F::kHeavy > 0)> {};

struct CopyF {
    static constexpr int NIN = 1;
    __device__ __forceinline__ float operator()(const float* v) const { return v[0]; }
    __device__ __forceinline__ float raw(const float* v) const { return v[0]; }
};
// where: C truthiness of an f32, i.e. (x != 0): NaN selects lhs, +-0 selects rhs (arraylib.c:788)
struct WhereAAA {
    static constexpr int NIN = 3;
    __device__ __forceinline__ float operator()(const float* v) const { return v[0] != 0.0f ? v[1] : v[2]; }
    __device__ __forceinline__ float raw(const float* v) const { return (*this)(v); }
};
struct WhereAAS {
    static constexpr int NIN = 2;
    float r;
    __device__ __forceinline__ float operator()(const float* v) const { return v[0] != 0.0f ? v[1] : r; }
    __device__ __forceinline__ float raw(const float* v) const { return (*this)(v); }
};
struct WhereASA {
    static constexpr int NIN = 2;
    float l;
    __device__ __forceinline__ float operator()(const float* v) const { return v[0] != 0.0f ? l : v[1]; }
    __device__ __forceinline__ float raw(const float* v) const { return (*this)(v); }
};
struct WhereASS {
    static constexpr int NIN = 1;
    float l, r;
    __device__ __forceinline__ float operator()(const float* v) const { return v[0] != 0.0f ? l : r; }
    __device__ __forceinline__ float raw(const float* v) const { return (*this)(v); }
};

// ---- N-d kernel --------------------------------------------------------------------------------
template <int NIN>
struct NdArgs {
    float* out;
    const float* in[NIN];
    uint32_t n_items;  // vector items in this launch (< 2^31)
    int ndim;
    uint32_t extent[kMaxDims];  // extent[0] counted in vector items
    FastDiv div[kMaxDims];
    long long stride[NIN][kMaxDims];  // element strides; stride[.][0] is pre-multiplied by VEC
    unsigned char inner_bcast[NIN];   // operand has stride 0 along the inner axis
    unsigned char reuse[NIN];         // operand is re-read (has a stride-0 axis): keep it cached
    unsigned char stream_store;
    unsigned char load_policy;  // 0: ld.global.cs (evict-first), 1: ld.global.nc, 2: nc + L1::no_allocate
    unsigned char out_aligned;  // ew_run: the output pointer is 16-byte aligned (128-bit stores allowed)
    // ew_run: how each operand is addressed. 0 = congruent with the output and 16-byte aligned (one 128-bit load),
    // 1 = "block broadcast": unit-stride over one run of axes and stride 0 over all others, so that the offset
    // of output element e is (e / blk_inner) mod blk_len, 3 = broadcast "in the middle" (contiguous below and above
    // a run of stride-0 axes, the [N,1,C] of pairwise [N,1,C] o [1,M,C]): offset = e mod blk_inner + blk_inner *
    // (e / blk_len), 2 = anything else (row offsets + inner stride).
    unsigned char mode[NIN];
    uint32_t blk_inner[NIN], blk_len[NIN];
    FastDiv blk_inner_div[NIN], blk_len_div[NIN];
};

template <int VEC>
struct VecT;
template <>
struct VecT<1> {
    using type = float;
};
template <>
struct VecT<4> {
    using type = float4;
};

__device__ __forceinline__ float4 ld_nc_na(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

template <int VEC>
__device__ __forceinline__ typename VecT<VEC>::type load_vec(const float* p, bool bcast, bool reuse, int policy) {
    if constexpr (VEC == 1) {
        return reuse ? __ldg(p) : __ldcs(p);
    } else {
        if (bcast) {
            float s = __ldg(p);
            return make_float4(s, s, s, s);
        }
        const float4* q = reinterpret_cast<const float4*>(p);
        if (reuse || policy == 1) return __ldg(q);
        if (policy == 2) return ld_nc_na(q);
        return __ldcs(q);
    }
}

template <int VEC>
__device__ __forceinline__ void store_vec(float* p, typename VecT<VEC>::type v, bool streaming) {
    if constexpr (VEC == 1) {
        if (streaming) __stcs(p, v);
        else *p = v;
    } else {
        if (streaming) __stcs(reinterpret_cast<float4*>(p), v);
        else *reinterpret_cast<float4*>(p) = v;
    }
}

constexpr int kEwThreads = 256;

// Resident CTAs ptxas is asked to make room for: without a hint it chose 40 registers AND spills for the libm functors.
// 5 CTAs (48 registers) fit every libm ufunc without a spill; pow needs 64.
template <typename F>
constexpr int ew_min_blocks() {
    if constexpr (is_heavy<F>::value) return F::kHeavy == 2 ? 5 : 3;
    else return 0;  // 0 = no hint: the streaming functors compile exactly as before
}

template <int VEC, int UNROLL, typename F>
__global__ void __launch_bounds__(kEwThreads, ew_min_blocks<F>()) ew_nd_kernel(const NdArgs<F::NIN> a, const F f) {
    constexpr int NIN = F::NIN;
    using V = typename VecT<VEC>::type;
    const uint32_t base = blockIdx.x * (uint32_t)(kEwThreads * UNROLL) + threadIdx.x;
    V v[UNROLL][NIN];
    // Phase 1: issue every load of this thread before any use (UNROLL*NIN requests in flight).
#pragma unroll
    for (int k = 0; k < UNROLL; k++) {
        const uint32_t item = base + k * kEwThreads;
        if (item < a.n_items) {
            long long off[NIN];
#pragma unroll
            for (int i = 0; i < NIN; i++) off[i] = 0;
            uint32_t rem = item;
#pragma unroll
            for (int d = 0; d < kMaxDims; d++) {
                uint32_t idx;
                const bool last = (d == a.ndim - 1);
                if (last) {
                    idx = rem;
                } else {
                    const uint32_t q = fdiv(rem, a.div[d]);
                    idx = rem - q * a.extent[d];
                    rem = q;
                }
#pragma unroll
                for (int i = 0; i < NIN; i++) off[i] += (long long)idx * a.stride[i][d];
                if (last) break;
            }
#pragma unroll
            for (int i = 0; i < NIN; i++) v[k][i] = load_vec<VEC>(a.in[i] + off[i], a.inner_bcast[i], a.reuse[i], a.load_policy);
        }
    }
    // Phase 2: compute and store (output is contiguous: item -> item*VEC).
#pragma unroll
    for (int k = 0; k < UNROLL; k++) {
        const uint32_t item = base + k * kEwThreads;
        if (item < a.n_items) {
            V r;
            if constexpr (VEC == 1) {
                float x[NIN];
#pragma unroll
                for (int i = 0; i < NIN; i++) x[i] = v[k][i];
                r = f(x);
            } else {
                float x[NIN];
#pragma unroll
                for (int i = 0; i < NIN; i++) x[i] = v[k][i].x;
                r.x = f(x);
#pragma unroll
                for (int i = 0; i < NIN; i++) x[i] = v[k][i].y;
                r.y = f(x);
#pragma unroll
                for (int i = 0; i < NIN; i++) x[i] = v[k][i].z;
                r.z = f(x);
#pragma unroll
                for (int i = 0; i < NIN; i++) x[i] = v[k][i].w;
                r.w = f(x);
            }
            store_vec<VEC>(a.out + (size_t)item * VEC, r, a.stream_store);
        }
    }
}

// ---- contiguous streams under a compute-heavy functor (the FP64 ufuncs, pow) ----------------------------------------
// ew_nd issues all of a thread's loads, waits, computes, stores and exits. With 40-130 instructions per element the
// compute phase is as long as the memory latency, and at 48-64 registers only 8 warps per scheduler are resident, so
// DRAM sat idle while the warps computed: exp ran at 4.5 TB/s with the FP64 pipe and the issue slots both ~60 % busy.
// Here the CTAs are persistent and every thread keeps the NEXT batch's 128-bit loads in flight while it computes the
// current one (register double buffer), so memory time and compute time overlap instead of adding up.
constexpr int kHeavyUnroll = 2;

template <typename F>
__global__ void __launch_bounds__(kEwThreads) ew_heavy_stream_kernel(float* __restrict__ out, const float* __restrict__ in0,
                                                                     const float* __restrict__ in1, uint32_t n4, const F f,
                                                                     int stream_store) {
    constexpr int NIN = F::NIN, U = kHeavyUnroll;
    static_assert(NIN <= 2, "one or two streamed operands");
    const uint32_t step = gridDim.x * (uint32_t)(kEwThreads * U);
    uint32_t i = blockIdx.x * (uint32_t)(kEwThreads * U) + threadIdx.x;
    float4 cur[U][NIN], nxt[U][NIN];
    auto fetch = [&](float4 (*dst)[NIN], uint32_t at) {
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t j = at + u * kEwThreads;
            if (j < n4) {
                dst[u][0] = __ldcs(reinterpret_cast<const float4*>(in0) + j);
                if constexpr (NIN == 2) dst[u][1] = __ldcs(reinterpret_cast<const float4*>(in1) + j);
            }
        }
    };
    fetch(cur, i);
    for (; i < n4; i += step) {
        fetch(nxt, i + step);
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t j = i + u * kEwThreads;
            if (j < n4) {
                float4 r;
                float x[NIN];
#pragma unroll
                for (int k = 0; k < NIN; k++) x[k] = cur[u][k].x;
                r.x = f(x);
#pragma unroll
                for (int k = 0; k < NIN; k++) x[k] = cur[u][k].y;
                r.y = f(x);
#pragma unroll
                for (int k = 0; k < NIN; k++) x[k] = cur[u][k].z;
                r.z = f(x);
#pragma unroll
                for (int k = 0; k < NIN; k++) x[k] = cur[u][k].w;
                r.w = f(x);
                if (stream_store) __stcs(reinterpret_cast<float4*>(out) + j, r);
                else reinterpret_cast<float4*>(out)[j] = r;
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int k = 0; k < NIN; k++) cur[u][k] = nxt[u][k];
    }
}

// ---- general N-d kernel, 4 consecutive outputs per thread ------------------------------------------------
// The fallback for everything the 128-bit plans cannot take (odd inner extents, misaligned rows,
// arbitrary as_strided operands). Resolving a linear index costs a divmod chain per operand; with one
// output per thread that put ~60-90 instructions on every element and capped write-bound cases at
// 1.5-2 TB/s. Here a thread owns 8 consecutive outputs (two 128-bit stores when the output is aligned)
// and each operand is addressed in the cheapest way its layout allows:
//   mode 0  congruent with the output (the big stream of "(N,3) + (3,)"): one 128-bit load, no index;
//   mode 1  block broadcast -- unit-stride over one run of axes, stride 0 over the rest, which is what
//           NumPy-style broadcasting of a contiguous array produces when its non-1 extents are adjacent
//           ((3,), (N,1), (1,C,1,1), a scalar ...): offset(e) = (e / inner) mod len, resolved with two
//           mul-hi divisions per THREAD and stepped with two compare/selects per element;
//   mode 4  mode 1 with inner = 1, a period that is a multiple of 8 and an aligned base: two 128-bit loads;
//   mode 3  broadcast in the middle ([N,1,C] against [1,M,C] with a small odd C): offset(e) = e mod C + C * (e / (M*C)),
//           three divisions per thread, three compare/selects per element;
//   mode 2  anything else: the 4 outputs lie in at most three consecutive rows (a row = one pass over
//           the inner axis), so the offsets of rows q, q+1 (q+2 only when the inner extent is 2) are
//           resolved once and every element is row offset + inner index * inner stride.
constexpr int kRun = 8;  // outputs per thread: the per-thread index set-up of modes 1-3 is paid once per 8 outputs

template <int NIN>
__device__ __forceinline__ void run_row_offsets(const NdArgs<NIN>& a, uint32_t q, long long* off) {
#pragma unroll
    for (int i = 0; i < NIN; i++) off[i] = 0;
    uint32_t rem = q;
#pragma unroll
    for (int d = 1; d < kMaxDims; d++) {
        if (d >= a.ndim) break;
        uint32_t idx;
        if (d == a.ndim - 1) {
            idx = rem;
        } else {
            const uint32_t qq = fdiv(rem, a.div[d]);
            idx = rem - qq * a.extent[d];
            rem = qq;
        }
#pragma unroll
        for (int i = 0; i < NIN; i++) off[i] += (long long)idx * a.stride[i][d];
    }
}

template <typename F>
__global__ void __launch_bounds__(kEwThreads) ew_run_kernel(const NdArgs<F::NIN> a, const F f) {
    constexpr int NIN = F::NIN;
    const uint32_t e0 = (blockIdx.x * (uint32_t)kEwThreads + threadIdx.x) * kRun;
    if (e0 >= a.n_items) return;
    const bool full = a.out_aligned && e0 + kRun <= a.n_items;
    float v[kRun][NIN];
    bool general = false;
#pragma unroll
    for (int i = 0; i < NIN; i++) {
        if (a.mode[i] == 0 && full) {
#pragma unroll
            for (int h = 0; h < kRun / 4; h++) {
                const float4 d = __ldcs(reinterpret_cast<const float4*>(a.in[i] + e0) + h);
                v[4 * h][i] = d.x, v[4 * h + 1][i] = d.y, v[4 * h + 2][i] = d.z, v[4 * h + 3][i] = d.w;
            }
        } else if (a.mode[i] == 4 && full) {
            // periodic, contiguous, period a multiple of kRun, aligned: the run never wraps and is two 128-bit loads
            const uint32_t o = e0 - fdiv(e0, a.blk_len_div[i]) * a.blk_len[i];
#pragma unroll
            for (int h = 0; h < kRun / 4; h++) {
                const float4 d = __ldg(reinterpret_cast<const float4*>(a.in[i] + o) + h);
                v[4 * h][i] = d.x, v[4 * h + 1][i] = d.y, v[4 * h + 2][i] = d.z, v[4 * h + 3][i] = d.w;
            }
        } else if (a.mode[i] <= 1 || a.mode[i] == 4) {
            // mode 0 / 4 on a ragged tail is the block formula with inner = 1, len = all (set up by the host)
            const uint32_t inner = a.blk_inner[i], len = a.blk_len[i];
            const uint32_t t = fdiv(e0, a.blk_inner_div[i]);
            uint32_t c = e0 - t * inner;
            uint32_t o = t - fdiv(t, a.blk_len_div[i]) * len;
#pragma unroll
            for (int j = 0; j < kRun; j++) {
                if (e0 + j < a.n_items) v[j][i] = __ldg(a.in[i] + o);
                if (++c == inner) {
                    c = 0;
                    if (++o == len) o = 0;
                }
            }
        } else if (a.mode[i] == 3) {
            const uint32_t p1 = a.blk_inner[i], p2 = a.blk_len[i];
            const uint32_t q2 = fdiv(e0, a.blk_len_div[i]);
            uint32_t r2 = e0 - q2 * p2;
            uint32_t r1 = r2 - fdiv(r2, a.blk_inner_div[i]) * p1;
            long long o = (long long)q2 * p1 + r1;
#pragma unroll
            for (int j = 0; j < kRun; j++) {
                if (e0 + j < a.n_items) v[j][i] = __ldg(a.in[i] + o);
                o++, r1++, r2++;
                if (r1 == p1) r1 = 0, o -= p1;
                if (r2 == p2) r2 = 0, o += p1;
            }
        } else {
            general = true;
        }
    }
    if (general) {
        const uint32_t ext0 = a.extent[0];
#pragma unroll
        for (int h = 0; h < kRun / 4; h++) {  // four outputs at a time: they lie in at most three consecutive rows
            const uint32_t eh = e0 + 4 * h;
            if (eh < a.n_items) {
                const uint32_t q0 = fdiv(eh, a.div[0]);
                const uint32_t r0 = eh - q0 * ext0;
                long long row[3][NIN];
                run_row_offsets<NIN>(a, q0, row[0]);
                if (r0 + 3 >= ext0) {  // the run crosses into the next row(s): rare when rows are long
                    run_row_offsets<NIN>(a, q0 + 1, row[1]);
                    if (ext0 == 2) {
                        run_row_offsets<NIN>(a, q0 + 2, row[2]);
                    } else {
#pragma unroll
                        for (int i = 0; i < NIN; i++) row[2][i] = row[1][i];
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < NIN; i++) row[1][i] = row[2][i] = row[0][i];
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    uint32_t r = r0 + j;
                    int c = 0;
                    if (r >= ext0) r -= ext0, c = 1;
                    if (r >= ext0) r -= ext0, c = 2;
                    if (eh + j < a.n_items) {
#pragma unroll
                        for (int i = 0; i < NIN; i++) {
                            if (a.mode[i] == 2) {
                                const long long base = c == 0 ? row[0][i] : (c == 1 ? row[1][i] : row[2][i]);
                                v[4 * h + j][i] = __ldg(a.in[i] + base + (long long)r * a.stride[i][0]);
                            }
                        }
                    }
                }
            }
        }
    }
    float r[kRun];
#pragma unroll
    for (int j = 0; j < kRun; j++) r[j] = (e0 + j < a.n_items) ? f(v[j]) : 0.0f;
    float* dst = a.out + e0;
    if (full) {
#pragma unroll
        for (int h = 0; h < kRun / 4; h++) {
            const float4 r4 = make_float4(r[4 * h], r[4 * h + 1], r[4 * h + 2], r[4 * h + 3]);
            if (a.stream_store) __stcs(reinterpret_cast<float4*>(dst) + h, r4);
            else reinterpret_cast<float4*>(dst)[h] = r4;
        }
    } else {
#pragma unroll
        for (int j = 0; j < kRun; j++)
            if (e0 + j < a.n_items) dst[j] = r[j];
    }
}

// ---- period kernel: short inner axes under operands that fit none of ew_run's cheap modes ---------------------------
// "[4096,1,5,3] + [1,2048,1,3]", "[N,1,7] + [1,9,7]": with inner extents of 3..7 the four outputs of an ew_run thread
// span up to three rows and every row costs a divmod chain per operand (1.2-3 TB/s, issue-bound). Here the innermost
// axes -- whole axes while their product stays <= kPerMax, then the largest divisor of the next axis that still fits --
// form a PERIOD of P outputs. Each CTA tabulates, once, what every operand does inside a period: its VALUES when the
// operand does not depend on anything above the period (kind 1: the broadcast operand's period staged in shared
// memory), else its element OFFSETS (kind 2). A thread then owns 4 consecutive outputs, needs one division (e / P) and
// one chain over the OUTER axes only, and reads its 4 table entries with one 128-bit shared load: the table is kept in
// four copies shifted by 0..3 entries (and extended by 3 entries past the wrap), so the load is aligned for every
// phase of e mod P. Only threads whose 4 outputs cross into the next period resolve a second outer chain.
constexpr int kPerMax = 512;

template <int NIN>
struct PerArgs {
    float* out;
    const float* in[NIN];
    uint32_t n_items, period;
    FastDiv period_div;
    int n_inner, n_outer;
    uint32_t in_ext[kMaxDims], out_ext[kMaxDims];
    FastDiv in_div[kMaxDims], out_div[kMaxDims];
    uint32_t in_stride[NIN][kMaxDims], out_stride[NIN][kMaxDims];  // element strides; every operand spans < 2^32 elements
    unsigned char kind[NIN];  // 0: laid out like the output (128-bit loads), 1: values staged, 2: offsets staged,
                              // 4: contiguous with period blk_len (a multiple of 4, aligned base): offset = e mod blk_len, one 128-bit load
    uint32_t blk_len[NIN];
    FastDiv blk_len_div[NIN];
    unsigned char stream_store, out_aligned;
};

template <int NIN>
__device__ __forceinline__ void per_outer_offsets(const PerArgs<NIN>& a, uint32_t q, uint32_t* off) {
#pragma unroll
    for (int i = 0; i < NIN; i++) off[i] = 0;
    uint32_t rem = q;
#pragma unroll
    for (int d = 0; d < kMaxDims; d++) {
        if (d >= a.n_outer) break;
        uint32_t idx;
        if (d == a.n_outer - 1) {
            idx = rem;
        } else {
            const uint32_t qq = fdiv(rem, a.out_div[d]);
            idx = rem - qq * a.out_ext[d];
            rem = qq;
        }
#pragma unroll
        for (int i = 0; i < NIN; i++) off[i] += idx * a.out_stride[i][d];
    }
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 t;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "r"(addr));
    return t;
}

// Operand values of one group of 4 consecutive outputs starting at e0. GUARD: the ragged last group (per-element bounds
// checks); K0..K2: operand kinds known at compile time (3 = read a.kind[i]), which removes every per-operand branch.
// The table entries past the wrap (m >= P) of an offset table already contain the step of the first outer axis, so the
// offset of an element is table + outer(q) unless the group wraps exactly where that outer axis carries (rare: fixed up).
template <typename F, int K0, int K1, int K2, bool GUARD>
__device__ __forceinline__ void per_load(const PerArgs<F::NIN>& a, uint32_t tbl_base, uint32_t e0, float (*v)[F::NIN]) {
    constexpr int NIN = F::NIN;
    constexpr int KS[3] = {K0, K1, K2};
    const uint32_t P = a.period;
    const uint32_t q = fdiv(e0, a.period_div), r = e0 - q * P;
    const uint32_t s = r & 3u, j = r - s;
    uint32_t o0[NIN];
    bool fix = false;
    if (a.n_outer == 1) {  // the usual case: one axis above the period, no carries
#pragma unroll
        for (int i = 0; i < NIN; i++) o0[i] = q * a.out_stride[i][0];
    } else if (a.n_outer == 2) {  // two axes above the period (one of them usually the remainder of a split axis)
        const uint32_t q1 = fdiv(q, a.out_div[0]), i0 = q - q1 * a.out_ext[0];
#pragma unroll
        for (int i = 0; i < NIN; i++) o0[i] = i0 * a.out_stride[i][0] + q1 * a.out_stride[i][1];
        fix = r + 3 >= P && i0 == a.out_ext[0] - 1;
    } else {
        per_outer_offsets<NIN>(a, q, o0);
        const uint32_t q1 = fdiv(q, a.out_div[0]);
        fix = r + 3 >= P && q - q1 * a.out_ext[0] == a.out_ext[0] - 1;
    }
    const uint32_t row = tbl_base + s * (uint32_t)((kPerMax + 4) * 4) + j * 4u;
#pragma unroll
    for (int i = 0; i < NIN; i++) {
        const int kind = KS[i] == 3 ? (int)a.kind[i] : KS[i];
        if (kind == 0) {
            if (!GUARD) {
                const float4 d = __ldcs(reinterpret_cast<const float4*>(a.in[i] + e0));
                v[0][i] = d.x, v[1][i] = d.y, v[2][i] = d.z, v[3][i] = d.w;
            } else {
#pragma unroll
                for (int c = 0; c < 4; c++)
                    if (e0 + c < a.n_items) v[c][i] = __ldg(a.in[i] + e0 + c);
            }
        } else if (kind == 4) {
            const uint32_t o = e0 - fdiv(e0, a.blk_len_div[i]) * a.blk_len[i];  // a multiple of 4; the group never wraps
            if (!GUARD) {
                const float4 d = __ldg(reinterpret_cast<const float4*>(a.in[i] + o));
                v[0][i] = d.x, v[1][i] = d.y, v[2][i] = d.z, v[3][i] = d.w;
            } else {
#pragma unroll
                for (int c = 0; c < 4; c++)
                    if (e0 + c < a.n_items) v[c][i] = __ldg(a.in[i] + o + c);
            }
        } else {
            const uint4 t = lds128(row + (uint32_t)(i * 4 * (kPerMax + 4) * 4));
            uint32_t tv[4] = {t.x, t.y, t.z, t.w};
            if (kind == 1) {
#pragma unroll
                for (int c = 0; c < 4; c++) v[c][i] = __uint_as_float(tv[c]);
            } else {
                if (fix) {  // elements past the wrap: drop the baked-in step, use the offsets of the next outer index
                    uint32_t o1[NIN];
                    per_outer_offsets<NIN>(a, q + 1, o1);
#pragma unroll
                    for (int c = 0; c < 4; c++)
                        if (r + c >= P) tv[c] += o1[i] - o0[i] - a.out_stride[i][0];
                }
#pragma unroll
                for (int c = 0; c < 4; c++)
                    if (!GUARD || e0 + c < a.n_items) v[c][i] = __ldg(a.in[i] + (tv[c] + o0[i]));
            }
        }
    }
}

template <typename F, bool GUARD>
__device__ __forceinline__ void per_store(const PerArgs<F::NIN>& a, const F& f, uint32_t e0, float (*v)[F::NIN]) {
    float* dst = a.out + e0;
    if (!GUARD) {
        // the bare arithmetic first; one NaN test for the four results (a NaN survives the sum; inf - inf only costs a
        // detour), and only then the functor's own NaN rules (sign and payload of the reference's build)
        float y[4];
#pragma unroll
        for (int c = 0; c < 4; c++) y[c] = f.raw(v[c]);
        const float probe = (y[0] + y[1]) + (y[2] + y[3]);
        if (probe != probe) {
#pragma unroll
            for (int c = 0; c < 4; c++) y[c] = f(v[c]);
        }
        if (a.out_aligned) {
            const float4 r4 = make_float4(y[0], y[1], y[2], y[3]);
            if (a.stream_store) __stcs(reinterpret_cast<float4*>(dst), r4);
            else *reinterpret_cast<float4*>(dst) = r4;
        } else {
#pragma unroll
            for (int c = 0; c < 4; c++) dst[c] = y[c];
        }
    } else {
#pragma unroll
        for (int c = 0; c < 4; c++)
            if (e0 + c < a.n_items) dst[c] = f(v[c]);
    }
}

template <typename F, int K0, int K1, int K2>
__global__ void __launch_bounds__(kEwThreads) ew_period_kernel(const PerArgs<F::NIN> a, const F f) {
    constexpr int NIN = F::NIN;
    constexpr int KS[3] = {K0, K1, K2};
    __shared__ __align__(16) uint32_t tbl[NIN][4][kPerMax + 4];
    const uint32_t P = a.period;
    for (uint32_t m = threadIdx.x; m < P + 3; m += kEwThreads) {
        uint32_t rem = m >= P ? m - P : m;
        uint32_t off[NIN];
#pragma unroll
        for (int i = 0; i < NIN; i++) off[i] = 0;
        for (int d = 0; d < a.n_inner; d++) {
            uint32_t idx;
            if (d == a.n_inner - 1) {
                idx = rem;
            } else {
                const uint32_t qq = fdiv(rem, a.in_div[d]);
                idx = rem - qq * a.in_ext[d];
                rem = qq;
            }
#pragma unroll
            for (int i = 0; i < NIN; i++) off[i] += idx * a.in_stride[i][d];
        }
#pragma unroll
        for (int i = 0; i < NIN; i++) {
            const int kind = KS[i] == 3 ? (int)a.kind[i] : KS[i];
            if (kind == 1 || kind == 2) {
                const uint32_t val = kind == 1 ? __float_as_uint(__ldg(a.in[i] + off[i])) : off[i] + (m >= P ? a.out_stride[i][0] : 0u);
#pragma unroll
                for (uint32_t s = 0; s < 4; s++)
                    if (m >= s) tbl[i][s][m - s] = val;
            }
        }
    }
    __syncthreads();
    const uint32_t tbl_base = (uint32_t)__cvta_generic_to_shared(&tbl[0][0][0]);
    const uint32_t n4 = a.n_items / 4;  // whole groups; the ragged one (if any) goes to one thread afterwards
    const uint32_t step = gridDim.x * (uint32_t)kEwThreads;
    // two groups per trip, all gathers of both issued before the first use
    for (uint32_t g = blockIdx.x * (uint32_t)kEwThreads + threadIdx.x; g < n4; g += 2 * step) {
        float v0[4][NIN], v1[4][NIN];
        const bool second = g + step < n4;
        per_load<F, K0, K1, K2, false>(a, tbl_base, 4 * g, v0);
        if (second) per_load<F, K0, K1, K2, false>(a, tbl_base, 4 * (g + step), v1);
        per_store<F, false>(a, f, 4 * g, v0);
        if (second) per_store<F, false>(a, f, 4 * (g + step), v1);
    }
    if ((a.n_items & 3u) && blockIdx.x == 0 && threadIdx.x == 0) {
        float v[4][NIN];
        per_load<F, K0, K1, K2, true>(a, tbl_base, 4 * n4, v);
        per_store<F, true>(a, f, 4 * n4, v);
    }
}

// ---- tiled transpose kernel ----------------------------------------------------------------------
template <int NIN>
struct TileArgs {
    float* out;
    const float* in[NIN];
    uint32_t ext_q, ext_p;  // q: output's inner axis; p: the axis some operand is unit-stride along
    uint32_t tiles_q, tiles_p;
    FastDiv div_tq, div_tp;
    long long out_sp;  // output stride along p (output stride along q is 1)
    long long sq[NIN], sp[NIN];
    uint32_t via_smem;  // bit i: operand i is read p-major and transposed through shared memory
    int nbatch;         // remaining axes, innermost first
    uint32_t bext[kMaxDims];
    FastDiv bdiv[kMaxDims];
    long long bstride_out[kMaxDims];
    long long bstride[NIN][kMaxDims];
    unsigned char stream_store;
    // rectangular tiles (ew_tile_rect): tile extents and the shared-memory pitch
    uint32_t tile_q, tile_p, pitch;
    FastDiv div_tile_q, div_tile_p;
};

constexpr int kTile = 32;  // extents below this use the rectangular / interleave plans

// Scalar (4-byte) transpose tile for extents that are not multiples of 4: 64x64 with 16 loads in flight
// per thread (32x32 for three-operand functors, to stay inside the 48 KB of static shared memory).
template <int NIN>
constexpr int scalar_tile_dim() { return NIN <= 2 ? 64 : 32; }

template <typename F>
__global__ void __launch_bounds__(256) ew_tile_kernel(const TileArgs<F::NIN> a, const F f) {
    constexpr int NIN = F::NIN;
    constexpr int kTile = scalar_tile_dim<NIN>(), kTileRows = 256 / kTile;
    __shared__ float tile[NIN][kTile][kTile + 1];
    uint32_t t = blockIdx.x;
    uint32_t tq = t;
    uint32_t r = fdiv(t, a.div_tq);
    tq = t - r * a.tiles_q;
    uint32_t tp = r;
    uint32_t b = fdiv(r, a.div_tp);
    tp = r - b * a.tiles_p;
    long long obase = 0, ibase[NIN];
#pragma unroll
    for (int i = 0; i < NIN; i++) ibase[i] = 0;
    for (int d = 0; d < a.nbatch; d++) {
        uint32_t idx;
        if (d == a.nbatch - 1) {
            idx = b;
        } else {
            uint32_t q = fdiv(b, a.bdiv[d]);
            idx = b - q * a.bext[d];
            b = q;
        }
        obase += (long long)idx * a.bstride_out[d];
#pragma unroll
        for (int i = 0; i < NIN; i++) ibase[i] += (long long)idx * a.bstride[i][d];
    }
    const uint32_t q0 = tq * kTile, p0 = tp * kTile;
    const int tx = threadIdx.x, ty = threadIdx.y;
    constexpr int kPer = kTile / kTileRows;  // elements per thread
    // Interior tiles (all but the last row / column of tiles) run without a bounds check and with one pointer per
    // operand that is stepped by a constant: ncu had counted 37 instructions per 32 elements on the guarded form
    // (two compares and a 64-bit multiply-add per element), which left odd-extent transposes issue-bound at 4.7 TB/s.
    const bool interior = q0 + kTile <= a.ext_q && p0 + kTile <= a.ext_p;
    float direct[NIN][kPer];
    if (interior) {
        // stage p-major operands: threadIdx.x runs along p (their unit-stride axis)
#pragma unroll
        for (int i = 0; i < NIN; i++) {
            if (a.via_smem >> i & 1) {
                const char* src = reinterpret_cast<const char*>(a.in[i] + ibase[i] + (long long)(q0 + ty) * a.sq[i] + (p0 + tx));
                const long long step = (long long)kTileRows * a.sq[i] * 4;  // bytes: one 64-bit add per load
                float v[kPer];
#pragma unroll
                for (int j = 0; j < kPer; j++) {
                    v[j] = __ldcs(reinterpret_cast<const float*>(src));
                    src += step;
                }
#pragma unroll
                for (int j = 0; j < kPer; j++) tile[i][ty + j * kTileRows][tx] = v[j];
            }
        }
        // q-major operands are fetched before the barrier so their latency overlaps the staging above
#pragma unroll
        for (int i = 0; i < NIN; i++) {
            if (!(a.via_smem >> i & 1)) {
                const char* src = reinterpret_cast<const char*>(a.in[i] + ibase[i] + (long long)(q0 + tx) * a.sq[i] + (long long)(p0 + ty) * a.sp[i]);
                const long long step = (long long)kTileRows * a.sp[i] * 4;
#pragma unroll
                for (int j = 0; j < kPer; j++) {
                    direct[i][j] = __ldg(reinterpret_cast<const float*>(src));
                    src += step;
                }
            }
        }
        __syncthreads();
        char* dst = reinterpret_cast<char*>(a.out + obase + (q0 + tx) + (long long)(p0 + ty) * a.out_sp);
        const long long ostep = (long long)kTileRows * a.out_sp * 4;
#pragma unroll
        for (int j = 0; j < kPer; j++) {
            float x[NIN];
#pragma unroll
            for (int i = 0; i < NIN; i++) {
                if (a.via_smem >> i & 1) x[i] = tile[i][tx][ty + j * kTileRows];
                else x[i] = direct[i][j];
            }
            const float y = f(x);
            if (a.stream_store) __stcs(reinterpret_cast<float*>(dst), y);
            else *reinterpret_cast<float*>(dst) = y;
            dst += ostep;
        }
        return;
    }
    // edge tiles: every access guarded
#pragma unroll
    for (int i = 0; i < NIN; i++) {
        if (a.via_smem >> i & 1) {
            const uint32_t p = p0 + tx;
#pragma unroll
            for (int j = 0; j < kTile; j += kTileRows) {
                const uint32_t q = q0 + ty + j;
                if (q < a.ext_q && p < a.ext_p)
                    tile[i][ty + j][tx] = __ldcs(a.in[i] + ibase[i] + (long long)q * a.sq[i] + (long long)p * a.sp[i]);
            }
        }
    }
    const uint32_t q = q0 + tx;
#pragma unroll
    for (int i = 0; i < NIN; i++) {
        if (!(a.via_smem >> i & 1)) {
#pragma unroll
            for (int j = 0; j < kTile; j += kTileRows) {
                const uint32_t p = p0 + ty + j;
                if (q < a.ext_q && p < a.ext_p)
                    direct[i][j / kTileRows] = __ldg(a.in[i] + ibase[i] + (long long)q * a.sq[i] + (long long)p * a.sp[i]);
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kTile; j += kTileRows) {
        const uint32_t p = p0 + ty + j;
        if (q < a.ext_q && p < a.ext_p) {
            float x[NIN];
#pragma unroll
            for (int i = 0; i < NIN; i++) {
                if (a.via_smem >> i & 1) x[i] = tile[i][tx][ty + j];
                else x[i] = direct[i][j / kTileRows];
            }
            float* dst = a.out + obase + q + (long long)p * a.out_sp;
            const float y = f(x);
            if (a.stream_store) __stcs(dst, y);
            else *dst = y;
        }
    }
}


// ---- rectangular tile kernel: transposes where one of the two extents is small ----------------------------
// [N,3] <-> [3,N] style layouts (interleaved <-> planar): a 32x32 tile would leave most lanes idle on the
// short axis. Here the tile is tile_q x tile_p (<= 1024 elements, the short axis taken whole), threads
// sweep it linearly: p-fastest while loading the p-major operands (coalesced reads), q-fastest while
// emitting (coalesced writes); the odd shared-memory pitch keeps both sweeps bank-conflict free.
constexpr int kRectElems = 1024;

template <typename F>
__global__ void __launch_bounds__(256) ew_tile_rect_kernel(const TileArgs<F::NIN> a, const F f) {
    constexpr int NIN = F::NIN;
    __shared__ float tile[NIN][kRectElems + kRectElems / 2];
    uint32_t t = blockIdx.x;
    uint32_t r = fdiv(t, a.div_tq);
    const uint32_t tq = t - r * a.tiles_q;
    uint32_t b = fdiv(r, a.div_tp);
    const uint32_t tp = r - b * a.tiles_p;
    long long obase = 0, ibase[NIN];
#pragma unroll
    for (int i = 0; i < NIN; i++) ibase[i] = 0;
    for (int d = 0; d < a.nbatch; d++) {
        uint32_t idx;
        if (d == a.nbatch - 1) {
            idx = b;
        } else {
            uint32_t q = fdiv(b, a.bdiv[d]);
            idx = b - q * a.bext[d];
            b = q;
        }
        obase += (long long)idx * a.bstride_out[d];
#pragma unroll
        for (int i = 0; i < NIN; i++) ibase[i] += (long long)idx * a.bstride[i][d];
    }
    const uint32_t q0 = tq * a.tile_q, p0 = tp * a.tile_p;
    const uint32_t n_tile = a.tile_q * a.tile_p;
    for (uint32_t e = threadIdx.x; e < n_tile; e += 256) {  // p fastest
        const uint32_t ql = fdiv(e, a.div_tile_p), pl = e - ql * a.tile_p;
        const uint32_t q = q0 + ql, p = p0 + pl;
        if (q < a.ext_q && p < a.ext_p) {
#pragma unroll
            for (int i = 0; i < NIN; i++)
                if (a.via_smem >> i & 1)
                    tile[i][ql * a.pitch + pl] = __ldcs(a.in[i] + ibase[i] + (long long)q * a.sq[i] + (long long)p * a.sp[i]);
        }
    }
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < n_tile; e += 256) {  // q fastest
        const uint32_t pl = fdiv(e, a.div_tile_q), ql = e - pl * a.tile_q;
        const uint32_t q = q0 + ql, p = p0 + pl;
        if (q < a.ext_q && p < a.ext_p) {
            float x[NIN];
#pragma unroll
            for (int i = 0; i < NIN; i++) {
                if (a.via_smem >> i & 1) x[i] = tile[i][ql * a.pitch + pl];
                else x[i] = __ldg(a.in[i] + ibase[i] + (long long)q * a.sq[i] + (long long)p * a.sp[i]);
            }
            float* dst = a.out + obase + q + (long long)p * a.out_sp;
            const float y = f(x);
            if (a.stream_store) __stcs(dst, y);
            else *dst = y;
        }
    }
}

// ---- packed interleave kernels: [N, C] <-> [C, N] with a short, densely packed axis C ------------------------
// The layouts of RGB pixels, xyz points or complex pairs. Because the interleaved side is one
// contiguous stream, it is moved with 128-bit accesses; the planar side is moved with 128-bit accesses
// along n; only the shared-memory side of the shuffle is scalar. A CTA handles `chunk` consecutive n.
constexpr int kIlvFloats = 4096;  // shared-memory floats per staged operand (chunk * C <= 4096)

template <int NIN>
struct IlvArgs {
    float* out;
    const float* in[NIN];
    uint32_t C, N, chunk, chunks;  // short extent, long extent, n per CTA, CTAs per batch item
    FastDiv div_c, div_chunks;
    long long plane_in;             // staged operand: stride between planes (planar input) -- unused when interleaved
    long long plane_out;            // output: stride between planes (planar output)
    uint32_t staged;                // index of the operand that goes through shared memory
    unsigned char congruent[NIN];   // other operands: 1 = same layout as the output, 0 = one broadcast scalar
    int nbatch;
    uint32_t bext[kMaxDims];
    FastDiv bdiv[kMaxDims];
    long long bstride_out[kMaxDims];
    long long bstride[NIN][kMaxDims];
    unsigned char stream_store;
};

template <typename F, bool PLANAR_OUT>
__global__ void __launch_bounds__(256) ew_interleave_kernel(const IlvArgs<F::NIN> a, const F f) {
    constexpr int NIN = F::NIN;
    __shared__ __align__(16) float plane[kIlvFloats + 4 * 32];
    uint32_t b = fdiv(blockIdx.x, a.div_chunks);
    const uint32_t n0 = (blockIdx.x - b * a.chunks) * a.chunk;
    long long obase = 0, ibase[NIN];
#pragma unroll
    for (int i = 0; i < NIN; i++) ibase[i] = 0;
    for (int d = 0; d < a.nbatch; d++) {
        uint32_t idx;
        if (d == a.nbatch - 1) {
            idx = b;
        } else {
            uint32_t q = fdiv(b, a.bdiv[d]);
            idx = b - q * a.bext[d];
            b = q;
        }
        obase += (long long)idx * a.bstride_out[d];
#pragma unroll
        for (int i = 0; i < NIN; i++) ibase[i] += (long long)idx * a.bstride[i][d];
    }
    const uint32_t len = min(a.chunk, a.N - n0);  // n handled here (a multiple of 4)
    const uint32_t pitch = a.chunk + 4;           // floats between planes in shared memory (16-byte aligned)
    const uint32_t stream4 = len * a.C / 4;       // float4 items of the interleaved stream
    const float* xin = a.in[a.staged] + ibase[a.staged];
    auto apply4 = [&](float4 x, long long other_off) -> float4 {
        float4 o[NIN];
#pragma unroll
        for (int i = 0; i < NIN; i++) {
            if ((uint32_t)i == a.staged) {
                o[i] = x;
            } else if (a.congruent[i]) {
                o[i] = __ldcs(reinterpret_cast<const float4*>(a.in[i] + ibase[i] + other_off));
            } else {
                const float sv = __ldg(a.in[i]);
                o[i] = make_float4(sv, sv, sv, sv);
            }
        }
        float4 y;
        float e[NIN];
#pragma unroll
        for (int i = 0; i < NIN; i++) e[i] = o[i].x;
        y.x = f(e);
#pragma unroll
        for (int i = 0; i < NIN; i++) e[i] = o[i].y;
        y.y = f(e);
#pragma unroll
        for (int i = 0; i < NIN; i++) e[i] = o[i].z;
        y.z = f(e);
#pragma unroll
        for (int i = 0; i < NIN; i++) e[i] = o[i].w;
        y.w = f(e);
        return y;
    };
    if constexpr (PLANAR_OUT) {
        // interleaved in -> planes in shared memory
        const float4* src = reinterpret_cast<const float4*>(xin + (long long)n0 * a.C);
        for (uint32_t v = threadIdx.x; v < stream4; v += 256) {
            const float4 x = __ldcs(src + v);
            uint32_t n = fdiv(4 * v, a.div_c), c = 4 * v - n * a.C;
            const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                plane[c * pitch + n] = xs[j];
                if (++c == a.C) c = 0, n++;
            }
        }
        __syncthreads();
        const uint32_t per_plane = len / 4;
        for (uint32_t v = threadIdx.x; v < per_plane * a.C; v += 256) {
            const uint32_t c = v / per_plane, t4 = v - c * per_plane;
            const float4 x = *reinterpret_cast<const float4*>(plane + c * pitch + 4 * t4);
            const long long off = (long long)c * a.plane_out + n0 + 4 * t4;
            float4* dst = reinterpret_cast<float4*>(a.out + obase + off);
            const float4 y = apply4(x, off);
            if (a.stream_store) __stcs(dst, y);
            else *dst = y;
        }
    } else {
        // planar in -> planes in shared memory -> interleaved stream out
        const uint32_t per_plane = len / 4;
        for (uint32_t v = threadIdx.x; v < per_plane * a.C; v += 256) {
            const uint32_t c = v / per_plane, t4 = v - c * per_plane;
            const float4 x = __ldcs(reinterpret_cast<const float4*>(xin + (long long)c * a.plane_in + n0 + 4 * t4));
            *reinterpret_cast<float4*>(plane + c * pitch + 4 * t4) = x;
        }
        __syncthreads();
        for (uint32_t v = threadIdx.x; v < stream4; v += 256) {
            uint32_t n = fdiv(4 * v, a.div_c), c = 4 * v - n * a.C;
            float xs[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                xs[j] = plane[c * pitch + n];
                if (++c == a.C) c = 0, n++;
            }
            const long long off = (long long)n0 * a.C + 4 * v;
            float4* dst = reinterpret_cast<float4*>(a.out + obase + off);
            const float4 y = apply4(make_float4(xs[0], xs[1], xs[2], xs[3]), off);
            if (a.stream_store) __stcs(dst, y);
            else *dst = y;
        }
    }
}

// ---- 64x64 float4 transpose kernel (fast path of ew_tile) ---------------------------------------------
// Every global access is 128-bit. A thread loads a 4x4 block (4 rows along q, one float4 along p),
// transposes it in registers and parks the 4 resulting float4 in shared memory; after the barrier
// threads re-read float4 with lanes running along q, so the global store is coalesced too.
// Shared memory is addressed in float4 units, 16 units per tile row, with unit' = unit ^ ((row>>2)&7):
// both the store phase (8 lanes of a quarter-warp hit rows 4 apart) and the load phase (8 lanes hit
// consecutive units of one row) then cover all 32 banks exactly once.
// The tile is TQ (along q) x TP (along p) with TQ*TP = 4096, TQ in {32, 64, 128}: a wider TP makes the
// p-major READ runs longer (TP*4 bytes per row), a wider TQ the q-major WRITE runs.
template <typename F, int TQ>
__global__ void __launch_bounds__(256) ew_tile64_kernel(const TileArgs<F::NIN> a, const F f) {
    constexpr int NIN = F::NIN;
    constexpr int TP = 4096 / TQ, UQ = TQ / 4, P4 = TP / 4, ROWS_PER_PASS = 256 / UQ;
    static_assert(UQ >= 8 && P4 >= 8, "swizzle needs at least 8 float4 units per row and 8 p-groups");
    __shared__ float4 tile[NIN][TP * UQ];
    // (A grouped 8x8 CTA order was measured and gave no gain over this q-fastest order.)
    uint32_t t = blockIdx.x;
    uint32_t r = fdiv(t, a.div_tq);
    const uint32_t tq = t - r * a.tiles_q;
    uint32_t b = fdiv(r, a.div_tp);
    const uint32_t tp = r - b * a.tiles_p;
    long long obase = 0, ibase[NIN];
#pragma unroll
    for (int i = 0; i < NIN; i++) ibase[i] = 0;
    for (int d = 0; d < a.nbatch; d++) {
        uint32_t idx;
        if (d == a.nbatch - 1) {
            idx = b;
        } else {
            uint32_t q = fdiv(b, a.bdiv[d]);
            idx = b - q * a.bext[d];
            b = q;
        }
        obase += (long long)idx * a.bstride_out[d];
#pragma unroll
        for (int i = 0; i < NIN; i++) ibase[i] += (long long)idx * a.bstride[i][d];
    }
    const uint32_t q0 = tq * TQ, p0 = tp * TP;
    const uint32_t lo = threadIdx.x % P4, hi = threadIdx.x / P4;
    // stage: lo indexes float4 along p, hi indexes 4-row groups along q
#pragma unroll
    for (int i = 0; i < NIN; i++) {
        if (a.via_smem >> i & 1) {
            const uint32_t p = p0 + 4 * lo;
            float4 v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t q = q0 + 4 * hi + k;
                v[k] = (q < a.ext_q && p < a.ext_p)
                           ? __ldcs(reinterpret_cast<const float4*>(a.in[i] + ibase[i] + (long long)q * a.sq[i] + p))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            const uint32_t sw = hi ^ (lo & 7);
            tile[i][(4 * lo + 0) * UQ + sw] = make_float4(v[0].x, v[1].x, v[2].x, v[3].x);
            tile[i][(4 * lo + 1) * UQ + sw] = make_float4(v[0].y, v[1].y, v[2].y, v[3].y);
            tile[i][(4 * lo + 2) * UQ + sw] = make_float4(v[0].z, v[1].z, v[2].z, v[3].z);
            tile[i][(4 * lo + 3) * UQ + sw] = make_float4(v[0].w, v[1].w, v[2].w, v[3].w);
        }
    }
    // operands that are already q-major are fetched now, so their latency overlaps the staging above
    const uint32_t eu = threadIdx.x % UQ, er = threadIdx.x / UQ;  // emit role: float4 unit along q, row along p
    const uint32_t q = q0 + 4 * eu;
    float4 direct[NIN][4];
#pragma unroll
    for (int i = 0; i < NIN; i++) {
        if (!(a.via_smem >> i & 1)) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t p = p0 + er + ROWS_PER_PASS * j;
                if (q < a.ext_q && p < a.ext_p) {
                    const float* src = a.in[i] + ibase[i] + (long long)q * a.sq[i] + (long long)p * a.sp[i];
                    if (a.sq[i] == 0) {
                        const float sv = __ldg(src);
                        direct[i][j] = make_float4(sv, sv, sv, sv);
                    } else {
                        direct[i][j] = __ldcs(reinterpret_cast<const float4*>(src));
                    }
                }
            }
        }
    }
    __syncthreads();
    // emit: eu indexes float4 along q, rows (p) = er + ROWS_PER_PASS*j
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t row = er + ROWS_PER_PASS * j;
        const uint32_t p = p0 + row;
        if (q < a.ext_q && p < a.ext_p) {
            float4 x[NIN];
#pragma unroll
            for (int i = 0; i < NIN; i++) {
                if (a.via_smem >> i & 1) x[i] = tile[i][row * UQ + (eu ^ ((row >> 2) & 7))];
                else x[i] = direct[i][j];
            }
            float4 y;
            float e[NIN];
#pragma unroll
            for (int i = 0; i < NIN; i++) e[i] = x[i].x;
            y.x = f(e);
#pragma unroll
            for (int i = 0; i < NIN; i++) e[i] = x[i].y;
            y.y = f(e);
#pragma unroll
            for (int i = 0; i < NIN; i++) e[i] = x[i].z;
            y.z = f(e);
#pragma unroll
            for (int i = 0; i < NIN; i++) e[i] = x[i].w;
            y.w = f(e);
            float4* dst = reinterpret_cast<float4*>(a.out + obase + q + (long long)p * a.out_sp);
            if (a.stream_store) __stcs(dst, y);
            else *dst = y;
        }
    }
}

// ---- register-tiled outer broadcast kernel ("outer sum": [I,1,W] o [1,J,W]) ---------------------------
// When operand 0 is broadcast along axis d1 and operand 1 along axis d2, the plain N-d kernel issues
// two loads per output and becomes L2-bandwidth bound (both operands live in L2; measured 4.9 TB/s
// on config C3). Here a thread owns a T x T block of (d1, d2) for one inner float4: it loads T
// vectors of each operand once and produces T*T outputs, i.e. 2/T loads per output.
constexpr int kOT = 4;

struct OuterArgs {
    float* out;
    const float* x;  // operand 0: stride 0 along d1, varies along d2
    const float* y;  // operand 1: stride 0 along d2, varies along d1
    uint32_t n_items;
    int ndim;  // iteration axes: [inner float4, ..., d1 blocks, ..., d2 blocks, ...]
    uint32_t extent[kMaxDims];
    FastDiv div[kMaxDims];
    long long xs[kMaxDims], ys[kMaxDims], os[kMaxDims];  // element strides per unit of each iteration axis
    int d1, d2;
    uint32_t e1, e2;
    long long x_s2, y_s1, o_s1, o_s2;
    unsigned char x_inner_bcast, y_inner_bcast, stream_store;
};

template <typename F>
__global__ void __launch_bounds__(256) ew_outer_kernel(const OuterArgs a, const F f) {
    const uint32_t item = blockIdx.x * 256u + threadIdx.x;
    if (item >= a.n_items) return;
    long long xo = 0, yo = 0, oo = 0;
    uint32_t rem = item, i1 = 0, i2 = 0;
#pragma unroll
    for (int d = 0; d < kMaxDims; d++) {
        uint32_t idx;
        const bool last = (d == a.ndim - 1);
        if (last) {
            idx = rem;
        } else {
            const uint32_t q = fdiv(rem, a.div[d]);
            idx = rem - q * a.extent[d];
            rem = q;
        }
        if (d == a.d1) i1 = idx * kOT;
        if (d == a.d2) i2 = idx * kOT;
        xo += (long long)idx * a.xs[d];
        yo += (long long)idx * a.ys[d];
        oo += (long long)idx * a.os[d];
        if (last) break;
    }
    float4 xv[kOT], yv[kOT];
#pragma unroll
    for (int b = 0; b < kOT; b++) {
        if (i2 + b < a.e2) {
            const float* p = a.x + xo + b * a.x_s2;
            if (a.x_inner_bcast) {
                const float s = __ldg(p);
                xv[b] = make_float4(s, s, s, s);
            } else {
                xv[b] = __ldg(reinterpret_cast<const float4*>(p));
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kOT; k++) {
        if (i1 + k < a.e1) {
            const float* p = a.y + yo + k * a.y_s1;
            if (a.y_inner_bcast) {
                const float s = __ldg(p);
                yv[k] = make_float4(s, s, s, s);
            } else {
                yv[k] = __ldg(reinterpret_cast<const float4*>(p));
            }
        }
    }
#pragma unroll
    for (int b = 0; b < kOT; b++) {
#pragma unroll
        for (int k = 0; k < kOT; k++) {
            if (i2 + b < a.e2 && i1 + k < a.e1) {
                float4 r;
                float e[2];
                e[0] = xv[b].x, e[1] = yv[k].x;
                r.x = f(e);
                e[0] = xv[b].y, e[1] = yv[k].y;
                r.y = f(e);
                e[0] = xv[b].z, e[1] = yv[k].z;
                r.z = f(e);
                e[0] = xv[b].w, e[1] = yv[k].w;
                r.w = f(e);
                float4* dst = reinterpret_cast<float4*>(a.out + oo + k * a.o_s1 + b * a.o_s2);
                if (a.stream_store) __stcs(dst, r);
                else *dst = r;
            }
        }
    }
}

// ---- host-side launch plans ------------------------------------------------------------------------

// Period plan (ew_period_kernel): 0 = not applicable, 1 = launched, -1 = error.
template <typename F>
static int try_launch_period(const F& f, const Folded& fo, const float* const* in, float* out, uint64_t n_items) {
    constexpr int NIN = F::NIN;
    uint64_t P = 1;
    int g = 0;
    while (g < fo.ndim && P * fo.extent[g] <= (uint64_t)kPerMax) P *= fo.extent[g++];
    if (g == fo.ndim) return 0;
    uint64_t split = 1;
    for (uint64_t c = (uint64_t)kPerMax / P; c >= 2; c--)
        if (fo.extent[g] % c == 0) {
            split = c;
            break;
        }
    P *= split;
    if (P < 8) return 0;
    PerArgs<NIN> a;
    memset(&a, 0, sizeof a);
    a.out = out;
    a.n_items = (uint32_t)n_items;
    a.period = (uint32_t)P;
    a.period_div = make_fastdiv((uint32_t)P);
    a.n_inner = g + (split > 1 ? 1 : 0);
    a.n_outer = fo.ndim - g;
    for (int d = 0; d < g; d++) a.in_ext[d] = (uint32_t)fo.extent[d];
    if (split > 1) a.in_ext[g] = (uint32_t)split;
    for (int d = 0; d < a.n_inner; d++) a.in_div[d] = make_fastdiv(a.in_ext[d]);
    for (int d = g; d < fo.ndim; d++) {
        a.out_ext[d - g] = (uint32_t)(d == g ? fo.extent[d] / split : fo.extent[d]);
        a.out_div[d - g] = make_fastdiv(a.out_ext[d - g]);
    }
    for (int i = 0; i < NIN; i++) {
        a.in[i] = in[i];
        uint64_t span = 0;
        bool congruent = (uintptr_t)in[i] % 16 == 0, periodic = true;
        int64_t run = 1;
        for (int d = 0; d < fo.ndim; d++) {
            const int64_t st = fo.in_stride[i][d];
            if (st < 0) return 0;
            span += (fo.extent[d] - 1) * (uint64_t)st;
            if (fo.extent[d] > 1 && st != run) congruent = false;
            if (d >= g && st != 0) periodic = false;
            run *= (int64_t)fo.extent[d];
        }
        if (span >= (1ull << 32)) return 0;
        for (int d = 0; d < g; d++) a.in_stride[i][d] = (uint32_t)fo.in_stride[i][d];
        if (split > 1) a.in_stride[i][g] = (uint32_t)fo.in_stride[i][g];
        for (int d = g; d < fo.ndim; d++)
            a.out_stride[i][d - g] = (uint32_t)(fo.in_stride[i][d] * (int64_t)(d == g ? split : 1));
        a.kind[i] = congruent ? 0 : (periodic ? 1 : 2);
        if (a.kind[i] == 2 && (uintptr_t)in[i] % 16 == 0) {  // (a period that fits the table is cheaper from shared memory: kind 1)
            // unit-stride over a leading run of axes and stride 0 above it (what broadcasting a contiguous array gives)
            uint64_t len = 1;
            int d2 = 0;
            while (d2 < fo.ndim && (fo.in_stride[i][d2] == (int64_t)len || fo.extent[d2] == 1)) len *= fo.extent[d2++];
            bool rest_zero = d2 > 0;
            for (int d = d2; d < fo.ndim; d++) rest_zero = rest_zero && fo.in_stride[i][d] == 0;
            if (rest_zero && len % 4 == 0 && len < (1ull << 31)) {
                a.kind[i] = 4;
                a.blk_len[i] = (uint32_t)len;
                a.blk_len_div[i] = make_fastdiv((uint32_t)len);
            }
        }
    }
    a.out_aligned = (uintptr_t)out % 16 == 0;
    a.stream_store = n_items * sizeof(float) >= (uint64_t)rt().stream_store_min_bytes;
    const uint64_t n4 = (n_items + 3) / 4;
    const uint64_t want = (n4 + 2 * kEwThreads - 1) / (2 * kEwThreads);  // two groups per thread and trip
    auto launch = [&](auto kernel) {
        static int resident = 0;  // persistent CTAs: as many as fit one SM (register-bound, 6)
        if (resident == 0) {
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kernel, kEwThreads, 0) != cudaSuccess || resident < 1) resident = 1;
        }
        const uint64_t cap = (uint64_t)kNumSMs * (uint64_t)resident;
        kernel<<<(unsigned)(want < cap ? want : cap), kEwThreads, 0, rt().stream>>>(a, f);
    };
    const int k0 = a.kind[0], k1 = NIN > 1 ? a.kind[1] : 3;
    if (NIN == 1 && k0 == 2) launch(ew_period_kernel<F, 2, 3, 3>);
    else if (NIN == 2 && k0 == 2 && k1 == 1) launch(ew_period_kernel<F, 2, 1, 3>);
    else if (NIN == 2 && k0 == 1 && k1 == 2) launch(ew_period_kernel<F, 1, 2, 3>);
    else if (NIN == 2 && k0 == 2 && k1 == 2) launch(ew_period_kernel<F, 2, 2, 3>);
    else if (NIN == 2 && k0 == 2 && k1 == 4) launch(ew_period_kernel<F, 2, 4, 3>);
    else if (NIN == 2 && k0 == 4 && k1 == 2) launch(ew_period_kernel<F, 4, 2, 3>);
    else launch(ew_period_kernel<F, 3, 3, 3>);
    note_launch("ew_period");
    return check_launch("ew_period") ? 1 : -1;
}
template <int VEC, typename F>
static bool launch_nd_one(const F& f, const Folded& fo, const float* const* in, float* out, uint64_t n_items) {
    constexpr int NIN = F::NIN;
    NdArgs<NIN> a;
    memset(&a, 0, sizeof a);
    a.out = out;
    a.n_items = (uint32_t)n_items;
    a.ndim = fo.ndim;
    for (int d = 0; d < fo.ndim; d++) {
        uint64_t e = fo.extent[d];
        if (d == 0) e /= VEC;
        a.extent[d] = (uint32_t)e;
        a.div[d] = make_fastdiv((uint32_t)e);
    }
    uint64_t out_bytes = n_items * VEC * sizeof(float);
    a.stream_store = out_bytes >= (uint64_t)rt().stream_store_min_bytes;
    a.load_policy = (unsigned char)rt().load_policy;
    for (int i = 0; i < NIN; i++) {
        a.in[i] = in[i];
        bool reuse = false;
        for (int d = 0; d < fo.ndim; d++) {
            a.stride[i][d] = fo.in_stride[i][d] * (d == 0 ? VEC : 1);
            if (fo.in_stride[i][d] == 0 && fo.extent[d] > 1) reuse = true;
        }
        a.inner_bcast[i] = fo.in_stride[i][0] == 0;
        a.reuse[i] = reuse;
    }
    cudaStream_t s = rt().stream;
    if constexpr (VEC == 4) {
        // (unroll 1/2/4/8 were swept on C2 and are within 1.5 % of each other; 4 is kept)
        constexpr int kUnroll = 4;
        const unsigned blocks = (unsigned)((n_items + (uint64_t)kEwThreads * kUnroll - 1) / ((uint64_t)kEwThreads * kUnroll));
        ew_nd_kernel<4, kUnroll, F><<<blocks, kEwThreads, 0, s>>>(a, f);
    } else if (fo.ndim == 1) {
        // a flat stream that is merely misaligned (or strided): one element per lane is already perfectly
        // coalesced and needs no index arithmetic (6.57 TB/s on a 4-byte-shifted a+b, vs 6.15 with the run kernel)
        constexpr int kUnroll = 4;
        const unsigned blocks = (unsigned)((n_items + (uint64_t)kEwThreads * kUnroll - 1) / ((uint64_t)kEwThreads * kUnroll));
        ew_nd_kernel<1, kUnroll, F><<<blocks, kEwThreads, 0, s>>>(a, f);
    } else {
        a.out_aligned = (uintptr_t)out % 16 == 0;
        for (int i = 0; i < NIN; i++) {
            // the run of axes [d1, d2) over which the operand is unit-stride-contiguous; stride 0 everywhere else?
            int d1 = 0;
            while (d1 < fo.ndim && fo.in_stride[i][d1] == 0) d1++;
            uint64_t inner = 1, len = 1;
            for (int d = 0; d < d1; d++) inner *= fo.extent[d];
            bool block = true;
            int d2 = d1;
            int64_t run = 1;
            while (d2 < fo.ndim && fo.in_stride[i][d2] == run) {
                run *= (int64_t)fo.extent[d2];
                len *= fo.extent[d2];
                d2++;
            }
            for (int d = d2; d < fo.ndim; d++) block = block && fo.in_stride[i][d] == 0;
            if (d1 == fo.ndim) inner = 1, len = 1;  // a single element broadcast everywhere
            block = block && inner < (1ull << 31) && len < (1ull << 31);
            const bool congruent = block && d1 == 0 && d2 == fo.ndim;
            a.mode[i] = !block ? 2 : (congruent && (uintptr_t)in[i] % 16 == 0 ? 0 : 1);
            if (a.mode[i] == 1 && inner == 1 && len % kRun == 0 && (uintptr_t)in[i] % 16 == 0) a.mode[i] = 4;  // periodic, vectorisable
            a.blk_inner[i] = (uint32_t)(block ? inner : 1);
            a.blk_len[i] = (uint32_t)(block ? len : 1);
            if (!block && d1 == 0 && d2 > 0 && d2 < fo.ndim && len < (1ull << 31)) {
                // contiguous over [0, d2), stride 0 over [d2, d3), contiguous again (as if the middle were absent) above
                int d3 = d2;
                uint64_t p2 = len;  // = prod extent[0..d2)
                while (d3 < fo.ndim && fo.in_stride[i][d3] == 0) p2 *= fo.extent[d3++];
                bool middle = d3 < fo.ndim;
                int64_t expect = (int64_t)len;
                for (int d = d3; d < fo.ndim && middle; d++) {
                    middle = fo.in_stride[i][d] == expect;
                    expect *= (int64_t)fo.extent[d];
                }
                if (middle && p2 < (1ull << 31)) {
                    a.mode[i] = 3;
                    a.blk_inner[i] = (uint32_t)len;
                    a.blk_len[i] = (uint32_t)p2;
                }
            }
            a.blk_inner_div[i] = make_fastdiv(a.blk_inner[i]);
            a.blk_len_div[i] = make_fastdiv(a.blk_len[i]);
        }
        bool hard = false;  // some operand is outside ew_run's cheap addressing modes
        for (int i = 0; i < NIN; i++) hard = hard || a.mode[i] == 2 || a.mode[i] == 3;
        if (hard && rt().period_plan && !rt().force_general && fo.extent[0] <= 128) {
            const int per = try_launch_period<F>(f, fo, in, out, n_items);
            if (per != 0) return per > 0;
        }
        const uint64_t threads = (n_items + kRun - 1) / kRun;
        ew_run_kernel<F><<<(unsigned)((threads + kEwThreads - 1) / kEwThreads), kEwThreads, 0, s>>>(a, f);
    }
    note_launch(VEC == 4 ? "ew_nd_vec4" : "ew_nd_scalar");
    return check_launch("ew_nd");
}

// Launches over `fo`, bisecting the outermost axis until a launch holds < 2^31 vector items.
template <int VEC, typename F>
static bool launch_nd(const F& f, Folded fo, const float* const* in_ptrs, float* out) {
    constexpr int NIN = F::NIN;
    uint64_t total = 1;
    for (int d = 0; d < fo.ndim; d++) total *= fo.extent[d];
    uint64_t n_items = total / VEC;
    if (n_items < (1ull << 31)) return launch_nd_one<VEC, F>(f, fo, in_ptrs, out, n_items);
    int top = fo.ndim - 1;
    uint64_t e = fo.extent[top];
    uint64_t a_ext = e / 2;
    if (top == 0) a_ext &= ~(uint64_t)(VEC - 1);
    uint64_t inner = total / e;  // output elements per step of the outermost axis
    Folded lo = fo, hi = fo;
    lo.extent[top] = a_ext;
    hi.extent[top] = e - a_ext;
    const float* in_hi[NIN];
    for (int i = 0; i < NIN; i++) in_hi[i] = in_ptrs[i] + (int64_t)a_ext * fo.in_stride[i][top];
    return launch_nd<VEC, F>(f, lo, in_ptrs, out) && launch_nd<VEC, F>(f, hi, in_hi, out + a_ext * inner);
}

// Packed interleave plan for a TileArgs already filled by launch_tile: 0 = not applicable, 1 = launched,
// -1 = error. Needs exactly one staged operand, densely packed on its interleaved side, every other
// operand either laid out like the output or a single broadcast element, and 4-divisible extents.
template <typename F>
static int try_launch_interleave(const F& f, const Folded& fo, const float* const* in, float* out, int P,
                                 const TileArgs<F::NIN>& t) {
    constexpr int NIN = F::NIN;
    if (!rt().interleave) return 0;
    int staged = -1;
    for (int i = 0; i < NIN; i++)
        if (t.via_smem >> i & 1) {
            if (staged >= 0) return 0;
            staged = i;
        }
    if (staged < 0) return 0;
    const bool planar_out = t.ext_p < t.ext_q;  // short axis p: operand is [n=q, c=p] interleaved, output is [c, n]
    const uint32_t C = planar_out ? t.ext_p : t.ext_q, N = planar_out ? t.ext_q : t.ext_p;
    if (C < 2 || C > 32 || N % 4 != 0 || N < 256) return 0;
    // long long strides of the staged operand: (sq along q, sp == 1 along p)
    if (planar_out) {
        if (t.sq[staged] != (long long)C) return 0;          // interleaved input must be dense: offset = n*C + c
        if (t.out_sp % 4 != 0) return 0;                     // planar output planes 16-byte aligned
    } else {
        if (t.sq[staged] % 4 != 0) return 0;                 // planar input: plane stride, 16-byte aligned
        if (t.out_sp != (long long)C) return 0;              // interleaved output must be dense
    }
    if ((uintptr_t)out % 16 != 0 || (uintptr_t)in[staged] % 16 != 0) return 0;
    IlvArgs<NIN> a;
    memset(&a, 0, sizeof a);
    a.out = out;
    a.C = C;
    a.N = N;
    a.staged = (uint32_t)staged;
    a.plane_in = planar_out ? 0 : t.sq[staged];
    a.plane_out = planar_out ? t.out_sp : 0;
    a.chunk = (uint32_t)(kIlvFloats / C / 32 * 32);
    if (a.chunk > N) a.chunk = N;
    a.chunks = (N + a.chunk - 1) / a.chunk;
    a.div_c = make_fastdiv(C);
    a.div_chunks = make_fastdiv(a.chunks);
    a.nbatch = t.nbatch;
    uint64_t nb = 1;
    for (int d = 0; d < t.nbatch; d++) {
        a.bext[d] = t.bext[d];
        a.bdiv[d] = t.bdiv[d];
        a.bstride_out[d] = t.bstride_out[d];
        if (t.bstride_out[d] % 4 != 0) return 0;
        for (int i = 0; i < NIN; i++) a.bstride[i][d] = t.bstride[i][d];
        if (t.bstride[staged][d] % 4 != 0) return 0;
        nb *= t.bext[d];
    }
    for (int i = 0; i < NIN; i++) {
        a.in[i] = in[i];
        if (i == staged) continue;
        bool all_zero = t.sq[i] == 0 && t.sp[i] == 0, same = t.sq[i] == 1 && t.sp[i] == t.out_sp;
        for (int d = 0; d < t.nbatch; d++) {
            all_zero = all_zero && t.bstride[i][d] == 0;
            same = same && t.bstride[i][d] == t.bstride_out[d];
        }
        if (same && (uintptr_t)in[i] % 16 == 0) a.congruent[i] = 1;
        else if (all_zero) a.congruent[i] = 0;
        else return 0;
    }
    a.stream_store = t.stream_store;
    const uint64_t blocks = (uint64_t)a.chunks * nb;
    if (blocks >= (1ull << 31)) return 0;
    (void)fo, (void)P;
    if (planar_out) ew_interleave_kernel<F, true><<<(unsigned)blocks, 256, 0, rt().stream>>>(a, f);
    else ew_interleave_kernel<F, false><<<(unsigned)blocks, 256, 0, rt().stream>>>(a, f);
    note_launch(planar_out ? "ew_deinterleave" : "ew_interleave");
    return check_launch("ew_interleave") ? 1 : -1;
}

template <typename F>
static bool launch_tile(const F& f, const Folded& fo, const float* const* in, float* out, int P) {
    constexpr int NIN = F::NIN;
    TileArgs<NIN> a;
    memset(&a, 0, sizeof a);
    a.out = out;
    a.ext_q = (uint32_t)fo.extent[0];
    a.ext_p = (uint32_t)fo.extent[P];
    constexpr int kScalarTile = scalar_tile_dim<NIN>();
    a.tiles_q = (a.ext_q + kScalarTile - 1) / kScalarTile;
    a.tiles_p = (a.ext_p + kScalarTile - 1) / kScalarTile;
    a.div_tq = make_fastdiv(a.tiles_q);
    a.div_tp = make_fastdiv(a.tiles_p);
    // contiguous output strides of the folded axes
    long long ostride[kMaxDims];
    long long run = 1;
    uint64_t total = 1;
    for (int d = 0; d < fo.ndim; d++) {
        ostride[d] = run;
        run *= (long long)fo.extent[d];
        total *= fo.extent[d];
    }
    a.out_sp = ostride[P];
    uint64_t nb = 1;
    int k = 0;
    for (int d = 1; d < fo.ndim; d++) {
        if (d == P) continue;
        a.bext[k] = (uint32_t)fo.extent[d];
        a.bdiv[k] = make_fastdiv(a.bext[k]);
        a.bstride_out[k] = ostride[d];
        for (int i = 0; i < NIN; i++) a.bstride[i][k] = fo.in_stride[i][d];
        nb *= fo.extent[d];
        k++;
    }
    a.nbatch = k;
    for (int i = 0; i < NIN; i++) {
        a.in[i] = in[i];
        a.sq[i] = fo.in_stride[i][0];
        a.sp[i] = fo.in_stride[i][P];
        if (a.sp[i] == 1 && a.sq[i] > 1) a.via_smem |= 1u << i;
    }
    a.stream_store = total * sizeof(float) >= (uint64_t)rt().stream_store_min_bytes;
    // 128-bit variant: every extent/stride/pointer that a float4 access touches must be 4-aligned
    const bool short_axis = a.ext_q < (uint32_t)kTile || a.ext_p < (uint32_t)kTile;  // -> rectangular tiles
    bool wide = rt().tile64 && !short_axis && a.ext_q % 4 == 0 && a.ext_p % 4 == 0 && (uintptr_t)out % 16 == 0;
    for (int i = 0; i < NIN && wide; i++) {
        const bool smem = a.via_smem >> i & 1;
        if (!smem && a.sq[i] > 1) wide = false;                 // would need a strided gather
        if ((smem || a.sq[i] == 1) && (uintptr_t)in[i] % 16 != 0) wide = false;
        if (smem && a.sq[i] % 4 != 0) wide = false;
        if (!smem && a.sq[i] == 1 && a.sp[i] % 4 != 0) wide = false;
        for (int d = 0; d < a.nbatch && wide; d++)
            if ((smem || a.sq[i] == 1) && a.bstride[i][d] % 4 != 0) wide = false;
    }
    if (wide) {
        const int tq = rt().tile_tq == 32 || rt().tile_tq == 128 ? (int)rt().tile_tq : 64;
        const int tp = 4096 / tq;
        a.tiles_q = (a.ext_q + tq - 1) / tq;
        a.tiles_p = (a.ext_p + tp - 1) / tp;
        a.div_tq = make_fastdiv(a.tiles_q);
        a.div_tp = make_fastdiv(a.tiles_p);
        uint64_t blocks64 = (uint64_t)a.tiles_q * a.tiles_p * nb;
        if (blocks64 >= (1ull << 31)) return false;
        if (tq == 32) ew_tile64_kernel<F, 32><<<(unsigned)blocks64, 256, 0, rt().stream>>>(a, f);
        else if (tq == 128) ew_tile64_kernel<F, 128><<<(unsigned)blocks64, 256, 0, rt().stream>>>(a, f);
        else ew_tile64_kernel<F, 64><<<(unsigned)blocks64, 256, 0, rt().stream>>>(a, f);
        note_launch("ew_tile64_transpose");
        return check_launch("ew_tile64");
    }
    if (short_axis) {
        const int ilv = try_launch_interleave<F>(f, fo, in, out, P, a);
        if (ilv != 0) return ilv > 0;
        // one extent is shorter than a 32x32 tile: take it whole and stretch the tile along the other
        a.tile_p = a.ext_p < (uint32_t)kTile ? a.ext_p : (uint32_t)kRectElems / (a.ext_q < (uint32_t)kTile ? a.ext_q : (uint32_t)kTile);
        a.tile_q = a.ext_q < (uint32_t)kTile ? a.ext_q : (uint32_t)kRectElems / a.tile_p;
        if (a.tile_p > a.ext_p) a.tile_p = a.ext_p;
        if (a.tile_q > a.ext_q) a.tile_q = a.ext_q;
        a.pitch = a.tile_p | 1;
        if ((uint64_t)a.tile_q * a.pitch > (uint64_t)(kRectElems + kRectElems / 2)) return false;
        a.div_tile_p = make_fastdiv(a.tile_p);
        a.div_tile_q = make_fastdiv(a.tile_q);
        a.tiles_q = (a.ext_q + a.tile_q - 1) / a.tile_q;
        a.tiles_p = (a.ext_p + a.tile_p - 1) / a.tile_p;
        a.div_tq = make_fastdiv(a.tiles_q);
        a.div_tp = make_fastdiv(a.tiles_p);
        const uint64_t rblocks = (uint64_t)a.tiles_q * a.tiles_p * nb;
        if (rblocks >= (1ull << 31)) return false;
        ew_tile_rect_kernel<F><<<(unsigned)rblocks, 256, 0, rt().stream>>>(a, f);
        note_launch("ew_tile_rect_transpose");
        return check_launch("ew_tile_rect");
    }
    uint64_t blocks = (uint64_t)a.tiles_q * a.tiles_p * nb;
    if (blocks >= (1ull << 31)) return false;  // caller falls back to the scalar gather
    ew_tile_kernel<F><<<(unsigned)blocks, dim3(kScalarTile, 256 / kScalarTile), 0, rt().stream>>>(a, f);
    note_launch("ew_tile_transpose");
    return check_launch("ew_tile");
}

// Outer-broadcast plan: returns 0 = not applicable, 1 = launched, -1 = error.
template <typename F>
static int try_launch_outer(const F& f, const Folded& fo, const float* const* in, float* out) {
    if constexpr (F::NIN != 2) {
        return 0;
    } else {
        // (a 4-element inner axis leaves one float4 per row: the tile's stores no longer coalesce and ew_nd is 20 % faster)
        if (!rt().outer_tile || fo.ndim < 3 || fo.extent[0] < 8) return 0;
        int d1 = -1, d2 = -1;
        for (int d = 1; d < fo.ndim; d++) {
            if (fo.extent[d] < (uint64_t)kOT) continue;
            if (d1 < 0 && fo.in_stride[0][d] == 0 && fo.in_stride[1][d] != 0) d1 = d;
            else if (d2 < 0 && fo.in_stride[1][d] == 0 && fo.in_stride[0][d] != 0) d2 = d;
        }
        if (d1 < 0 || d2 < 0) return 0;
        OuterArgs a;
        memset(&a, 0, sizeof a);
        a.out = out;
        a.x = in[0];
        a.y = in[1];
        a.ndim = fo.ndim;
        a.d1 = d1;
        a.d2 = d2;
        a.e1 = (uint32_t)fo.extent[d1];
        a.e2 = (uint32_t)fo.extent[d2];
        long long run = 1;
        uint64_t items = 1, total = 1;
        for (int d = 0; d < fo.ndim; d++) {
            uint64_t e = fo.extent[d];
            long long unit = 1;  // elements of the logical axis per iteration unit
            if (d == 0) e /= 4, unit = 4;
            if (d == d1 || d == d2) e = (e + kOT - 1) / kOT, unit = kOT;
            if (e >= (1ull << 31)) return 0;
            a.extent[d] = (uint32_t)e;
            a.div[d] = make_fastdiv((uint32_t)e);
            a.xs[d] = fo.in_stride[0][d] * unit;
            a.ys[d] = fo.in_stride[1][d] * unit;
            a.os[d] = run * unit;
            if (d == d1) a.o_s1 = run, a.y_s1 = fo.in_stride[1][d];
            if (d == d2) a.o_s2 = run, a.x_s2 = fo.in_stride[0][d];
            run *= (long long)fo.extent[d];
            items *= e;
            total *= fo.extent[d];
        }
        if (items >= (1ull << 31)) return 0;
        a.n_items = (uint32_t)items;
        a.x_inner_bcast = fo.in_stride[0][0] == 0;
        a.y_inner_bcast = fo.in_stride[1][0] == 0;
        a.stream_store = total * sizeof(float) >= (uint64_t)rt().stream_store_min_bytes;
        ew_outer_kernel<F><<<(unsigned)((items + 255) / 256), 256, 0, rt().stream>>>(a, f);
        note_launch("ew_outer_bcast");
        return check_launch("ew_outer") ? 1 : -1;
    }
}

template <typename F>
static bool run_functor(const F& f, const EwCall& c, const Folded& fo) {
    constexpr int NIN = F::NIN;
    const float* in[NIN];
    for (int i = 0; i < NIN; i++) in[i] = c.in[i];
    bool fast = !rt().force_general;
    uint64_t total = 1;
    for (int d = 0; d < fo.ndim; d++) total *= fo.extent[d];

    if (fast && fo.ndim >= 2 && fo.extent[0] < (1ull << 31)) {
        // transposed operand: inner stride > 1 but unit stride along another (large enough) axis
        for (int i = 0; i < NIN; i++) {
            if (fo.in_stride[i][0] <= 1) continue;
            for (int P = 1; P < fo.ndim; P++) {
                if (fo.in_stride[i][P] == 1 && fo.extent[P] >= 2 && fo.extent[P] < (1ull << 31) && fo.extent[0] >= 2 &&
                    fo.extent[P] * fo.extent[0] >= 256) {
                    uint64_t nb = total / fo.extent[0] / fo.extent[P];
                    if (nb < (1ull << 31) && launch_tile<F>(f, fo, in, c.out, P)) return true;
                    if (!last_error_string().empty()) return false;
                }
            }
        }
    }
    bool vec_ok = fast && ((uintptr_t)c.out % 16 == 0);
    for (int i = 0; i < NIN && vec_ok; i++) {
        int64_t s0 = fo.in_stride[i][0];
        if (s0 == 0) continue;  // broadcast along the inner axis: scalar load, no alignment needed
        if (s0 != 1 || (uintptr_t)in[i] % 16 != 0) vec_ok = false;
        for (int d = 1; d < fo.ndim && vec_ok; d++)
            if (fo.in_stride[i][d] % 4 != 0) vec_ok = false;
    }
    if constexpr (is_heavy<F>::value && NIN <= 2) {
        bool flat = vec_ok && fo.ndim == 1 && fo.extent[0] >= 4 && fo.extent[0] / 4 < (1ull << 31) && rt().heavy_stream >= F::kHeavy;
        for (int i = 0; i < NIN; i++) flat = flat && fo.in_stride[i][0] == 1;
        if (flat) {
            const uint64_t nb = fo.extent[0] & ~3ull, n4 = nb / 4;
            uint64_t blocks = (n4 + (uint64_t)kEwThreads * kHeavyUnroll - 1) / ((uint64_t)kEwThreads * kHeavyUnroll);
            static int resident = 0;  // CTAs of this instantiation that fit one SM (register-bound: 3..5)
            if (resident == 0) {
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, ew_heavy_stream_kernel<F>, kEwThreads, 0) != cudaSuccess || resident < 1)
                    resident = 1;
            }
            const uint64_t cap = (uint64_t)kNumSMs * (uint64_t)(rt().heavy_ctas_per_sm > 0 ? rt().heavy_ctas_per_sm : resident);
            if (blocks > cap) blocks = cap;
            ew_heavy_stream_kernel<F><<<(unsigned)blocks, kEwThreads, 0, rt().stream>>>(
                c.out, in[0], NIN == 2 ? in[NIN - 1] : nullptr, (uint32_t)n4, f,
                nb * sizeof(float) >= (uint64_t)rt().stream_store_min_bytes);
            note_launch("ew_heavy_stream");
            if (!check_launch("ew_heavy_stream")) return false;
            if (nb == fo.extent[0]) return true;
            Folded tail = fo;
            tail.extent[0] = fo.extent[0] - nb;
            const float* in_tail[NIN];
            for (int i = 0; i < NIN; i++) in_tail[i] = in[i] + nb;
            return launch_nd<1, F>(f, tail, in_tail, c.out + nb);
        }
    }
    if (vec_ok && fo.extent[0] % 4 == 0) {
        const int outer = try_launch_outer<F>(f, fo, in, c.out);
        if (outer != 0) return outer > 0;
        return launch_nd<4, F>(f, fo, in, c.out);
    }
    if (vec_ok && fo.ndim == 1 && fo.extent[0] >= 4) {
        // contiguous stream whose length is not a multiple of 4: vector body + scalar tail
        Folded body = fo, tail = fo;
        uint64_t nb = fo.extent[0] & ~3ull;
        body.extent[0] = nb;
        tail.extent[0] = fo.extent[0] - nb;
        const float* in_tail[NIN];
        for (int i = 0; i < NIN; i++) in_tail[i] = in[i] + (int64_t)nb * fo.in_stride[i][0];
        return launch_nd<4, F>(f, body, in, c.out) && launch_nd<1, F>(f, tail, in_tail, c.out + nb);
    }
    return launch_nd<1, F>(f, fo, in, c.out);
}


}  // namespace al
