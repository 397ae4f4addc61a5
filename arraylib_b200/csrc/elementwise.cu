// Elementwise engine: out(contiguous) = f(in_0[strided], in_1[strided], ...).
//
// Replaces the reference loops array_array_op (arraylib/src/arraylib.c:569-589), array_scalar_op
// (:491-501), array_op (:638-645), array_deep_copy (:248-256) and the where family (:774-838),
// all of which call offset_indexer (:116-123) per element. Here the layout is folded once on the
// host (layout.cu) and one of three sm_100a kernels is chosen:
//
//   ew_nd<VEC=4>   inner axis contiguous (stride 1) or broadcast (stride 0) in every operand:
//                  128-bit loads/stores, outer axes resolved with one mul-hi divmod per axis per
//                  vector. ndim==1 degenerates to the pure streaming kernel (config C2); stride-0
//                  outer axes give the broadcast case (config C3).
//   ew_tile        some operand is unit-stride along a different axis than the output
//                  (transposed view, config C5b): 32x32 tiles staged through padded shared memory
//                  so that both the global read and the global write are coalesced.
//   ew_nd<VEC=1>   anything else (arbitrary as_strided operands): scalar gather.
#include "al_internal.h"
#include "device_utils.cuh"

namespace al {


// ---- scalar functions: bit-for-bit the reference's (arraylib.c:12-49) ---------------------------
template <int OP>
__device__ __forceinline__ float binop_apply(float l, float r) {
    if constexpr (OP == AL_SUM) return __fadd_rn(l, r);
    else if constexpr (OP == AL_SUB) return __fsub_rn(l, r);
    else if constexpr (OP == AL_MUL) return __fmul_rn(l, r);
    else if constexpr (OP == AL_DIV) return __fdiv_rn(l, r);
    else if constexpr (OP == AL_MOD) return fmodf(l, r);  // fmod is exact, so f32 == round(f64)
    else if constexpr (OP == AL_POW) return (float)pow((double)l, (double)r);
    // glibc fmax/fmin: NaN-ignoring, and on a +-0 tie the LEFT operand is returned
    else if constexpr (OP == AL_MAX) return (l >= r || r != r) ? l : r;
    else if constexpr (OP == AL_MIN) return (l <= r || r != r) ? l : r;
    else if constexpr (OP == AL_EQ) return l == r ? 1.0f : 0.0f;
    else if constexpr (OP == AL_NE) return l != r ? 1.0f : 0.0f;
    else if constexpr (OP == AL_GT) return l > r ? 1.0f : 0.0f;
    else if constexpr (OP == AL_GE) return l >= r ? 1.0f : 0.0f;
    else if constexpr (OP == AL_LT) return l < r ? 1.0f : 0.0f;
    else return l <= r ? 1.0f : 0.0f;
}

// The reference promotes to double, calls libm and rounds back (arraylib.c:31-49). sqrt/ceil/
// floor/abs/neg are exact under that double rounding, so they stay in f32; the transcendentals
// run in FP64 on the device and round once.
template <int OP>
__device__ __forceinline__ float ufunc_apply(float v) {
    if constexpr (OP == AL_NEG) return __uint_as_float(__float_as_uint(v) ^ 0x80000000u);
    else if constexpr (OP == AL_ABS) return __uint_as_float(__float_as_uint(v) & 0x7fffffffu);
    else if constexpr (OP == AL_SQRT) return __fsqrt_rn(v);
    else if constexpr (OP == AL_CEIL) return ceilf(v);
    else if constexpr (OP == AL_FLOOR) return floorf(v);
    else {
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

// Functors: NIN inputs in, one float out.
template <int OP>
struct BinF {
    static constexpr int NIN = 2;
    __device__ __forceinline__ float operator()(const float* v) const { return binop_apply<OP>(v[0], v[1]); }
};
template <int OP>
struct ScalarF {
    static constexpr int NIN = 1;
    float s;
    __device__ __forceinline__ float operator()(const float* v) const { return binop_apply<OP>(v[0], s); }
};
template <int OP>
struct UnaryF {
    static constexpr int NIN = 1;
    __device__ __forceinline__ float operator()(const float* v) const { return ufunc_apply<OP>(v[0]); }
};
struct CopyF {
    static constexpr int NIN = 1;
    __device__ __forceinline__ float operator()(const float* v) const { return v[0]; }
};
// where: C truthiness of an f32, i.e. (x != 0): NaN selects lhs, +-0 selects rhs (arraylib.c:788)
struct WhereAAA {
    static constexpr int NIN = 3;
    __device__ __forceinline__ float operator()(const float* v) const { return v[0] != 0.0f ? v[1] : v[2]; }
};
struct WhereAAS {
    static constexpr int NIN = 2;
    float r;
    __device__ __forceinline__ float operator()(const float* v) const { return v[0] != 0.0f ? v[1] : r; }
};
struct WhereASA {
    static constexpr int NIN = 2;
    float l;
    __device__ __forceinline__ float operator()(const float* v) const { return v[0] != 0.0f ? l : v[1]; }
};
struct WhereASS {
    static constexpr int NIN = 1;
    float l, r;
    __device__ __forceinline__ float operator()(const float* v) const { return v[0] != 0.0f ? l : r; }
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

template <int VEC>
__device__ __forceinline__ typename VecT<VEC>::type load_vec(const float* p, bool bcast, bool reuse) {
    if constexpr (VEC == 1) {
        return reuse ? __ldg(p) : __ldcs(p);
    } else {
        if (bcast) {
            float s = __ldg(p);
            return make_float4(s, s, s, s);
        }
        return reuse ? __ldg(reinterpret_cast<const float4*>(p)) : __ldcs(reinterpret_cast<const float4*>(p));
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

template <int VEC, int UNROLL, typename F>
__global__ void __launch_bounds__(kEwThreads) ew_nd_kernel(const NdArgs<F::NIN> a, const F f) {
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
            for (int i = 0; i < NIN; i++) v[k][i] = load_vec<VEC>(a.in[i] + off[i], a.inner_bcast[i], a.reuse[i]);
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
};

constexpr int kTile = 32, kTileRows = 8;

template <typename F>
__global__ void __launch_bounds__(kTile * kTileRows) ew_tile_kernel(const TileArgs<F::NIN> a, const F f) {
    constexpr int NIN = F::NIN;
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
    // stage p-major operands: threadIdx.x runs along p (their unit-stride axis)
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
    __syncthreads();
    const uint32_t q = q0 + tx;
#pragma unroll
    for (int j = 0; j < kTile; j += kTileRows) {
        const uint32_t p = p0 + ty + j;
        if (q < a.ext_q && p < a.ext_p) {
            float x[NIN];
#pragma unroll
            for (int i = 0; i < NIN; i++) {
                if (a.via_smem >> i & 1) x[i] = tile[i][tx][ty + j];
                else x[i] = __ldg(a.in[i] + ibase[i] + (long long)q * a.sq[i] + (long long)p * a.sp[i]);
            }
            float* dst = a.out + obase + q + (long long)p * a.out_sp;
            const float y = f(x);
            if (a.stream_store) __stcs(dst, y);
            else *dst = y;
        }
    }
}

// ---- host-side launch plans ------------------------------------------------------------------------
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
    const long unroll = rt().ew_unroll;
    auto blocks = [&](int u) { return (unsigned)((n_items + (uint64_t)kEwThreads * u - 1) / ((uint64_t)kEwThreads * u)); };
    cudaStream_t s = rt().stream;
    if (unroll >= 8) ew_nd_kernel<VEC, 8, F><<<blocks(8), kEwThreads, 0, s>>>(a, f);
    else if (unroll >= 4) ew_nd_kernel<VEC, 4, F><<<blocks(4), kEwThreads, 0, s>>>(a, f);
    else if (unroll >= 2) ew_nd_kernel<VEC, 2, F><<<blocks(2), kEwThreads, 0, s>>>(a, f);
    else ew_nd_kernel<VEC, 1, F><<<blocks(1), kEwThreads, 0, s>>>(a, f);
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

template <typename F>
static bool launch_tile(const F& f, const Folded& fo, const float* const* in, float* out, int P) {
    constexpr int NIN = F::NIN;
    TileArgs<NIN> a;
    memset(&a, 0, sizeof a);
    a.out = out;
    a.ext_q = (uint32_t)fo.extent[0];
    a.ext_p = (uint32_t)fo.extent[P];
    a.tiles_q = (a.ext_q + kTile - 1) / kTile;
    a.tiles_p = (a.ext_p + kTile - 1) / kTile;
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
    uint64_t blocks = (uint64_t)a.tiles_q * a.tiles_p * nb;
    if (blocks >= (1ull << 31)) return false;  // caller falls back to the scalar gather
    ew_tile_kernel<F><<<(unsigned)blocks, dim3(kTile, kTileRows), 0, rt().stream>>>(a, f);
    note_launch("ew_tile_transpose");
    return check_launch("ew_tile");
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
                if (fo.in_stride[i][P] == 1 && fo.extent[P] >= 8 && fo.extent[P] < (1ull << 31) && fo.extent[0] >= 8) {
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
    if (vec_ok && fo.extent[0] % 4 == 0) return launch_nd<4, F>(f, fo, in, c.out);
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

template <template <int> class FT, int N, typename Make>
static bool dispatch_op(int op, const EwCall& c, const Folded& fo, Make make) {
    if constexpr (N == 0) {
        set_error("ValueError: unknown op code %d", op);
        return false;
    } else {
        if (op == N - 1) return run_functor(make(FT<N - 1>{}), c, fo);
        return dispatch_op<FT, N - 1>(op, c, fo, make);
    }
}

bool launch_elementwise(const EwCall& c, const Folded& fo) {
    switch (c.kind) {
        case EW_BINARY:
            return dispatch_op<BinF, AL_BINOP_COUNT>(c.op, c, fo, [](auto f) { return f; });
        case EW_SCALAR: {
            float s = c.s0;
            return dispatch_op<ScalarF, AL_BINOP_COUNT>(c.op, c, fo, [s](auto f) { f.s = s; return f; });
        }
        case EW_UNARY:
            return dispatch_op<UnaryF, AL_UFUNC_COUNT>(c.op, c, fo, [](auto f) { return f; });
        case EW_COPY: return run_functor(CopyF{}, c, fo);
        case EW_WHERE_AAA: return run_functor(WhereAAA{}, c, fo);
        case EW_WHERE_AAS: return run_functor(WhereAAS{c.s0}, c, fo);
        case EW_WHERE_ASA: return run_functor(WhereASA{c.s0}, c, fo);
        case EW_WHERE_ASS: return run_functor(WhereASS{c.s0, c.s1}, c, fo);
    }
    return false;
}

// ---- strided destination (setters; compat path) ------------------------------------------------------
struct AssignArgs {
    float* dst;
    const float* src;
    float scalar;
    uint32_t n;
    int ndim;
    uint32_t extent[kMaxDims];
    FastDiv div[kMaxDims];
    long long dstride[kMaxDims], sstride[kMaxDims];
};

__global__ void __launch_bounds__(256) strided_assign_kernel(const AssignArgs a) {
    uint32_t item = blockIdx.x * 256u + threadIdx.x;
    if (item >= a.n) return;
    long long doff = 0, soff = 0;
    uint32_t rem = item;
    for (int d = 0; d < a.ndim; d++) {
        uint32_t idx;
        if (d == a.ndim - 1) {
            idx = rem;
        } else {
            uint32_t q = fdiv(rem, a.div[d]);
            idx = rem - q * a.extent[d];
            rem = q;
        }
        doff += (long long)idx * a.dstride[d];
        soff += (long long)idx * a.sstride[d];
    }
    a.dst[doff] = a.src ? a.src[soff] : a.scalar;
}

bool launch_strided_assign(float* dst, const Folded& fo, const float* src, float scalar) {
    uint64_t total = 1;
    for (int d = 0; d < fo.ndim; d++) total *= fo.extent[d];
    if (total >= (1ull << 31)) {
        set_error("NotImplementedError: strided assignment of %llu elements (limit 2^31-1)", (unsigned long long)total);
        return false;
    }
    AssignArgs a;
    memset(&a, 0, sizeof a);
    a.dst = dst;
    a.src = src;
    a.scalar = scalar;
    a.n = (uint32_t)total;
    a.ndim = fo.ndim;
    for (int d = 0; d < fo.ndim; d++) {
        a.extent[d] = (uint32_t)fo.extent[d];
        a.div[d] = make_fastdiv(a.extent[d]);
        a.dstride[d] = fo.out_stride[d];
        a.sstride[d] = src ? fo.in_stride[0][d] : 0;
    }
    strided_assign_kernel<<<(unsigned)((total + 255) / 256), 256, 0, rt().stream>>>(a);
    note_launch("strided_assign");
    return check_launch("strided_assign");
}

// ---- shared front end ----------------------------------------------------------------------------------
// Runs an elementwise call whose inputs all share `shape` (strides may be 0 for broadcast axes)
// and returns a fresh contiguous array of that shape.
static NDArray* run_to_new(EwCall call, size_t ndim, const size_t* shape, const size_t* const* strides) {
    NDArray* out = new_array(shape, ndim);
    if (!out) return nullptr;
    Folded fo;
    call.out = out->ptr;
    if (!fold_layout(ndim, shape, out->lay->stride, call.nin, strides, &fo) || !launch_elementwise(call, fo)) {
        std::string keep = last_error_string();
        array_free(out);
        restore_error(keep);
        return nullptr;
    }
    return out;
}

static NDArray* same_shape_call(EwCall call, const NDArray* const* arrs, const char* who) {
    for (int i = 0; i < call.nin; i++)
        if (!valid(arrs[i], who)) return nullptr;
    const size_t* strides[3];
    for (int i = 0; i < call.nin; i++) {
        AL_REQUIRE(arrs[i]->lay->ndim == arrs[0]->lay->ndim, "ValueError: %s rank mismatch", who);
        for (size_t d = 0; d < arrs[0]->lay->ndim; d++)
            AL_REQUIRE(arrs[i]->lay->shape[d] == arrs[0]->lay->shape[d], "ValueError: %s shape mismatch", who);
        call.in[i] = arrs[i]->ptr;
        strides[i] = arrs[i]->lay->stride;
    }
    return run_to_new(call, arrs[0]->lay->ndim, arrs[0]->lay->shape, strides);
}

}  // namespace al

using namespace al;

extern "C" {

// arraylib.c:248-256
NDArray* array_deep_copy(const NDArray* src) {
    AL_ENTER;
    EwCall c;
    c.kind = EW_COPY;
    c.nin = 1;
    const NDArray* arrs[1] = {src};
    return same_shape_call(c, arrs, "array_deep_copy");
}

// arraylib.c:569-589 with is_broadcastable (:171-178) and layout_broadcast (:180-206) folded in.
NDArray* array_array_op_named(int op, const NDArray* lhs, const NDArray* rhs) {
    AL_ENTER;
    if (!valid(lhs, "array_array_op") || !valid(rhs, "array_array_op")) return nullptr;
    AL_REQUIRE(op >= 0 && op < AL_BINOP_COUNT, "ValueError: unknown binary op %d", op);
    size_t nl = lhs->lay->ndim, nr = rhs->lay->ndim, n = nl > nr ? nl : nr;
    AL_REQUIRE(n <= 64, "ValueError: too many dimensions");
    size_t shape[64], sl[64], sr[64];
    for (size_t k = 0; k < n; k++) {  // k counts axes from the right
        size_t o = n - 1 - k;
        size_t el = k < nl ? lhs->lay->shape[nl - 1 - k] : 1;
        size_t er = k < nr ? rhs->lay->shape[nr - 1 - k] : 1;
        AL_REQUIRE(el == er || el == 1 || er == 1, "ValueError: can not broadcast.");
        size_t e = el > er ? el : er;
        shape[o] = e;
        sl[o] = (k < nl && el == e) ? lhs->lay->stride[nl - 1 - k] : 0;
        sr[o] = (k < nr && er == e) ? rhs->lay->stride[nr - 1 - k] : 0;
    }
    EwCall c;
    c.kind = EW_BINARY;
    c.op = op;
    c.nin = 2;
    c.in[0] = lhs->ptr;
    c.in[1] = rhs->ptr;
    const size_t* strides[2] = {sl, sr};
    return run_to_new(c, n, shape, strides);
}

// arraylib.c:491-501
NDArray* array_scalar_op_named(int op, const NDArray* src, f32 rhs) {
    AL_ENTER;
    AL_REQUIRE(op >= 0 && op < AL_BINOP_COUNT, "ValueError: unknown binary op %d", op);
    EwCall c;
    c.kind = EW_SCALAR;
    c.op = op;
    c.s0 = rhs;
    c.nin = 1;
    const NDArray* arrs[1] = {src};
    return same_shape_call(c, arrs, "array_scalar_op");
}

// arraylib.c:638-645 with the uniop given as a table entry
NDArray* array_op_named(int ufunc, const NDArray* src) {
    AL_ENTER;
    AL_REQUIRE(ufunc >= 0 && ufunc < AL_UFUNC_COUNT, "ValueError: unknown ufunc %d", ufunc);
    EwCall c;
    c.kind = EW_UNARY;
    c.op = ufunc;
    c.nin = 1;
    const NDArray* arrs[1] = {src};
    return same_shape_call(c, arrs, "array_op");
}

NDArray* array_op(uniop, const NDArray*) {
    AL_ENTER;
    set_error("NotImplementedError: host callbacks cannot run on the device; use array_op_named with an al_ufunc entry");
    return nullptr;
}

#define AL_BINARY_WRAPPERS(name, code)                                                                          \
    NDArray* array_array_##name(const NDArray* l, const NDArray* r) { return array_array_op_named(code, l, r); } \
    NDArray* array_scalar_##name(const NDArray* s, f32 r) { return array_scalar_op_named(code, s, r); }
AL_BINARY_WRAPPERS(sum, AL_SUM)
AL_BINARY_WRAPPERS(sub, AL_SUB)
AL_BINARY_WRAPPERS(mul, AL_MUL)
AL_BINARY_WRAPPERS(div, AL_DIV)
AL_BINARY_WRAPPERS(mod, AL_MOD)
AL_BINARY_WRAPPERS(pow, AL_POW)
AL_BINARY_WRAPPERS(max, AL_MAX)
AL_BINARY_WRAPPERS(min, AL_MIN)
AL_BINARY_WRAPPERS(eq, AL_EQ)
AL_BINARY_WRAPPERS(ne, AL_NE)
AL_BINARY_WRAPPERS(gt, AL_GT)
AL_BINARY_WRAPPERS(ge, AL_GE)
AL_BINARY_WRAPPERS(lt, AL_LT)
AL_BINARY_WRAPPERS(le, AL_LE)

#define AL_UNARY_WRAPPER(name, code) \
    NDArray* array_##name(const NDArray* s) { return array_op_named(code, s); }
AL_UNARY_WRAPPER(neg, AL_NEG)
AL_UNARY_WRAPPER(abs, AL_ABS)
AL_UNARY_WRAPPER(sqrt, AL_SQRT)
AL_UNARY_WRAPPER(exp, AL_EXP)
AL_UNARY_WRAPPER(log, AL_LOG)
AL_UNARY_WRAPPER(sin, AL_SIN)
AL_UNARY_WRAPPER(cos, AL_COS)
AL_UNARY_WRAPPER(tan, AL_TAN)
AL_UNARY_WRAPPER(asin, AL_ASIN)
AL_UNARY_WRAPPER(acos, AL_ACOS)
AL_UNARY_WRAPPER(atan, AL_ATAN)
AL_UNARY_WRAPPER(sinh, AL_SINH)
AL_UNARY_WRAPPER(cosh, AL_COSH)
AL_UNARY_WRAPPER(tanh, AL_TANH)
AL_UNARY_WRAPPER(asinh, AL_ASINH)
AL_UNARY_WRAPPER(acosh, AL_ACOSH)
AL_UNARY_WRAPPER(atanh, AL_ATANH)
AL_UNARY_WRAPPER(ceil, AL_CEIL)
AL_UNARY_WRAPPER(floor, AL_FLOOR)

// arraylib.c:774-838 (no broadcasting: shapes must match)
NDArray* array_array_array_where(const NDArray* cond, const NDArray* lhs, const NDArray* rhs) {
    AL_ENTER;
    EwCall c;
    c.kind = EW_WHERE_AAA;
    c.nin = 3;
    const NDArray* arrs[3] = {cond, lhs, rhs};
    return same_shape_call(c, arrs, "where");
}
NDArray* array_array_scalar_where(const NDArray* cond, const NDArray* lhs, f32 rhs) {
    AL_ENTER;
    EwCall c;
    c.kind = EW_WHERE_AAS;
    c.nin = 2;
    c.s0 = rhs;
    const NDArray* arrs[2] = {cond, lhs};
    return same_shape_call(c, arrs, "where");
}
NDArray* array_scalar_array_where(const NDArray* cond, f32 lhs, const NDArray* rhs) {
    AL_ENTER;
    EwCall c;
    c.kind = EW_WHERE_ASA;
    c.nin = 2;
    c.s0 = lhs;
    const NDArray* arrs[2] = {cond, rhs};
    return same_shape_call(c, arrs, "where");
}
NDArray* array_scalar_scalar_where(const NDArray* cond, f32 lhs, f32 rhs) {
    AL_ENTER;
    EwCall c;
    c.kind = EW_WHERE_ASS;
    c.nin = 1;
    c.s0 = lhs;
    c.s1 = rhs;
    const NDArray* arrs[1] = {cond};
    return same_shape_call(c, arrs, "where");
}

// ---- slice assignment (arraylib.c:359-415) --------------------------------------------------------------
static bool slice_layout(const NDArray* dst, const size_t* start, const size_t* end, const size_t* step,
                         size_t* shape, size_t* stride, size_t* offset) {
    size_t nd = dst->lay->ndim;
    *offset = 0;
    for (size_t i = 0; i < nd; i++) {
        if (!(start[i] < end[i])) return set_error("ValueError: start index >= end index."), false;
        if (!(end[i] <= dst->lay->shape[i])) return set_error("ValueError: end index out of bounds."), false;
        if (!(step[i] > 0)) return set_error("ValueError: non-positive step size."), false;
        shape[i] = (end[i] - start[i] + step[i] - 1) / step[i];
        stride[i] = dst->lay->stride[i] * step[i];
        *offset += start[i] * dst->lay->stride[i];
    }
    return true;
}

NDArray* array_set_view_from_scalar(NDArray* dst, f32 value, const size_t* start, const size_t* end,
                                    const size_t* step) {
    AL_ENTER;
    if (!valid(dst, "array_set_view_from_scalar")) return nullptr;
    size_t nd = dst->lay->ndim, shape[64], stride[64], off;
    AL_REQUIRE(nd <= 64, "ValueError: too many dimensions");
    if (!slice_layout(dst, start, end, step, shape, stride, &off)) return nullptr;
    Folded fo;
    if (!fold_layout(nd, shape, stride, 0, nullptr, &fo)) return nullptr;
    if (!launch_strided_assign(dst->ptr + off, fo, nullptr, value)) return nullptr;
    return dst;
}

NDArray* array_set_view_from_array(NDArray* dst, const NDArray* src, const size_t* start, const size_t* end,
                                   const size_t* step) {
    AL_ENTER;
    if (!valid(dst, "array_set_view_from_array") || !valid(src, "array_set_view_from_array")) return nullptr;
    size_t nd = dst->lay->ndim, shape[64], stride[64], off;
    AL_REQUIRE(nd <= 64, "ValueError: too many dimensions");
    AL_REQUIRE(src->lay->ndim == nd, "ValueError: dimension mismatch.");
    if (!slice_layout(dst, start, end, step, shape, stride, &off)) return nullptr;
    // the reference only checks the element COUNT (:406-407) and walks both sides in logical
    // order; iterate over the source's shape when the per-axis extents differ.
    AL_REQUIRE(count_of(shape, nd) == count_of(src->lay->shape, nd), "ValueError: shape mismatch.");
    for (size_t i = 0; i < nd; i++)
        AL_REQUIRE(shape[i] == src->lay->shape[i], "NotImplementedError: slice assignment with differing per-axis extents");
    Folded fo;
    const size_t* sstr[1] = {src->lay->stride};
    if (!fold_layout(nd, shape, stride, 1, sstr, &fo)) return nullptr;
    if (!launch_strided_assign(dst->ptr + off, fo, src->ptr, 0.f)) return nullptr;
    return dst;
}

}  // extern "C"
