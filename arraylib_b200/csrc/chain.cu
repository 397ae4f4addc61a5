// Fused elementwise chains: out = step_n(... step_1(in_0) ...) in ONE pass over memory.
//
// The reference evaluates `a + b + c + 1.0` as three eager loops with two materialised temporaries
// (three array_array_op / array_scalar_op calls, arraylib/src/arraylib.c:569-589,
// 491-501). On config C3 that is 51.6 GB of traffic for a result that needs 17.2 GB. Because the
// path is HBM-bound, a small per-element interpreter is free: the kernel loads up to four strided /
// broadcast operands once, runs a left-deep chain of up to eight steps on registers and stores once.
// Every step is the same correctly rounded f32 (or f64-libm-then-round) operation the eager kernels
// apply, in the same order, so fused results are bit-identical to the eager ones.
#include "elementwise.cuh"

namespace al {

constexpr int kChainMaxSteps = 8;
constexpr int kChainMaxInputs = 4;

enum ChainKind { CH_ACC_IN = 0, CH_IN_ACC = 1, CH_ACC_SCALAR = 2, CH_SCALAR_ACC = 3, CH_UNARY = 4 };

template <int NIN>
struct ChainArgs {
    NdArgs<NIN> nd;        // iteration space; along axis `blk_axis` one item covers kChainBlock rows
    int n_steps;
    int blk_axis;          // -1: no blocking (one output vector per thread)
    uint32_t blk_extent;   // true extent of the blocked axis
    long long blk_out;     // output stride (elements) of the blocked axis
    long long blk_in[NIN]; // operand strides of the blocked axis (0 = invariant: loaded once)
    unsigned char kind[kChainMaxSteps];
    unsigned char op[kChainMaxSteps];
    float scalar[kChainMaxSteps];
    int* nan_flag;         // raised when any output of this launch is a NaN (chain_nan_fix_kernel then replays those)
};

// One chain step over N lanes. The op is decoded ONCE (the switch is outside the lane loops), so the
// interpreter costs a handful of instructions per step per thread, not per element. The steps use the
// *_raw functions: a fused chain is issue-bound, and the per-element NaN fix-up of the eager kernels
// (x86 sign and payload, elementwise.cuh) after EVERY step cost 2.5 -> 6.0 ms on the C3 chain. Instead the
// fix-up is deferred: every non-NaN final result is bit-identical to the eager path whatever happened to NaNs on
// the way (max/min and the comparisons swallow a NaN the same way for any payload). A thread that produced a NaN
// (one compare per element) raises a device flag, and chain_nan_fix_kernel -- launched behind every chain, a few
// microseconds when the flag is down -- replays exactly those elements from their operands with the eager kernels'
// *_apply functions, so a chain returns the reference's NaN sign and payload too. (Replaying inside the hot kernel
// was measured: the extra live state took the C3 chain from 74 to 96 registers and from 2.67 to 3.28 ms.)
template <int OP, int N>
__device__ __forceinline__ void lanes_bin(float* acc, const float* other, bool reversed) {
    if (reversed) {
#pragma unroll
        for (int c = 0; c < N; c++) acc[c] = binop_raw<OP>(other[c], acc[c]);
    } else {
#pragma unroll
        for (int c = 0; c < N; c++) acc[c] = binop_raw<OP>(acc[c], other[c]);
    }
}
template <int OP, int N>
__device__ __forceinline__ void lanes_un(float* acc) {
#pragma unroll
    for (int c = 0; c < N; c++) acc[c] = ufunc_value<OP>(acc[c]);
}

template <bool FULL, int N>
__device__ __forceinline__ void chain_step(float* acc, const float* other, int kind, int op) {
    if (kind == CH_UNARY) {
        switch (op) {
            case AL_NEG: lanes_un<AL_NEG, N>(acc); return;
            case AL_ABS: lanes_un<AL_ABS, N>(acc); return;
            case AL_SQRT: lanes_un<AL_SQRT, N>(acc); return;
            case AL_CEIL: lanes_un<AL_CEIL, N>(acc); return;
            case AL_FLOOR: lanes_un<AL_FLOOR, N>(acc); return;
            default: break;
        }
        if constexpr (FULL) {
            switch (op) {
                case AL_EXP: lanes_un<AL_EXP, N>(acc); return;
                case AL_LOG: lanes_un<AL_LOG, N>(acc); return;
                case AL_SIN: lanes_un<AL_SIN, N>(acc); return;
                case AL_COS: lanes_un<AL_COS, N>(acc); return;
                case AL_TAN: lanes_un<AL_TAN, N>(acc); return;
                case AL_ASIN: lanes_un<AL_ASIN, N>(acc); return;
                case AL_ACOS: lanes_un<AL_ACOS, N>(acc); return;
                case AL_ATAN: lanes_un<AL_ATAN, N>(acc); return;
                case AL_SINH: lanes_un<AL_SINH, N>(acc); return;
                case AL_COSH: lanes_un<AL_COSH, N>(acc); return;
                case AL_TANH: lanes_un<AL_TANH, N>(acc); return;
                case AL_ASINH: lanes_un<AL_ASINH, N>(acc); return;
                case AL_ACOSH: lanes_un<AL_ACOSH, N>(acc); return;
                case AL_ATANH: lanes_un<AL_ATANH, N>(acc); return;
                default: break;
            }
        }
        return;
    }
    const bool rev = (kind == CH_IN_ACC || kind == CH_SCALAR_ACC);
    switch (op) {
        case AL_SUM: lanes_bin<AL_SUM, N>(acc, other, rev); return;
        case AL_SUB: lanes_bin<AL_SUB, N>(acc, other, rev); return;
        case AL_MUL: lanes_bin<AL_MUL, N>(acc, other, rev); return;
        case AL_DIV: lanes_bin<AL_DIV, N>(acc, other, rev); return;
        case AL_MAX: lanes_bin<AL_MAX, N>(acc, other, rev); return;
        case AL_MIN: lanes_bin<AL_MIN, N>(acc, other, rev); return;
        case AL_EQ: lanes_bin<AL_EQ, N>(acc, other, rev); return;
        case AL_NE: lanes_bin<AL_NE, N>(acc, other, rev); return;
        case AL_GT: lanes_bin<AL_GT, N>(acc, other, rev); return;
        case AL_GE: lanes_bin<AL_GE, N>(acc, other, rev); return;
        case AL_LT: lanes_bin<AL_LT, N>(acc, other, rev); return;
        case AL_LE: lanes_bin<AL_LE, N>(acc, other, rev); return;
        default: break;
    }
    if constexpr (FULL) {
        if (op == AL_MOD) lanes_bin<AL_MOD, N>(acc, other, rev);
        else if (op == AL_POW) lanes_bin<AL_POW, N>(acc, other, rev);
    }
}

template <int OP>
__device__ __forceinline__ float bin_exact_dir(float acc, float other, bool rev) {
    return rev ? binop_apply<OP>(other, acc) : binop_apply<OP>(acc, other);
}
__device__ __noinline__ float chain_exact_step(float acc, float other, int kind, int op) {
    if (kind == CH_UNARY) {
        switch (op) {
            case AL_NEG: return ufunc_apply<AL_NEG>(acc);
            case AL_ABS: return ufunc_apply<AL_ABS>(acc);
            case AL_SQRT: return ufunc_apply<AL_SQRT>(acc);
            case AL_CEIL: return ufunc_apply<AL_CEIL>(acc);
            case AL_FLOOR: return ufunc_apply<AL_FLOOR>(acc);
            case AL_EXP: return ufunc_apply<AL_EXP>(acc);
            case AL_LOG: return ufunc_apply<AL_LOG>(acc);
            case AL_SIN: return ufunc_apply<AL_SIN>(acc);
            case AL_COS: return ufunc_apply<AL_COS>(acc);
            case AL_TAN: return ufunc_apply<AL_TAN>(acc);
            case AL_ASIN: return ufunc_apply<AL_ASIN>(acc);
            case AL_ACOS: return ufunc_apply<AL_ACOS>(acc);
            case AL_ATAN: return ufunc_apply<AL_ATAN>(acc);
            case AL_SINH: return ufunc_apply<AL_SINH>(acc);
            case AL_COSH: return ufunc_apply<AL_COSH>(acc);
            case AL_TANH: return ufunc_apply<AL_TANH>(acc);
            case AL_ASINH: return ufunc_apply<AL_ASINH>(acc);
            case AL_ACOSH: return ufunc_apply<AL_ACOSH>(acc);
            default: return ufunc_apply<AL_ATANH>(acc);
        }
    }
    const bool rev = (kind == CH_IN_ACC || kind == CH_SCALAR_ACC);
    switch (op) {
        case AL_SUM: return bin_exact_dir<AL_SUM>(acc, other, rev);
        case AL_SUB: return bin_exact_dir<AL_SUB>(acc, other, rev);
        case AL_MUL: return bin_exact_dir<AL_MUL>(acc, other, rev);
        case AL_DIV: return bin_exact_dir<AL_DIV>(acc, other, rev);
        case AL_MOD: return bin_exact_dir<AL_MOD>(acc, other, rev);
        case AL_POW: return bin_exact_dir<AL_POW>(acc, other, rev);
        case AL_MAX: return bin_exact_dir<AL_MAX>(acc, other, rev);
        case AL_MIN: return bin_exact_dir<AL_MIN>(acc, other, rev);
        case AL_EQ: return bin_exact_dir<AL_EQ>(acc, other, rev);
        case AL_NE: return bin_exact_dir<AL_NE>(acc, other, rev);
        case AL_GT: return bin_exact_dir<AL_GT>(acc, other, rev);
        case AL_GE: return bin_exact_dir<AL_GE>(acc, other, rev);
        case AL_LT: return bin_exact_dir<AL_LT>(acc, other, rev);
        default: return bin_exact_dir<AL_LE>(acc, other, rev);
    }
}

// Replays ONE element's chain with the eager kernels' NaN rules (slow path: only for elements whose fast result is a NaN).
template <class Steps>
__device__ __noinline__ float chain_exact(const Steps& ca, float v0, float v1, float v2, float v3) {
    float acc = v0;
    int next = 1;
    for (int s = 0; s < ca.n_steps; s++) {
        const int kind = ca.kind[s];
        float other = ca.scalar[s];
        if (kind == CH_ACC_IN || kind == CH_IN_ACC) {
            other = next == 1 ? v1 : next == 2 ? v2 : v3;
            next++;
        }
        acc = chain_exact_step(acc, other, kind, ca.op[s]);
    }
    return acc;
}

// A thread owns U output vectors: U consecutive rows of the blocked axis (U = kChainBlock) or a
// single vector (U = 1). Operands that do not vary along the blocked axis are loaded once, and each
// chain step is decoded once for all U*VEC lanes, which keeps the interpreter off the critical path
// (measured on C3: 8.8 ms with per-vector decoding and 3 loads per output -> see DESIGN.md).
constexpr int kChainBlock = 4;

template <int VEC, int NIN, bool FULL, int U>
__global__ void __launch_bounds__(kEwThreads) ew_chain_kernel(const __grid_constant__ ChainArgs<NIN> ca) {
    using V = typename VecT<VEC>::type;
    const NdArgs<NIN>& a = ca.nd;
    const uint32_t item = blockIdx.x * (uint32_t)kEwThreads + threadIdx.x;
    if (item >= a.n_items) return;
    long long off[NIN], ooff = 0;
#pragma unroll
    for (int i = 0; i < NIN; i++) off[i] = 0;
    uint32_t rem = item, row0 = 0;
    long long ostride = VEC;  // contiguous output: running product of the (true) extents
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
        const bool blocked = (U > 1 && d == ca.blk_axis);
        if (blocked) row0 = idx * U;
#pragma unroll
        for (int i = 0; i < NIN; i++) off[i] += (long long)idx * a.stride[i][d];
        ooff += (long long)idx * ostride * (blocked ? U : 1);
        ostride *= blocked ? (long long)ca.blk_extent : (long long)a.extent[d];
        if (last) break;
    }
    float lanes[NIN][U][VEC];
#pragma unroll
    for (int i = 0; i < NIN; i++) {
#pragma unroll
        for (int u = 0; u < U; u++) {
            const bool live = U == 1 || row0 + u < ca.blk_extent;
            const bool fresh = u == 0 || ca.blk_in[i] != 0;  // invariant operands reuse the first load
            if (live && fresh) {
                const V t = load_vec<VEC>(a.in[i] + off[i] + (U > 1 ? u * ca.blk_in[i] : 0), a.inner_bcast[i], a.reuse[i], a.load_policy);
                if constexpr (VEC == 1) {
                    lanes[i][u][0] = t;
                } else {
                    lanes[i][u][0] = t.x, lanes[i][u][1] = t.y, lanes[i][u][2] = t.z, lanes[i][u][3] = t.w;
                }
            } else {
#pragma unroll
                for (int c = 0; c < VEC; c++) lanes[i][u][c] = lanes[i][0][c];
            }
        }
    }
    float acc[U][VEC];
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
        for (int c = 0; c < VEC; c++) acc[u][c] = lanes[0][u][c];
    int next = 1;
    for (int s = 0; s < ca.n_steps; s++) {
        const int kind = ca.kind[s], op = ca.op[s];
        float other[U][VEC];
        if (kind == CH_ACC_IN || kind == CH_IN_ACC) {
#pragma unroll
            for (int u = 0; u < U; u++)
#pragma unroll
                for (int c = 0; c < VEC; c++) {
                    float o = lanes[NIN > 1 ? 1 : 0][u][c];
                    if (NIN > 2 && next == 2) o = lanes[NIN > 2 ? 2 : 0][u][c];
                    if (NIN > 3 && next == 3) o = lanes[NIN > 3 ? 3 : 0][u][c];
                    other[u][c] = o;
                }
            next++;
        } else {
#pragma unroll
            for (int u = 0; u < U; u++)
#pragma unroll
                for (int c = 0; c < VEC; c++) other[u][c] = ca.scalar[s];
        }
        chain_step<FULL, U * VEC>(&acc[0][0], &other[0][0], kind, op);
    }
    bool any_nan = false;
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
        for (int c = 0; c < VEC; c++) any_nan |= acc[u][c] != acc[u][c];
    if (any_nan) atomicOr(ca.nan_flag, 1);  // rare; chain_nan_fix_kernel repairs the NaN elements afterwards
#pragma unroll
    for (int u = 0; u < U; u++) {
        if (U == 1 || row0 + u < ca.blk_extent) {
            V r;
            if constexpr (VEC == 1) r = acc[u][0];
            else r = make_float4(acc[u][0], acc[u][1], acc[u][2], acc[u][3]);
            store_vec<VEC>(a.out + ooff + (U > 1 ? u * ca.blk_out : 0), r, a.stream_store);
        }
    }
}

// ---- specialised blocked kernel: float4 lanes, arithmetic ops only, invariance known at compile time ----
// INV is a bitmask of the operands that do not vary along the blocked axis: they occupy one float4
// instead of U, every load is unconditional (ragged blocks clamp the row instead of predicating) and
// nothing about the operand layout is decided at run time. On C3 this took the fused chain from
// 511 to well under 200 issued instructions per thread (ncu), i.e. back under the issue budget of a
// write-bound kernel.
constexpr int kChainRows = 8;

__device__ __forceinline__ void f4_to_lanes(const float4& v, float* l) { l[0] = v.x, l[1] = v.y, l[2] = v.z, l[3] = v.w; }

template <int NIN, int INV, int I>
__device__ __forceinline__ void chain_apply_operand(float (&acc)[kChainRows][4], const float4 (&inv)[NIN],
                                                    const float4 (&var)[NIN][kChainRows], int kind, int op) {
    float other[kChainRows][4];
#pragma unroll
    for (int u = 0; u < kChainRows; u++) f4_to_lanes((INV >> I & 1) ? inv[I] : var[I][u], other[u]);
    chain_step<false, kChainRows * 4>(&acc[0][0], &other[0][0], kind, op);
}

template <int NIN, int INV>
__global__ void __launch_bounds__(kEwThreads) ew_chain_blocked_kernel(const __grid_constant__ ChainArgs<NIN> ca) {
    constexpr int U = kChainRows;
    const NdArgs<NIN>& a = ca.nd;
    const uint32_t item = blockIdx.x * (uint32_t)kEwThreads + threadIdx.x;
    if (item >= a.n_items) return;
    long long off[NIN], ooff = 0;
#pragma unroll
    for (int i = 0; i < NIN; i++) off[i] = 0;
    uint32_t rem = item, row0 = 0;
    long long ostride = 4;
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
        const bool blocked = d == ca.blk_axis;
        if (blocked) row0 = idx * U;
#pragma unroll
        for (int i = 0; i < NIN; i++) off[i] += (long long)idx * a.stride[i][d];
        ooff += (long long)idx * ostride * (blocked ? U : 1);
        ostride *= blocked ? (long long)ca.blk_extent : (long long)a.extent[d];
        if (last) break;
    }
    const uint32_t last_row = ca.blk_extent - 1 - row0;  // rows past the end re-read the last valid row
    float4 inv[NIN], var[NIN][U];
#pragma unroll
    for (int i = 0; i < NIN; i++) {
        if ((INV >> i) & 1) {
            inv[i] = __ldg(reinterpret_cast<const float4*>(a.in[i] + off[i]));
        } else {
#pragma unroll
            for (int u = 0; u < U; u++)
                var[i][u] = __ldg(reinterpret_cast<const float4*>(a.in[i] + off[i] + (long long)min((uint32_t)u, last_row) * ca.blk_in[i]));
        }
    }
    float acc[U][4];
#pragma unroll
    for (int u = 0; u < U; u++) f4_to_lanes((INV & 1) ? inv[0] : var[0][u], acc[u]);
    int next = 1;
    for (int s = 0; s < ca.n_steps; s++) {
        const int kind = ca.kind[s], op = ca.op[s];
        if (kind == CH_ACC_IN || kind == CH_IN_ACC) {
            if (NIN > 1 && next == 1) chain_apply_operand<NIN, INV, (NIN > 1 ? 1 : 0)>(acc, inv, var, kind, op);
            if (NIN > 2 && next == 2) chain_apply_operand<NIN, INV, (NIN > 2 ? 2 : 0)>(acc, inv, var, kind, op);
            if (NIN > 3 && next == 3) chain_apply_operand<NIN, INV, (NIN > 3 ? 3 : 0)>(acc, inv, var, kind, op);
            next++;
        } else {
            float other[U][4];
#pragma unroll
            for (int u = 0; u < U; u++)
#pragma unroll
                for (int c = 0; c < 4; c++) other[u][c] = ca.scalar[s];
            chain_step<false, U * 4>(&acc[0][0], &other[0][0], kind, op);
        }
    }
    bool any_nan = false;
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
        for (int c = 0; c < 4; c++) any_nan |= acc[u][c] != acc[u][c];
    if (any_nan) atomicOr(ca.nan_flag, 1);
#pragma unroll
    for (int u = 0; u < U; u++) {
        if ((uint32_t)u <= last_row) {
            float4* dst = reinterpret_cast<float4*>(a.out + ooff + u * ca.blk_out);
            const float4 r = make_float4(acc[u][0], acc[u][1], acc[u][2], acc[u][3]);
            if (a.stream_store) __stcs(dst, r);
            else *dst = r;
        }
    }
}

template <int NIN, int INV>
static void launch_blocked_inv(const ChainArgs<NIN>& ca, unsigned blocks, int inv) {
    if constexpr (INV < 0) {
        (void)ca, (void)blocks, (void)inv;
    } else {
        if (inv == INV) ew_chain_blocked_kernel<NIN, INV><<<blocks, kEwThreads, 0, rt().stream>>>(ca);
        else launch_blocked_inv<NIN, INV - 1>(ca, blocks, inv);
    }
}

// ---- FP64 evaluation: compute in double precision and round to f32 ONCE at the end ------------------------
// This is the arithmetic of the reference's `apply(fn)`: cffi hands the Python callback a C float
// widened to a Python float (a double), the lambda computes in double, and the returned double is
// narrowed to f32 (/root/reference/arraylib/core.py:455-460, arraylib.c:638-645). A traced lambda
// therefore matches the reference bit for bit for + - * / comparisons, neg, abs, sqrt, ceil, floor (all
// correctly rounded in IEEE double) and to the usual libm tolerance for the transcendentals.
__device__ __forceinline__ double dbin(int op, double x, double y) {
    switch (op) {
        case AL_SUM: return __dadd_rn(x, y);
        case AL_SUB: return __dsub_rn(x, y);
        case AL_MUL: return __dmul_rn(x, y);
        case AL_DIV: return __ddiv_rn(x, y);
        case AL_MOD: return fmod(x, y);
        case AL_POW: return pow(x, y);
        case AL_MAX: return (x >= y || y != y) ? x : y;
        case AL_MIN: return (x <= y || y != y) ? x : y;
        case AL_EQ: return x == y ? 1.0 : 0.0;
        case AL_NE: return x != y ? 1.0 : 0.0;
        case AL_GT: return x > y ? 1.0 : 0.0;
        case AL_GE: return x >= y ? 1.0 : 0.0;
        case AL_LT: return x < y ? 1.0 : 0.0;
        default: return x <= y ? 1.0 : 0.0;
    }
}
__device__ __forceinline__ double dun(int op, double x) {
    switch (op) {
        case AL_NEG: return -x;
        case AL_ABS: return fabs(x);
        case AL_SQRT: return __dsqrt_rn(x);
        case AL_EXP: return exp(x);
        case AL_LOG: return log(x);
        case AL_SIN: return sin(x);
        case AL_COS: return cos(x);
        case AL_TAN: return tan(x);
        case AL_ASIN: return asin(x);
        case AL_ACOS: return acos(x);
        case AL_ATAN: return atan(x);
        case AL_SINH: return sinh(x);
        case AL_COSH: return cosh(x);
        case AL_TANH: return tanh(x);
        case AL_ASINH: return asinh(x);
        case AL_ACOSH: return acosh(x);
        case AL_ATANH: return atanh(x);
        case AL_CEIL: return ceil(x);
        default: return floor(x);
    }
}

// ---- FP64 tape: a small register machine for traced apply() lambdas ---------------------------------------
constexpr int kTapeMax = 32;
constexpr int kTapeConst = -100;  // operand source: the instruction's own constant

struct TapeArgs {
    float* out;
    const float* in[kChainMaxInputs];
    long long stride[kChainMaxInputs][kMaxDims];
    uint32_t n_items;
    int ndim, n_in, n_insn, result;
    uint32_t extent[kMaxDims];
    FastDiv div[kMaxDims];
    unsigned char kind[kTapeMax], op[kTapeMax];
    signed char a_src[kTapeMax], b_src[kTapeMax];
    double a_const[kTapeMax], b_const[kTapeMax];
};

#pragma nv_diag_suppress 549  // r[] is written by instruction s before any later instruction may read it (validated on the host)
__global__ void __launch_bounds__(kEwThreads) ew_tape_f64_kernel(const __grid_constant__ TapeArgs t) {
    const uint32_t item = blockIdx.x * (uint32_t)kEwThreads + threadIdx.x;
    if (item >= t.n_items) return;
    long long off[kChainMaxInputs] = {0, 0, 0, 0};
    uint32_t rem = item;
    for (int d = 0; d < t.ndim; d++) {
        uint32_t idx;
        if (d == t.ndim - 1) {
            idx = rem;
        } else {
            const uint32_t q = fdiv(rem, t.div[d]);
            idx = rem - q * t.extent[d];
            rem = q;
        }
#pragma unroll
        for (int i = 0; i < kChainMaxInputs; i++) off[i] += (long long)idx * t.stride[i][d];
    }
    double in[kChainMaxInputs];
#pragma unroll
    for (int i = 0; i < kChainMaxInputs; i++) in[i] = i < t.n_in ? (double)__ldg(t.in[i] + off[i]) : 0.0;
    double r[kTapeMax];  // register file of the tape (dynamic indexing -> L1-resident local memory)
    auto fetch = [&](int src, double c) -> double {
        if (src >= 0) return r[src];
        if (src == kTapeConst) return c;
        const int i = -1 - src;
        return i == 0 ? in[0] : i == 1 ? in[1] : i == 2 ? in[2] : in[3];
    };
    for (int s = 0; s < t.n_insn; s++) {
        const double a = fetch(t.a_src[s], t.a_const[s]);
        r[s] = t.kind[s] ? dun(t.op[s], a) : dbin(t.op[s], a, fetch(t.b_src[s], t.b_const[s]));
    }
    t.out[item] = (float)r[t.result];
}

// ---- NaN repair pass behind every chain launch --------------------------------------------------------------
struct ChainFixArgs {
    float* out;
    const float* in[kChainMaxInputs];
    int ndim, nin;
    unsigned long long extent[kMaxDims];        // innermost first, in ELEMENTS (the un-vectorised, un-blocked space)
    long long stride[kChainMaxInputs][kMaxDims];
    unsigned long long total;
    int n_steps;
    unsigned char kind[kChainMaxSteps], op[kChainMaxSteps];
    float scalar[kChainMaxSteps];
    int* flag;     // [0] = some output of the chain launch is a NaN, [1] = CTAs of this pass that have finished
};

__global__ void __launch_bounds__(256) chain_nan_fix_kernel(const __grid_constant__ ChainFixArgs f) {
    __shared__ int go;
    if (threadIdx.x == 0) go = *(volatile int*)f.flag;
    __syncthreads();
    if (go) {  // some element is a NaN: find them (one pass over the output) and replay each from its operands
        const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
        for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < f.total; e += step) {
            const float v = f.out[e];
            if (v == v) continue;
            long long off[kChainMaxInputs] = {0, 0, 0, 0};
            unsigned long long rem = e;
            for (int d = 0; d < f.ndim; d++) {
                const unsigned long long idx = d == f.ndim - 1 ? rem : rem % f.extent[d];
                rem /= f.extent[d];
#pragma unroll
                for (int i = 0; i < kChainMaxInputs; i++) off[i] += (long long)idx * f.stride[i][d];
            }
            float x[kChainMaxInputs];
#pragma unroll
            for (int i = 0; i < kChainMaxInputs; i++) x[i] = i < f.nin ? f.in[i][off[i]] : 0.f;
            f.out[e] = chain_exact(f, x[0], x[1], x[2], x[3]);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // every CTA has read the flag before it checks in; the last one lowers it for the next chain
        __threadfence();
        if (atomicAdd(f.flag + 1, 1) == (int)gridDim.x - 1) {
            f.flag[0] = 0;
            f.flag[1] = 0;
            __threadfence();
        }
    }
}

static int* g_chain_nan_flag = nullptr;
static int* chain_nan_flag() {
    if (!g_chain_nan_flag) {
        if (cudaMalloc(&g_chain_nan_flag, 2 * sizeof(int)) != cudaSuccess || cudaMemsetAsync(g_chain_nan_flag, 0, 2 * sizeof(int), rt().stream) != cudaSuccess) {
            cudaGetLastError();
            g_chain_nan_flag = nullptr;
        }
    }
    return g_chain_nan_flag;
}

struct ChainPlan {
    int n_steps;
    unsigned char kind[kChainMaxSteps], op[kChainMaxSteps];
    float scalar[kChainMaxSteps];
    bool full;
};

template <int VEC, int NIN>
static bool launch_chain_one(const ChainPlan& plan, const Folded& fo_in, const float* const* in, float* out, uint64_t) {
    // A flat contiguous stream has no outer axis to block over: view [n] as [rows, n/rows] so that the
    // row-blocked kernel (one op decode per 8 vectors) applies; lanes stay coalesced along n/rows.
    Folded fo = fo_in;
    if (VEC == 4 && fo.ndim == 1 && !plan.full && rt().chain_block >= 2 && NIN <= 2 &&
        fo.extent[0] % (4 * kChainRows) == 0 && fo.extent[0] >= (1u << 16)) {
        bool ok = true;
        for (int i = 0; i < NIN; i++) ok = ok && (fo.in_stride[i][0] == 1);
        if (ok) {
            const uint64_t cols = fo.extent[0] / kChainRows;
            fo.ndim = 2;
            fo.extent[0] = cols;
            fo.extent[1] = kChainRows;
            fo.out_stride[1] = (int64_t)cols;
            for (int i = 0; i < NIN; i++) fo.in_stride[i][1] = (int64_t)cols;
        }
    }
    ChainArgs<NIN> ca;
    memset(&ca, 0, sizeof ca);
    NdArgs<NIN>& a = ca.nd;
    a.out = out;
    a.ndim = fo.ndim;
    // Block along an outer axis with >= kChainBlock rows; prefer one along which some operand is
    // invariant (stride 0), since that operand is then loaded once per block.
    bool any_inner_bcast = false;
    for (int i = 0; i < NIN; i++) any_inner_bcast |= fo.in_stride[i][0] == 0;
    const bool special = VEC == 4 && !plan.full && !any_inner_bcast && rt().chain_block >= 2;
    const int rows = special ? kChainRows : kChainBlock;
    // Prefer the OUTERMOST axis along which some operand is invariant: the invariant operand is then
    // loaded once per block, and the operands that DO vary along it are re-read by the CTAs that follow
    // (they differ only in the faster axes), i.e. they hit in L1 instead of L2. On C3 (A[i], B[j], c)
    // blocking over i instead of j moved the chain from 3.2 ms to the value in DESIGN.md.
    int blk = -1;
    for (int d = fo.ndim - 1; d >= 1 && rt().chain_block; d--) {
        if (fo.extent[d] < (uint64_t)rows) continue;
        bool invariant = false;
        for (int i = 0; i < NIN; i++) invariant |= fo.in_stride[i][d] == 0;
        if (blk < 0 || invariant) blk = d;
        if (invariant) break;
    }
    ca.blk_axis = blk;
    uint64_t n_items = 1, total = 1;
    long long run = 1;
    for (int d = 0; d < fo.ndim; d++) {
        uint64_t e = fo.extent[d];
        if (d == blk) {
            ca.blk_extent = (uint32_t)e;
            ca.blk_out = run;
            for (int i = 0; i < NIN; i++) ca.blk_in[i] = fo.in_stride[i][d];
            e = (e + rows - 1) / rows;
        }
        if (d == 0) e /= VEC;
        a.extent[d] = (uint32_t)e;
        a.div[d] = make_fastdiv((uint32_t)e);
        n_items *= e;
        run *= (long long)fo.extent[d];
        total *= fo.extent[d];
    }
    a.n_items = (uint32_t)n_items;
    a.stream_store = total * sizeof(float) >= (uint64_t)rt().stream_store_min_bytes;
    a.load_policy = (unsigned char)rt().load_policy;
    for (int i = 0; i < NIN; i++) {
        a.in[i] = in[i];
        bool reuse = false;
        for (int d = 0; d < fo.ndim; d++) {
            a.stride[i][d] = fo.in_stride[i][d] * (d == 0 ? VEC : 1) * (d == blk ? rows : 1);
            if (fo.in_stride[i][d] == 0 && fo.extent[d] > 1) reuse = true;
        }
        a.inner_bcast[i] = fo.in_stride[i][0] == 0;
        a.reuse[i] = reuse;
    }
    ca.nan_flag = chain_nan_flag();
    AL_REQUIRE(ca.nan_flag, "MemoryError: cannot allocate the chain NaN flag");
    ca.n_steps = plan.n_steps;
    memcpy(ca.kind, plan.kind, sizeof ca.kind);
    memcpy(ca.op, plan.op, sizeof ca.op);
    memcpy(ca.scalar, plan.scalar, sizeof ca.scalar);
    const unsigned blocks = (unsigned)((n_items + kEwThreads - 1) / kEwThreads);
    cudaStream_t st = rt().stream;
    if (blk >= 0 && special) {
        int inv = 0;
        for (int i = 0; i < NIN; i++) inv |= (ca.blk_in[i] == 0) << i;
        launch_blocked_inv<NIN, (1 << NIN) - 1>(ca, blocks, inv);
        note_launch("ew_chain_vec4_rows8");
        return check_launch("ew_chain_blocked");
    }
    if (blk >= 0) {
        if (plan.full) ew_chain_kernel<VEC, NIN, true, kChainBlock><<<blocks, kEwThreads, 0, st>>>(ca);
        else ew_chain_kernel<VEC, NIN, false, kChainBlock><<<blocks, kEwThreads, 0, st>>>(ca);
    } else {
        if (plan.full) ew_chain_kernel<VEC, NIN, true, 1><<<blocks, kEwThreads, 0, st>>>(ca);
        else ew_chain_kernel<VEC, NIN, false, 1><<<blocks, kEwThreads, 0, st>>>(ca);
    }
    note_launch(VEC == 4 ? (blk >= 0 ? "ew_chain_vec4_blocked" : "ew_chain_vec4") : "ew_chain_scalar");
    return check_launch("ew_chain");
}

template <int NIN>
static bool launch_nan_fix(const ChainPlan& plan, const Folded& fo, const float* const* in, float* out) {
    ChainFixArgs f;
    memset(&f, 0, sizeof f);
    f.out = out;
    f.ndim = fo.ndim;
    f.nin = NIN;
    f.total = 1;
    for (int d = 0; d < fo.ndim; d++) {
        f.extent[d] = fo.extent[d];
        f.total *= fo.extent[d];
        for (int i = 0; i < NIN; i++) f.stride[i][d] = fo.in_stride[i][d];
    }
    for (int i = 0; i < NIN; i++) f.in[i] = in[i];
    f.n_steps = plan.n_steps;
    memcpy(f.kind, plan.kind, sizeof f.kind);
    memcpy(f.op, plan.op, sizeof f.op);
    memcpy(f.scalar, plan.scalar, sizeof f.scalar);
    f.flag = chain_nan_flag();
    const unsigned long long want = (f.total + 255) / 256;
    const unsigned blocks = (unsigned)std::min<unsigned long long>(want, (unsigned long long)kNumSMs * 4);
    chain_nan_fix_kernel<<<blocks, 256, 0, rt().stream>>>(f);
    rt().launches++;  // counted; al_last_kernel() keeps naming the chain kernel
    return check_launch("chain_nan_fix");
}

template <int VEC, int NIN>
static bool launch_chain(const ChainPlan& plan, Folded fo, const float* const* in_ptrs, float* out) {
    uint64_t total = 1;
    for (int d = 0; d < fo.ndim; d++) total *= fo.extent[d];
    const uint64_t n_items = total / VEC;
    if (n_items < (1ull << 31))
        return launch_chain_one<VEC, NIN>(plan, fo, in_ptrs, out, n_items) && launch_nan_fix<NIN>(plan, fo, in_ptrs, out);
    const int top = fo.ndim - 1;  // bisect the outermost axis, like launch_nd
    const uint64_t e = fo.extent[top];
    uint64_t a_ext = e / 2;
    if (top == 0) a_ext &= ~(uint64_t)(VEC - 1);
    const uint64_t inner = total / e;
    Folded lo = fo, hi = fo;
    lo.extent[top] = a_ext;
    hi.extent[top] = e - a_ext;
    const float* in_hi[NIN];
    for (int i = 0; i < NIN; i++) in_hi[i] = in_ptrs[i] + (int64_t)a_ext * fo.in_stride[i][top];
    return launch_chain<VEC, NIN>(plan, lo, in_ptrs, out) && launch_chain<VEC, NIN>(plan, hi, in_hi, out + a_ext * inner);
}

template <int NIN>
static bool run_chain(const ChainPlan& plan, const Folded& fo, const float* const* in, float* out) {
    bool vec_ok = !rt().force_general && ((uintptr_t)out % 16 == 0) && fo.extent[0] % 4 == 0;
    for (int i = 0; i < NIN && vec_ok; i++) {
        const int64_t s0 = fo.in_stride[i][0];
        if (s0 == 0) continue;
        if (s0 != 1 || (uintptr_t)in[i] % 16 != 0) vec_ok = false;
        for (int d = 1; d < fo.ndim && vec_ok; d++)
            if (fo.in_stride[i][d] % 4 != 0) vec_ok = false;
    }
    return vec_ok ? launch_chain<4, NIN>(plan, fo, in, out) : launch_chain<1, NIN>(plan, fo, in, out);
}

// Right-aligned NumPy broadcast of n layouts (arraylib.c:171-206 generalised from two operands): fills the
// common shape and, per operand, strides that are 0 wherever the operand is broadcast.
static bool broadcast_all(const NDArray* const* inputs, size_t n_inputs, size_t nd, size_t* shape,
                          size_t (*strides)[64]) {
    for (size_t k = 0; k < nd; k++) {  // k counts axes from the right
        size_t e = 1;
        for (size_t i = 0; i < n_inputs; i++) {
            const Layout* l = inputs[i]->lay;
            const size_t ei = k < l->ndim ? l->shape[l->ndim - 1 - k] : 1;
            AL_REQUIRE(ei == e || ei == 1 || e == 1, "ValueError: can not broadcast.");
            if (ei > e) e = ei;
        }
        shape[nd - 1 - k] = e;
        for (size_t i = 0; i < n_inputs; i++) {
            const Layout* l = inputs[i]->lay;
            const bool present = k < l->ndim && l->shape[l->ndim - 1 - k] == e;
            strides[i][nd - 1 - k] = present ? l->stride[l->ndim - 1 - k] : 0;
        }
    }
    return true;
}

}  // namespace al

using namespace al;

extern "C" NDArray* array_chain_eval(const NDArray* const* inputs, size_t n_inputs, const al_chain_step* steps,
                                     size_t n_steps) {
    AL_ENTER;
    AL_REQUIRE(inputs && n_inputs >= 1 && n_inputs <= (size_t)kChainMaxInputs, "ValueError: a fused chain takes 1..%d arrays", kChainMaxInputs);
    AL_REQUIRE(n_steps <= (size_t)kChainMaxSteps && (steps || n_steps == 0), "ValueError: a fused chain holds at most %d steps", kChainMaxSteps);
    size_t nd = 0;
    for (size_t i = 0; i < n_inputs; i++) {
        if (!valid(inputs[i], "array_chain_eval")) return nullptr;
        if (inputs[i]->lay->ndim > nd) nd = inputs[i]->lay->ndim;
    }
    AL_REQUIRE(nd <= 64, "ValueError: too many dimensions");
    size_t shape[64], strides[kChainMaxInputs][64];
    if (!broadcast_all(inputs, n_inputs, nd, shape, strides)) return nullptr;
    ChainPlan plan;
    memset(&plan, 0, sizeof plan);
    plan.n_steps = (int)n_steps;
    size_t consumed = 1;
    for (size_t s = 0; s < n_steps; s++) {
        const int kind = steps[s].kind, op = steps[s].op;
        AL_REQUIRE(kind >= CH_ACC_IN && kind <= CH_UNARY, "ValueError: unknown chain step kind %d", kind);
        if (kind == CH_UNARY) {
            AL_REQUIRE(op >= 0 && op < AL_UFUNC_COUNT, "ValueError: unknown ufunc %d", op);
            if (!(op == AL_NEG || op == AL_ABS || op == AL_SQRT || op == AL_CEIL || op == AL_FLOOR)) plan.full = true;
        } else {
            AL_REQUIRE(op >= 0 && op < AL_BINOP_COUNT, "ValueError: unknown binary op %d", op);
            if (op == AL_MOD || op == AL_POW) plan.full = true;
            if (kind == CH_ACC_IN || kind == CH_IN_ACC) consumed++;
        }
        plan.kind[s] = (unsigned char)kind;
        plan.op[s] = (unsigned char)op;
        plan.scalar[s] = steps[s].scalar;
    }
    AL_REQUIRE(consumed == n_inputs, "ValueError: the chain consumes %zu arrays but %zu were given", consumed, n_inputs);
    NDArray* out = new_array(shape, nd);
    if (!out) return nullptr;
    Folded fo;
    const size_t* sp[kChainMaxInputs];
    const float* in[kChainMaxInputs];
    for (size_t i = 0; i < n_inputs; i++) sp[i] = strides[i], in[i] = inputs[i]->ptr;
    bool ok = fold_layout(nd, shape, out->lay->stride, (int)n_inputs, sp, &fo);
    if (ok) {
        switch (n_inputs) {
            case 1: ok = run_chain<1>(plan, fo, in, out->ptr); break;
            case 2: ok = run_chain<2>(plan, fo, in, out->ptr); break;
            case 3: ok = run_chain<3>(plan, fo, in, out->ptr); break;
            default: ok = run_chain<4>(plan, fo, in, out->ptr); break;
        }
    }
    if (!ok) {
        std::string keep = last_error_string();
        release_array(out);
        restore_error(keep);
        return nullptr;
    }
    return out;
}

extern "C" NDArray* array_tape_eval_f64(const NDArray* const* inputs, size_t n_inputs, const al_tape_insn* tape,
                                        size_t n_insn, size_t result) {
    AL_ENTER;
    AL_REQUIRE(inputs && n_inputs >= 1 && n_inputs <= (size_t)kChainMaxInputs, "ValueError: a tape takes 1..%d arrays", kChainMaxInputs);
    AL_REQUIRE(tape && n_insn >= 1 && n_insn <= (size_t)kTapeMax, "NotImplementedError: a tape holds 1..%d instructions", kTapeMax);
    AL_REQUIRE(result < n_insn, "ValueError: tape result index out of range");
    size_t nd = 0;
    for (size_t i = 0; i < n_inputs; i++) {
        if (!valid(inputs[i], "array_tape_eval_f64")) return nullptr;
        if (inputs[i]->lay->ndim > nd) nd = inputs[i]->lay->ndim;
    }
    AL_REQUIRE(nd <= 64, "ValueError: too many dimensions");
    size_t shape[64], strides[kChainMaxInputs][64];
    if (!broadcast_all(inputs, n_inputs, nd, shape, strides)) return nullptr;
    TapeArgs t;
    memset(&t, 0, sizeof t);
    for (size_t s = 0; s < n_insn; s++) {
        const al_tape_insn& in = tape[s];
        AL_REQUIRE(in.kind == 0 || in.kind == 1, "ValueError: unknown tape instruction kind %d", in.kind);
        AL_REQUIRE(in.op >= 0 && in.op < (in.kind ? (int)AL_UFUNC_COUNT : (int)AL_BINOP_COUNT), "ValueError: unknown tape op %d", in.op);
        const int srcs[2] = {in.a_src, in.kind ? kTapeConst : in.b_src};
        for (int src : srcs)
            AL_REQUIRE(src == kTapeConst || (src >= 0 && (size_t)src < s) || (src < 0 && src >= -(int)n_inputs),
                       "ValueError: tape instruction %zu reads an undefined operand %d", s, src);
        t.kind[s] = (unsigned char)in.kind;
        t.op[s] = (unsigned char)in.op;
        t.a_src[s] = (signed char)in.a_src;
        t.b_src[s] = (signed char)(in.kind ? kTapeConst : in.b_src);
        t.a_const[s] = in.a_const;
        t.b_const[s] = in.b_const;
    }
    NDArray* out = new_array(shape, nd);
    if (!out) return nullptr;
    Folded fo;
    const size_t* sp[kChainMaxInputs];
    for (size_t i = 0; i < n_inputs; i++) sp[i] = strides[i];
    bool ok = fold_layout(nd, shape, out->lay->stride, (int)n_inputs, sp, &fo);
    uint64_t total = 1;
    if (ok) {
        for (int d = 0; d < fo.ndim; d++) total *= fo.extent[d];
        if (total >= (1ull << 31)) {
            set_error("NotImplementedError: FP64 tapes are limited to 2^31-1 elements per call");
            ok = false;
        }
    }
    if (ok) {
        t.out = out->ptr;
        t.n_items = (uint32_t)total;
        t.ndim = fo.ndim;
        t.n_in = (int)n_inputs;
        t.n_insn = (int)n_insn;
        t.result = (int)result;
        for (int d = 0; d < fo.ndim; d++) {
            t.extent[d] = (uint32_t)fo.extent[d];
            t.div[d] = make_fastdiv(t.extent[d]);
            for (size_t i = 0; i < n_inputs; i++) t.stride[i][d] = fo.in_stride[i][d];
        }
        for (size_t i = 0; i < n_inputs; i++) t.in[i] = inputs[i]->ptr;
        ew_tape_f64_kernel<<<(unsigned)((total + kEwThreads - 1) / kEwThreads), kEwThreads, 0, rt().stream>>>(t);
        note_launch("ew_tape_f64");
        ok = check_launch("ew_tape_f64");
    }
    if (!ok) {
        std::string keep = last_error_string();
        release_array(out);
        restore_error(keep);
        return nullptr;
    }
    return out;
}
