// Elementwise engine front end: copy / where instantiations, the strided-destination kernel and the
// extern "C" entry points. The kernels and launch plans live in elementwise.cuh; the binary, scalar
// and unary functor tables are instantiated in ew_binary.cu / ew_scalar.cu / ew_unary.cu.
#include "elementwise.cuh"

namespace al {

bool launch_ew_binary(const EwCall& c, const Folded& fo);
bool launch_ew_scalar(const EwCall& c, const Folded& fo);
bool launch_ew_unary(const EwCall& c, const Folded& fo);

bool launch_elementwise(const EwCall& c, const Folded& fo) {
    switch (c.kind) {
        case EW_BINARY: return launch_ew_binary(c, fo);
        case EW_SCALAR: return launch_ew_scalar(c, fo);
        case EW_UNARY: return launch_ew_unary(c, fo);
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
        release_array(out);
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
    if (!valid_mut(dst, "array_set_view_from_scalar")) return nullptr;
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
    if (!valid_mut(dst, "array_set_view_from_array") || !valid(src, "array_set_view_from_array")) return nullptr;
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
