/*
 * oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C CPU restatement of the reference's array-core hot path, written against flat
 * (pointer, shape[], stride[]) arguments instead of the reference's NDArray structs. Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it. The product (arraylib_b200) never does.
 *
 * Parity status: PINNED. tests/test_oracle.py checks every function below against
 *   (1) the reference's own known-answer cases (tests/test_array.py:73-77,87-103,158-171,
 *       222-252,318-334 restated in tests/golden/), and
 *   (2) the unmodified reference translation unit compiled into oracle/_ref/ (see Makefile),
 *       differentially on random layouts, and golden vectors generated from it
 *       (tests/golden/make_golden.py).
 *
 * Strides are in ELEMENTS and unsigned, like the reference (arraylib.h:63-67).
 * Each function cites the reference lines it restates (paths relative to /root/reference).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

/* op numbering shared with include/arraylib_b200.h */
enum { B_SUM, B_SUB, B_MUL, B_DIV, B_MOD, B_POW, B_MAX, B_MIN, B_EQ, B_NE, B_GT, B_GE, B_LT, B_LE };
enum { U_NEG, U_ABS, U_SQRT, U_EXP, U_LOG, U_SIN, U_COS, U_TAN, U_ASIN, U_ACOS, U_ATAN, U_SINH,
       U_COSH, U_TANH, U_ASINH, U_ACOSH, U_ATANH, U_CEIL, U_FLOOR };
enum { R_SUM, R_MAX, R_MIN, R_MUL };

/* arraylib/src/arraylib.c:12-25 -- scalar binary functions; mod/pow/max/min go through the
 * double-precision libm entry points and round back to f32. */
static float bin_apply(int op, float l, float r) {
    switch (op) {
        case B_SUM: return l + r;
        case B_SUB: return l - r;
        case B_MUL: return l * r;
        case B_DIV: return l / r;
        case B_MOD: return (float)fmod((double)l, (double)r);
        case B_POW: return (float)pow((double)l, (double)r);
        case B_MAX: return (float)fmax((double)l, (double)r);
        case B_MIN: return (float)fmin((double)l, (double)r);
        case B_EQ: return l == r;
        case B_NE: return l != r;
        case B_GT: return l > r;
        case B_GE: return l >= r;
        case B_LT: return l < r;
        case B_LE: return l <= r;
    }
    return NAN;
}

/* arraylib/src/arraylib.c:31-49 -- unary functions: f32 -> double libm -> f32. */
static float un_apply(int op, float v) {
    double x = (double)v;
    switch (op) {
        case U_NEG: return -v;
        case U_ABS: return (float)fabs(x);
        case U_SQRT: return (float)sqrt(x);
        case U_EXP: return (float)exp(x);
        case U_LOG: return (float)log(x);
        case U_SIN: return (float)sin(x);
        case U_COS: return (float)cos(x);
        case U_TAN: return (float)tan(x);
        case U_ASIN: return (float)asin(x);
        case U_ACOS: return (float)acos(x);
        case U_ATAN: return (float)atan(x);
        case U_SINH: return (float)sinh(x);
        case U_COSH: return (float)cosh(x);
        case U_TANH: return (float)tanh(x);
        case U_ASINH: return (float)asinh(x);
        case U_ACOSH: return (float)acosh(x);
        case U_ATANH: return (float)atanh(x);
        case U_CEIL: return (float)ceil(x);
        case U_FLOOR: return (float)floor(x);
    }
    return NAN;
}

/* arraylib/src/arraylib.c:116-123 -- logical linear index -> element offset, peeling the
 * innermost dimension first. */
size_t orc_offset(size_t index, size_t ndim, const size_t* shape, const size_t* stride) {
    size_t off = 0;
    for (size_t d = ndim; d-- > 0;) {
        off += (index % shape[d]) * stride[d];
        index /= shape[d];
    }
    return off;
}

/* arraylib/src/arraylib.c:66-72 -- element count; a 0 extent counts as 1. */
size_t orc_count(const size_t* shape, size_t ndim) {
    size_t n = 1;
    for (size_t d = 0; d < ndim; d++) n *= shape[d] ? shape[d] : 1;
    return n;
}

/* arraylib/src/arraylib.c:171-178 -- right-aligned broadcast compatibility. */
int orc_broadcastable(const size_t* sa, size_t na, const size_t* sb, size_t nb) {
    while (na > 0 && nb > 0) {
        na--, nb--;
        if (sa[na] != sb[nb] && sa[na] != 1 && sb[nb] != 1) return 0;
    }
    return 1;
}

/* arraylib/src/arraylib.c:180-206 -- broadcast two layouts to a common rank. Output extents are
 * the per-axis maximum; an operand gets stride 0 on an axis it lacks or where its extent differs
 * from that maximum. out_shape/out_sa/out_sb must hold max(na, nb) entries. Returns the rank. */
size_t orc_broadcast(const size_t* sha, const size_t* sta, size_t na, const size_t* shb,
                     const size_t* stb, size_t nb, size_t* out_shape, size_t* out_sa,
                     size_t* out_sb) {
    size_t n = na > nb ? na : nb;
    for (size_t k = 0; k < n; k++) { /* k counts axes from the right */
        size_t o = n - 1 - k;
        size_t ea = k < na ? sha[na - 1 - k] : 1;
        size_t eb = k < nb ? shb[nb - 1 - k] : 1;
        size_t e = ea > eb ? ea : eb;
        out_shape[o] = e;
        out_sa[o] = (k < na && ea == e) ? sta[na - 1 - k] : 0;
        out_sb[o] = (k < nb && eb == e) ? stb[nb - 1 - k] : 0;
    }
    return n;
}

/* arraylib/src/arraylib.c:569-589 -- out[i] = fn(a[off_a(i)], b[off_b(i)]) over an already
 * broadcast shape; out is contiguous row-major. */
void orc_binary(int op, const float* a, const size_t* sa, const float* b, const size_t* sb,
                const size_t* shape, size_t ndim, float* out) {
    size_t n = orc_count(shape, ndim);
#pragma omp parallel for
    for (size_t i = 0; i < n; i++)
        out[i] = bin_apply(op, a[orc_offset(i, ndim, shape, sa)], b[orc_offset(i, ndim, shape, sb)]);
}

/* arraylib/src/arraylib.c:491-501 -- array (strided) o scalar -> contiguous. Serial in the
 * reference even under OpenMP (:494). */
void orc_binary_scalar(int op, const float* a, const size_t* sa, const size_t* shape, size_t ndim,
                       float rhs, float* out) {
    size_t n = orc_count(shape, ndim);
    for (size_t i = 0; i < n; i++) out[i] = bin_apply(op, a[orc_offset(i, ndim, shape, sa)], rhs);
}

/* arraylib/src/arraylib.c:638-645 -- strided unary map -> contiguous. */
void orc_unary(int op, const float* a, const size_t* sa, const size_t* shape, size_t ndim,
               float* out) {
    size_t n = orc_count(shape, ndim);
#pragma omp parallel for
    for (size_t i = 0; i < n; i++) out[i] = un_apply(op, a[orc_offset(i, ndim, shape, sa)]);
}

/* arraylib/src/arraylib.c:248-256 -- gather a strided view into a contiguous buffer. */
void orc_gather(const float* a, const size_t* sa, const size_t* shape, size_t ndim, float* out) {
    size_t n = orc_count(shape, ndim);
#pragma omp parallel for
    for (size_t i = 0; i < n; i++) out[i] = a[orc_offset(i, ndim, shape, sa)];
}

/* arraylib/src/arraylib.c:690-756 (serial build) -- keepdims reduction. One pass over the source
 * in ascending logical order, accumulating in f32 into the output slot found by zeroing the
 * reduced axes' strides. The accumulator starts at `init`, and the reference folds `init` in a
 * second time when it merges the (single) thread-local accumulator into dst (:746-750).
 * out must hold prod(kept extents) floats. */
void orc_reduce(int op, const float* a, const size_t* sa, const size_t* shape, size_t ndim,
                const uint8_t* reduced, float init, float* out) {
    size_t dshape[64], dstride[64], scatter[64];
    for (size_t d = 0; d < ndim; d++) dshape[d] = reduced[d] ? 1 : shape[d];
    size_t run = 1;
    for (size_t d = ndim; d-- > 0;) {
        dstride[d] = run;
        run *= dshape[d];
    }
    for (size_t d = 0; d < ndim; d++) scatter[d] = reduced[d] ? 0 : dstride[d];
    size_t nout = orc_count(dshape, ndim), nsrc = orc_count(shape, ndim);
    int bop = op == R_SUM ? B_SUM : op == R_MAX ? B_MAX : op == R_MIN ? B_MIN : B_MUL;
    for (size_t i = 0; i < nout; i++) out[i] = init;
    for (size_t i = 0; i < nsrc; i++) {
        size_t so = orc_offset(i, ndim, shape, sa);
        size_t dofs = orc_offset(i, ndim, shape, scatter);
        out[dofs] = bin_apply(bop, out[dofs], a[so]);
    }
    for (size_t i = 0; i < nout; i++) out[i] = bin_apply(bop, init, out[i]);
}

/* arraylib/src/arraylib.c:298-308 -- arange. The element count comes from f32-rounded
 * (end - start) and step, both truncated to size_t; the running value is a size_t that is pushed
 * through an f32 add every step, so it stops being exact at 2^24. Returns the count; writes when
 * out != NULL. */
size_t orc_arange(float start, float end, float step, float* out) {
    size_t span = (size_t)(end - start), st = (size_t)step;
    size_t total = (span + st - 1) / st;
    if (!out) return total;
    size_t running = (size_t)start;
    for (size_t i = 0; i < total; i++) {
        out[i] = (float)running;
        running = (size_t)((float)running + step);
    }
    return total;
}

/* arraylib/src/arraylib.c:310-316 -- linspace = start + arange(0, n) * dx. The reference build
 * (gcc, -ffp-contract=fast, FMA available) fuses the multiply-add, hence fmaf here. */
size_t orc_linspace(float start, float end, float n, float* out) {
    float dx = (end - start) / (n - 1);
    size_t total = orc_arange(0.0f, n, 1.0f, out);
    if (out)
        for (size_t i = 0; i < total; i++) out[i] = fmaf(out[i], dx, start);
    return total;
}

/* arraylib/src/arraylib.c:774-838 -- where: f32 truthiness (NaN true, +-0 false), no
 * broadcasting. A NULL lhs/rhs pointer means "use the scalar". */
void orc_where(const float* c, const size_t* sc, const float* l, const size_t* sl, float lscalar,
               const float* r, const size_t* sr, float rscalar, const size_t* shape, size_t ndim,
               float* out) {
    size_t n = orc_count(shape, ndim);
    for (size_t i = 0; i < n; i++) {
        float cv = c[orc_offset(i, ndim, shape, sc)];
        float lv = l ? l[orc_offset(i, ndim, shape, sl)] : lscalar;
        float rv = r ? r[orc_offset(i, ndim, shape, sr)] : rscalar;
        out[i] = cv ? lv : rv;
    }
}

/* arraylib/src/arraylib.c:522-564 -- 2-D matmul; for a fixed (i, j) the k terms are added in
 * ascending order into an f32 that starts at 0, with the product fused into the add. */
void orc_matmul(const float* a, size_t as0, size_t as1, const float* b, size_t bs0, size_t bs1,
                size_t M, size_t K, size_t N, float* out) {
    for (size_t i = 0; i < M; i++)
        for (size_t j = 0; j < N; j++) {
            float acc = 0.0f;
            for (size_t k = 0; k < K; k++) acc = fmaf(a[i * as0 + k * as1], b[k * bs0 + j * bs1], acc);
            out[i * N + j] = acc;
        }
}
