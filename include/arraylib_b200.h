/*
 * arraylib_b200.h -- C ABI of libarraylib_b200.so, the B200 (sm_100a) array core.
 *
 * This is the drop-in boundary for the reference's `ffi.dlopen(".../arraylib.so")`
 * (arraylib/core.py:17-33). The struct layouts and every `array_*` symbol below are ABI
 * compatible with what the reference's cffi `cdef` binds (arraylib/src/arraylib.h:28-584), so
 * a binding written against the reference header resolves against this library. What differs:
 *
 *   - `Data.mem` / `NDArray.ptr` are DEVICE (HBM) pointers. The host must not dereference
 *     them; use array_to_host / array_read_base / array_fill for host <-> device traffic.
 *   - Nothing asserts/aborts. A failing call returns NULL (or a non-zero int) and
 *     al_last_error() returns "<PythonExceptionName>: message" for the calling thread.
 *   - Work is enqueued asynchronously on one library-owned CUDA stream; only the host-copy
 *     entry points and al_sync() block.
 *   - Host callbacks cannot run on the device: array_op / array_reduce (which take C function
 *     pointers in the reference) fail with NotImplementedError. Their replacements are
 *     array_op_named / array_reduce_named, which take an entry of the ufunc tables below.
 *   - Numerics: elementwise results -- eager or fused chain -- are bit-identical to the reference's x86-64 / glibc
 *     build, NaN sign and payload included; reduce_max / reduce_min are bit-identical,
 *     the sign of a zero result included; reduce_sum sums in a tree (stated tolerance); the FP64 tape extension
 *     returns the GPU's canonical NaN.
 *   - array_fill / array_to_host accept ordinary pageable host memory, like the reference; transfers of 4 MiB and
 *     more are staged through a pinned ring inside the library and do not hold the API lock while bytes move.
 *
 * Reference citations are relative to /root/reference.
 */
#ifndef ARRAYLIB_B200_H
#define ARRAYLIB_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef float f32;

/* ---- structs: layout-compatible with arraylib/src/arraylib.h:57-77 ------------------------ */
typedef struct {
    f32* mem;    /* DEVICE pointer to the base allocation */
    size_t size; /* elements in the base allocation */
    size_t refs; /* reference count (updated atomically) */
} Data;

typedef struct {
    size_t ndim;
    size_t* shape;  /* host array, ndim entries */
    size_t* stride; /* host array, ndim entries, in ELEMENTS, unsigned */
} Layout;

typedef struct {
    Data* data;
    f32* ptr; /* DEVICE pointer: mem + view offset */
    Layout* lay;
    bool view;
} NDArray;

/* ---- op tables ---------------------------------------------------------------------------- */
/* binary ops, in the order of arraylib/src/arraylib.c:12-25 */
enum al_binop {
    AL_SUM = 0, AL_SUB, AL_MUL, AL_DIV, AL_MOD, AL_POW, AL_MAX, AL_MIN,
    AL_EQ, AL_NE, AL_GT, AL_GE, AL_LT, AL_LE, AL_BINOP_COUNT
};
/* unary ufunc table, in the order of arraylib/src/arraylib.c:31-49 */
enum al_ufunc {
    AL_NEG = 0, AL_ABS, AL_SQRT, AL_EXP, AL_LOG, AL_SIN, AL_COS, AL_TAN, AL_ASIN, AL_ACOS,
    AL_ATAN, AL_SINH, AL_COSH, AL_TANH, AL_ASINH, AL_ACOSH, AL_ATANH, AL_CEIL, AL_FLOOR,
    AL_UFUNC_COUNT
};
/* reduce accumulators: sum/max/min are arraylib/src/arraylib.c:758-768; prod is an extension */
enum al_redop { AL_RSUM = 0, AL_RMAX, AL_RMIN, AL_RPROD, AL_REDOP_COUNT };

/* ---- runtime (new; the reference has no device, stream or error channel) ------------------ */
int al_init(int device);             /* bind the process to one GPU; idempotent. <0 = use ARRAYLIB_DEVICE / LOCAL_RANK / 0 */
int al_device(void);                 /* bound device ordinal, -1 before al_init */
int al_device_count(void);            /* CUDA devices visible to this process */
int al_sync(void);                   /* block until the library stream is idle */
void* al_stream(void);               /* the cudaStream_t all work is enqueued on */
const char* al_last_error(void);     /* thread-local; "" when the last call succeeded */
const char* al_version(void);
/* instrumentation for bench.py / tests (counts kernels THIS library launched) */
uint64_t al_launch_count(void);
const char* al_last_kernel(void);    /* name of the most recently launched kernel family */
int al_timer_start(void);            /* cudaEventRecord on the library stream */
int al_timer_stop(float* ms);        /* record + synchronize + elapsed milliseconds */
int al_event_record(int idx);        /* record pooled event #idx on the library stream (no sync) */
int al_event_elapsed(int first, int second, float* ms); /* waits for `second`, then the time between the two */
uint64_t al_bytes_in_use(void);      /* device bytes currently held by live Data blocks */
uint64_t al_bytes_cached(void);      /* idle bytes held by the block cache (bounded by the cache_limit_mb tuning knob, default half the device) */
int al_trim_pool(void);              /* return cached pool memory to the driver */
void* al_host_alloc(size_t bytes);   /* pinned host memory for staging */
int al_host_free(void* p);
int al_set_tuning(const char* key, long value); /* kernel selection overrides, see DESIGN.md */

/* ---- low-level helpers of the reference header (not used by its Python; present so its cdef binds) ---- */
typedef f32 (*binop)(f32, f32);
typedef f32 (*uniop)(f32);
void* alloc(size_t size);                                           /* arraylib.c:55-60 (host memory) */
Data* data_empty(size_t size);                                      /* arraylib.c:129-137; mem is device memory */
Layout* layout_alloc(size_t ndim);                                  /* arraylib.c:143-149 */
Layout* layout_copy(Layout* dst, const Layout* src);                /* arraylib.c:151-156 */
void layout_free(Layout* lay);                                      /* arraylib.c:158-162 */
Layout** layout_broadcast(const Layout** lays, size_t nlay);        /* arraylib.c:180-206 */
NDArray* array_scalar_op(binop fn, const NDArray* src, f32 rhs);    /* arraylib.c:491-501: always fails (NotImplementedError) */
NDArray* array_array_op(binop fn, const NDArray* lhs, const NDArray* rhs); /* arraylib.c:569-589: always fails (NotImplementedError) */

/* ---- storage (arraylib.h "ARRAY CREATION AND DESTRUCTION"; arraylib.c:212-231) ------------ */
NDArray* array_empty(const size_t* shape, size_t ndim);             /* arraylib.c:212-221 */
void array_free(NDArray* array);                                    /* arraylib.c:223-231 */
NDArray* array_shallow_copy(const NDArray* src);                    /* arraylib.c:237-246 */
NDArray* array_deep_copy(const NDArray* src);                       /* arraylib.c:248-256 */

/* ---- initialisers (arraylib.c:275-316) ---------------------------------------------------- */
NDArray* array_zeros(const size_t* shape, size_t ndim);             /* arraylib.c:275-280 */
NDArray* array_ones(const size_t* shape, size_t ndim);              /* arraylib.c:290-296 */
NDArray* array_fill(f32* elems, const size_t* shape, size_t ndim);  /* arraylib.c:282-288; elems is a HOST buffer (H2D upload) */
NDArray* array_arange(f32 start, f32 end, f32 step);                /* arraylib.c:298-308, incl. the f32 recurrence above 2^24 */
NDArray* array_linspace(f32 start, f32 end, f32 n);                 /* arraylib.c:310-316 */
NDArray* array_full(const size_t* shape, size_t ndim, f32 value);   /* extension: constant fill */

/* ---- host materialisation (replaces core.py:189,193,340-363 which read Data.mem directly) - */
int array_to_host(const NDArray* src, f32* dst);                    /* logical view -> contiguous host buffer of prod(shape) */
NDArray* array_fill_async(f32* elems, const size_t* shape, size_t ndim); /* extension: upload from PINNED memory on the H2D stream; returns at once; the compute stream waits for it when the array is first used */
int array_to_host_async(const NDArray* src, f32* dst);              /* extension: download to PINNED memory on the D2H stream; dst valid after al_sync() */
int array_read_base(const NDArray* src, size_t offset, size_t count, f32* dst); /* raw base buffer elements */
int array_write_base(NDArray* dst, size_t offset, size_t count, const f32* src);

/* ---- element / slice access (arraylib.c:322-415) ------------------------------------------ */
f32 array_get_elem(const NDArray* array, const size_t* index);      /* arraylib.c:322-324 (blocking D2H of one element; NaN + error on failure) */
NDArray* array_get_view(const NDArray* src, const size_t* start, const size_t* end, const size_t* step); /* arraylib.c:326-349 */
NDArray* array_set_elem_from_scalar(NDArray* array, f32 value, const size_t* index); /* arraylib.c:355-358 */
NDArray* array_set_view_from_scalar(NDArray* dst, f32 value, const size_t* start, const size_t* end, const size_t* step); /* arraylib.c:359-384 */
NDArray* array_set_view_from_array(NDArray* dst, const NDArray* src, const size_t* start, const size_t* end, const size_t* step); /* arraylib.c:386-415 */

/* ---- views (arraylib.c:258-269, 421-485): metadata only unless a contiguous copy is needed - */
NDArray* array_as_strided(const NDArray* src, const Layout* lay);   /* arraylib.c:258-269 */
NDArray* array_reshape(const NDArray* src, const size_t* shape, size_t ndim); /* arraylib.c:421-430 */
NDArray* array_transpose(const NDArray* src, const size_t* dims);   /* arraylib.c:432-442 */
NDArray* array_move_dim(const NDArray* src, const size_t* from, const size_t* to, size_t ndim); /* arraylib.c:444-475 */
NDArray* array_ravel(const NDArray* src);                           /* arraylib.c:477-485 */

/* ---- array o scalar (arraylib.c:491-516) -------------------------------------------------- */
NDArray* array_scalar_op_named(int binop, const NDArray* src, f32 rhs); /* arraylib.c:491-501 with the binop as a table entry */
NDArray* array_scalar_sum(const NDArray* src, f32 rhs);
NDArray* array_scalar_sub(const NDArray* src, f32 rhs);
NDArray* array_scalar_mul(const NDArray* src, f32 rhs);
NDArray* array_scalar_div(const NDArray* src, f32 rhs);
NDArray* array_scalar_mod(const NDArray* src, f32 rhs);
NDArray* array_scalar_pow(const NDArray* src, f32 rhs);
NDArray* array_scalar_max(const NDArray* src, f32 rhs);
NDArray* array_scalar_min(const NDArray* src, f32 rhs);
NDArray* array_scalar_eq(const NDArray* src, f32 rhs);
NDArray* array_scalar_ne(const NDArray* src, f32 rhs);
NDArray* array_scalar_gt(const NDArray* src, f32 rhs);
NDArray* array_scalar_ge(const NDArray* src, f32 rhs);
NDArray* array_scalar_lt(const NDArray* src, f32 rhs);
NDArray* array_scalar_le(const NDArray* src, f32 rhs);

/* ---- array o array with broadcasting (arraylib.c:569-632) --------------------------------- */
NDArray* array_array_op_named(int binop, const NDArray* lhs, const NDArray* rhs); /* arraylib.c:569-589 */
NDArray* array_array_sum(const NDArray* lhs, const NDArray* rhs);
NDArray* array_array_sub(const NDArray* lhs, const NDArray* rhs);
NDArray* array_array_mul(const NDArray* lhs, const NDArray* rhs);
NDArray* array_array_div(const NDArray* lhs, const NDArray* rhs);
NDArray* array_array_mod(const NDArray* lhs, const NDArray* rhs);
NDArray* array_array_pow(const NDArray* lhs, const NDArray* rhs);
NDArray* array_array_max(const NDArray* lhs, const NDArray* rhs);
NDArray* array_array_min(const NDArray* lhs, const NDArray* rhs);
NDArray* array_array_eq(const NDArray* lhs, const NDArray* rhs);
NDArray* array_array_ne(const NDArray* lhs, const NDArray* rhs);
NDArray* array_array_gt(const NDArray* lhs, const NDArray* rhs);
NDArray* array_array_ge(const NDArray* lhs, const NDArray* rhs);
NDArray* array_array_lt(const NDArray* lhs, const NDArray* rhs);
NDArray* array_array_le(const NDArray* lhs, const NDArray* rhs);
NDArray* array_array_matmul(const NDArray* lhs, const NDArray* rhs); /* arraylib.c:522-564 (compat) */
NDArray* array_array_dot(const NDArray* lhs, const NDArray* rhs);    /* arraylib.c:671-688 (compat) */

/* ---- fused elementwise chain (extension; replaces several eager array_array_op / array_scalar_op
 *      passes, arraylib.c:491-501,569-589, by one pass; results are bit-identical to the eager ones) */
typedef struct {
    int kind;     /* 0: acc = binop(acc, next array)  1: acc = binop(next array, acc)
                     2: acc = binop(acc, scalar)      3: acc = binop(scalar, acc)   4: acc = ufunc(acc) */
    int op;       /* al_binop (kinds 0-3) or al_ufunc (kind 4) */
    f32 scalar;   /* kinds 2 and 3 */
} al_chain_step;
/* acc starts as inputs[0]; array-consuming steps take inputs[1], inputs[2], ... in order.
 * At most 4 arrays and 8 steps; all arrays broadcast against each other (NumPy rule). */
NDArray* array_chain_eval(const NDArray* const* inputs, size_t n_inputs, const al_chain_step* steps, size_t n_steps);

/* ---- FP64 tape (extension; the device form of the reference's apply(fn) Python callback, core.py:455-460:
 *      every element is widened to double, a traced expression is evaluated in double precision and the
 *      result is rounded to f32 once). Instruction s defines value #s; operands name an earlier value
 *      (src >= 0), input array i (src = -1 - i) or the instruction's own constant (src = AL_TAPE_CONST). */
enum { AL_TAPE_CONST = -100 };
typedef struct {
    int kind;            /* 0: value = binop(a, b)   1: value = ufunc(a) */
    int op;              /* al_binop or al_ufunc */
    int a_src, b_src;
    double a_const, b_const;
} al_tape_insn;
NDArray* array_tape_eval_f64(const NDArray* const* inputs, size_t n_inputs, const al_tape_insn* tape, size_t n_insn, size_t result);

/* ---- unary ufunc table (arraylib.c:638-665) ----------------------------------------------- */
NDArray* array_op(uniop fn, const NDArray* src);       /* arraylib.c:638-645: always fails (NotImplementedError) */
NDArray* array_op_named(int ufunc, const NDArray* src);
NDArray* array_neg(const NDArray* src);
NDArray* array_abs(const NDArray* src);
NDArray* array_sqrt(const NDArray* src);
NDArray* array_exp(const NDArray* src);
NDArray* array_log(const NDArray* src);
NDArray* array_sin(const NDArray* src);
NDArray* array_cos(const NDArray* src);
NDArray* array_tan(const NDArray* src);
NDArray* array_asin(const NDArray* src);
NDArray* array_acos(const NDArray* src);
NDArray* array_atan(const NDArray* src);
NDArray* array_sinh(const NDArray* src);
NDArray* array_cosh(const NDArray* src);
NDArray* array_tanh(const NDArray* src);
NDArray* array_asinh(const NDArray* src);
NDArray* array_acosh(const NDArray* src);
NDArray* array_atanh(const NDArray* src);
NDArray* array_ceil(const NDArray* src);
NDArray* array_floor(const NDArray* src);

/* ---- reductions (arraylib.c:690-768): keepdims output, contiguous ------------------------- */
NDArray* array_reduce(binop acc_fn, const NDArray* src, const size_t* reduce_dims, size_t reduce_ndim, f32 acc_init); /* always fails (NotImplementedError) */
NDArray* array_reduce_named(int redop, const NDArray* src, const size_t* reduce_dims, size_t reduce_ndim, f32 acc_init);
NDArray* array_reduce_sum(const NDArray* src, const size_t* reduce_dims, size_t reduce_ndim);
NDArray* array_reduce_max(const NDArray* src, const size_t* reduce_dims, size_t reduce_ndim);
NDArray* array_reduce_min(const NDArray* src, const size_t* reduce_dims, size_t reduce_ndim);

/* ---- where (arraylib.c:774-838) ----------------------------------------------------------- */
NDArray* array_array_array_where(const NDArray* cond, const NDArray* lhs, const NDArray* rhs);
NDArray* array_array_scalar_where(const NDArray* cond, const NDArray* lhs, f32 rhs);
NDArray* array_scalar_array_where(const NDArray* cond, f32 lhs, const NDArray* rhs);
NDArray* array_scalar_scalar_where(const NDArray* cond, f32 lhs, f32 rhs);

/* ---- multi-GPU (new): one process per GPU; the only exchange step is the partial-result combine of a
 * reduction that crosses the sharded leading axis (the reference has no multi-device path,
 * its nearest relative is the per-thread accumulator merge, arraylib.c:732-750) ------------------------- */
int al_comm_unique_id(char out[128]);                       /* rank 0: ncclGetUniqueId */
int al_comm_init(const char id[128], int rank, int world);  /* ncclCommInitRank on the bound device */
int al_comm_allreduce(NDArray* a, int redop);               /* in place over a contiguous array (sum/max/min/prod) */
int al_comm_allgather(const NDArray* shard, NDArray* full); /* concatenate equal leading-dim shards */
int al_comm_alltoall(const NDArray* send, NDArray* recv);   /* block r of send -> rank r, block r of recv <- rank r (re-sharding) */
/* reduce_<op>(shard, dims) completed across ranks when dims contains axis 0; the full keepdims result on every
 * rank. One node: the local reduce kernel stores into CUDA-IPC peer windows over NVLink and the partials are
 * combined in rank order (deterministic); otherwise local reduce + ncclAllReduce. Collective call. */
NDArray* al_comm_reduce(int redop, const NDArray* shard, const size_t* reduce_dims, size_t reduce_ndim);
/* The same reduction with the result left SHARDED along its leading non-reduced axis (first axis whose keepdims
 * extent is > 1), rank r receiving block r: the peer-memory path then needs one hand-over and no result broadcast
 * (reduce-scatter). Needs that extent to be a multiple of the world size and 16-byte aligned blocks;
 * al_comm_can_scatter answers whether a (shape, dims) pair qualifies. Collective call. */
NDArray* al_comm_reduce_scatter(int redop, const NDArray* shard, const size_t* reduce_dims, size_t reduce_ndim);
/* both variants for a shard cut along `shard_axis` (every axis before it has extent 1) */
NDArray* al_comm_reduce_ex(int redop, const NDArray* shard, const size_t* reduce_dims, size_t reduce_ndim, size_t shard_axis, int scatter);
int al_comm_can_scatter(const size_t* shape, size_t ndim, const size_t* reduce_dims, size_t reduce_ndim);
const char* al_comm_transport(void);                        /* "peer-memory" | "nccl (<why>)" | "single process" */
int al_comm_destroy(void);

#ifdef __cplusplus
}
#endif
#endif /* ARRAYLIB_B200_H */
