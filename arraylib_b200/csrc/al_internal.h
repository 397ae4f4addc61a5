// Internal declarations shared by the translation units of libarraylib_b200.so.
// Nothing here crosses the C ABI (see include/arraylib_b200.h for that).
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "arraylib_b200.h"

namespace al {

constexpr int kMaxDims = 8;   // folded (post-merge) rank the kernels accept
constexpr int kNumSMs = 148;  // B200

// ---- per-process runtime -------------------------------------------------------------------
struct Runtime {
    int device = -1;
    cudaStream_t stream = nullptr;      // every kernel and every blocking copy
    cudaStream_t h2d_stream = nullptr;  // asynchronous uploads (array_fill_async)
    cudaStream_t d2h_stream = nullptr;  // asynchronous downloads (array_to_host_async)
    cudaEvent_t ev_copy = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    uint64_t launches = 0;
    const char* last_kernel = "";
    uint64_t bytes_in_use = 0;
    // tunables (al_set_tuning)
    long stream_store_min_bytes = 64l << 20;  // outputs at least this big use st.global.cs
    long load_policy = 0;      // streaming loads: 0 ld.global.cs, 1 ld.global.nc, 2 nc+L1::no_allocate
    long force_general = 0;    // 1: disable the vector / tile fast paths (tests the fallback)
    long reduce_splits = 0;    // 0 = auto
    long reduce_loads_per_lane = 0;  // reduce_inner team sizing: reduced items per lane (0 = default 4)
    long reduce_peel = 1;      // 0: misaligned unit-stride rows use scalar lanes instead of masked aligned float4s
    long reduce_sequential = 0;  // 1: one thread per output, the reference's summation order (bit-exact, slow)
    long skip_zero_sign_fix = 0; // 1: max/min keep the tree's zero sign (+0 / -0) instead of the reference's first-zero rule
    long disable_window = 0;
    long window_whole_blocks = 0;  // 1: fixed-w windows use the whole-block scan kernel instead of the half-block one
    long window_shfl = 1;      // fixed-w windows: 1..4 = register/shuffle kernel (cp.async ring shapes), 0 = shared-memory scan kernels
    long window_any_l16 = 1;   // generic-w window kernel: 16 elements per lane for w < 24 (0: always 32)
    long window_ctas_per_sm = 0;  // 0 = default (32 CTAs per SM, 4 of them resident)
    long tile64 = 1;           // 0: use the 32x32 scalar transpose tile even when the float4 one applies
    long tile_tq = 64;         // q extent of the 128-bit transpose tile: 32, 64 or 128 (p extent = 4096 / tq)
    long chain_block = 2;      // fused chains: 0 one vector per thread, 1 generic 4-row blocks, 2 also the specialised 8-row kernel
    long stream_tma = 0;       // 1: contiguous add/mul through the cp.async.bulk + mbarrier pipeline (experiment)
    long interleave = 1;       // 0: short-axis transposes always use the rectangular tile kernel
    long outer_tile = 1;       // 0: do not use the register-tiled outer-broadcast kernel
    long period_plan = 1;      // 0: short-inner-axis layouts stay on ew_run (no period tables)
    long heavy_stream = 1;     // persistent prefetching kernel for contiguous streams: 0 never, 1 pow, 2 pow and the FP64 ufuncs
    long heavy_ctas_per_sm = 0;  // persistent CTAs per SM of ew_heavy_stream_kernel (0 = as many as are resident)
    long host_staging = 1;     // 0: pageable from_numpy / to_numpy use plain cudaMemcpy instead of the library's pinned ring + copy threads
    long host_copy_nt = 1;     // 1: the staging ring's host copies use streaming stores (hostcopy.cpp), 0: memcpy
    long peer_collective = 1;  // 0: cross-GPU reductions go through ncclAllReduce instead of the peer-memory windows
};
Runtime& rt();
bool ensure_init();  // lazily al_init(-1); false (with error set) on failure

// ---- error channel ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void clear_error();
bool cuda_ok(cudaError_t e, const char* what);

#define AL_CUDA(expr)                                 \
    do {                                              \
        if (!::al::cuda_ok((expr), #expr)) return 0;  \
    } while (0)
#define AL_REQUIRE(cond, ...)             \
    do {                                  \
        if (!(cond)) {                    \
            ::al::set_error(__VA_ARGS__); \
            return 0;                     \
        }                                 \
    } while (0)

inline void note_launch(const char* name) {
    rt().launches++;
    rt().last_kernel = name;
}
bool check_launch(const char* name);  // cudaGetLastError after a launch

// ---- storage ---------------------------------------------------------------------------------
void* device_alloc(size_t bytes);  // cached cudaMalloc block; nullptr (error set) on failure
void device_free(void* p);         // back to the cache (safe immediately: single-stream ordering)
size_t count_of(const size_t* shape, size_t ndim);  // zero extents count as 1 (arraylib.c:66-72)
NDArray* new_array(const size_t* shape, size_t ndim);  // contiguous, fresh device buffer
NDArray* new_view(const NDArray* src, const size_t* shape, const size_t* stride, size_t ndim,
                  size_t extra_offset);
void release_array(NDArray* a);  // array_free without touching the thread's error slot (internal clean-up)
bool is_contiguous(const NDArray* a);
size_t view_offset(const NDArray* a);  // ptr - mem, in elements
bool valid(const NDArray* a, const char* who);
bool valid_mut(const NDArray* a, const char* who);  // valid() for an array about to be WRITTEN in place: also orders the write after pending async downloads of its buffer
void await_downloads(const Data* d);

// ---- folded layouts --------------------------------------------------------------------------
// Dimensions are stored INNERMOST FIRST (index 0 = fastest-varying logical axis).
struct Folded {
    int ndim = 0;
    uint64_t extent[kMaxDims];
    int64_t out_stride[kMaxDims];
    int64_t in_stride[4][kMaxDims];  // up to 4 operands (the fused chain kernel takes 4)
};
// Collapse an iteration space of `ndim` logical axes over `nin` strided operands and one
// strided output: extent-1 axes are dropped and neighbours that are mutually contiguous in every
// operand are merged. Returns false (error set) when more than kMaxDims axes survive.
bool fold_layout(size_t ndim, const size_t* shape, const size_t* out_stride, int nin,
                 const size_t* const* in_stride, Folded* f);

// ---- elementwise engine (elementwise.cu) -------------------------------------------------------
enum EwKind { EW_BINARY, EW_SCALAR, EW_UNARY, EW_COPY, EW_WHERE_AAA, EW_WHERE_AAS, EW_WHERE_ASA, EW_WHERE_ASS };
struct EwCall {
    EwKind kind;
    int op = 0;                    // al_binop / al_ufunc
    float s0 = 0.f, s1 = 0.f;      // scalar operands
    int nin = 0;
    const float* in[3] = {nullptr, nullptr, nullptr};
    float* out = nullptr;
};
// Runs `call` over a folded iteration space whose output is contiguous (out_stride ignored).
bool launch_elementwise(const EwCall& call, const Folded& f);
// Generic strided scatter: dst(strided) = src(strided) or scalar. Compat path (setters, cat).
bool launch_strided_assign(float* dst, const Folded& f, const float* src, float scalar);

// ---- reduce engine (reduce.cu) -----------------------------------------------------------------
NDArray* reduce_impl(int redop, const NDArray* src, const size_t* dims, size_t ndims, float init,
                     bool custom_init);

// ---- multi-GPU (comm.cu) -----------------------------------------------------------------------
bool comm_error_pending();  // true (error set) once a peer-memory collective has timed out

// ---- fills (runtime.cu) ------------------------------------------------------------------------
bool fill_const(float* p, size_t n, float v);

// Every extern "C" entry point serialises on one lock (ctypes drops the GIL around calls) and
// starts with a clean error slot.
std::recursive_mutex& api_mutex();
void bind_thread();  // makes the library's device current on the calling host thread (new threads start on device 0)
const std::string& last_error_string();
void restore_error(const std::string& s);

}  // namespace al

#define AL_ENTER                                                     \
    std::lock_guard<std::recursive_mutex> _guard(::al::api_mutex()); \
    ::al::bind_thread();                                             \
    ::al::clear_error()
