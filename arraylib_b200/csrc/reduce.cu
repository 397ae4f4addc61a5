// Reduce engine: keepdims reduction of a strided source over an arbitrary set of axes.
//
// Replaces array_reduce (arraylib/src/arraylib.c:690-756), which walks the SOURCE once in logical
// order and scatters into local_acc[dst_offset] with two offset_indexer calls per element. Here
// the axes are split on the host into a kept list and a reduced list (each folded, innermost
// first) and one of these sm_100a kernels is chosen per call:
//
//   reduce_outer   the innermost kept axis is unit-stride in the source ("output-contiguous"):
//                  threads run along the kept axis (coalesced, 128-bit when aligned), blockDim.y
//                  and gridDim.y split the reduced range, partials meet in shared memory.
//                  Config C4a: [64,(512*512),64] -> [64,1,1,64].
//   reduce_inner   the innermost reduced axis is unit-stride ("reduce-contiguous"): a team of
//                  2..32 lanes (warp shuffles) or a whole CTA per output. Config C4b: 2^24 x 64.
//   reduce_rows_packed  dense [N, 2..8] rows: four rows per thread as aligned 128-bit loads.
//   reduce_window  overlapping as_strided windows (kept stride == reduced stride == 1): persistent CTAs
//                  stage a tile + halo in shared memory and form all window sums from per-block
//                  prefix/suffix scans (half-block lanes for w = 8/16/32/64, config C5a; segmented
//                  scans over 32-element lane chunks for any other w up to 1024).
//   reduce_inner_peel  unit-stride rows that are not 16-byte aligned: masked aligned float4 loads.
//   reduce_sequential  verification mode: the reference's own summation order, one thread per output.
//   reduce_finish  second pass when the reduced range was split across CTAs (long reduced
//                  extent, few outputs): combines the per-split partials in a fixed order.
//
// Summation order differs from the reference (tree instead of sequential), so reduce_sum carries
// the tolerance stated in DESIGN.md; max/min are order independent.
#include <algorithm>
#include <type_traits>

#include "al_internal.h"
#include "device_utils.cuh"
#include "peer.cuh"

namespace al {


// accumulate step (arraylib.c:12-19). max/min use the hardware NaN-ignoring FMNMX: identical to the
// reference's fmax/fmin except for the SIGN of a zero result when the reduced set holds both +0 and -0
// (the reference returns whichever zero comes first in logical order; FMNMX gives +0 for max and -0 for
// min, independent of order). reduce_zero_sign_kernel repairs those outputs after the tree.
template <int OP>
__device__ __forceinline__ float racc(float a, float b) {
    if constexpr (OP == AL_RSUM) return __fadd_rn(a, b);
    else if constexpr (OP == AL_RMAX) return fmaxf(a, b);
    else if constexpr (OP == AL_RMIN) return fminf(a, b);
    else return __fmul_rn(a, b);
}
template <int OP>
__device__ __forceinline__ float rident() {
    if constexpr (OP == AL_RSUM) return 0.0f;
    else if constexpr (OP == AL_RMAX) return -INFINITY;
    else if constexpr (OP == AL_RMIN) return INFINITY;
    else return 1.0f;
}
// The reference seeds both dst and the thread-local accumulator with `init` and then merges them
// (arraylib.c:723-750): result = fn(init, fn(init, x0, x1, ...)).
template <int OP>
__device__ __forceinline__ float rfinal(float init, float v) {
    return racc<OP>(init, racc<OP>(init, v));
}

struct RedArgs {
    const float* src;
    float* dst;          // final output, or the partial buffer [split][nout]
    uint32_t nk_items;   // kept items (vector items in reduce_outer)
    uint32_t nr_items;   // reduced items (vector items in reduce_inner)
    uint32_t nout;       // output elements
    int nk, nr;
    uint32_t kext[kMaxDims];
    FastDiv kdiv[kMaxDims];
    long long kstride[kMaxDims];
    uint32_t rext[kMaxDims];
    FastDiv rdiv[kMaxDims];
    long long rstride[kMaxDims];
    uint32_t r_per_split;
    float init;
    int finalize;
    PeerMap peer;        // world > 0: final stores go to the owners' peer slots, un-finalised (peer.cuh)
    uint32_t rot;        // reduce_outer: CTA x-index rotation (rank * gridDim.x / world when storing to peers)
};

__device__ __forceinline__ long long kept_offset(const RedArgs& a, uint32_t g) {
    long long off = 0;
    for (int d = 0; d < a.nk; d++) {
        uint32_t idx;
        if (d == a.nk - 1) {
            idx = g;
        } else {
            uint32_t q = fdiv(g, a.kdiv[d]);
            idx = g - q * a.kext[d];
            g = q;
        }
        off += (long long)idx * a.kstride[d];
    }
    return off;
}
__device__ __forceinline__ long long red_offset(const RedArgs& a, uint32_t r) {
    if (a.nr == 1) return (long long)r * a.rstride[0];
    long long off = 0;
    for (int d = 0; d < a.nr; d++) {
        uint32_t idx;
        if (d == a.nr - 1) {
            idx = r;
        } else {
            uint32_t q = fdiv(r, a.rdiv[d]);
            idx = r - q * a.rext[d];
            r = q;
        }
        off += (long long)idx * a.rstride[d];
    }
    return off;
}

template <int VEC>
struct Acc;
template <>
struct Acc<1> {
    float v;
};
template <>
struct Acc<4> {
    float4 v;
};

template <int VEC>
__device__ __forceinline__ Acc<VEC> ld_acc(const float* p) {
    Acc<VEC> r;
    if constexpr (VEC == 1) r.v = __ldcs(p);
    else r.v = __ldcs(reinterpret_cast<const float4*>(p));
    return r;
}
template <int OP, int VEC>
__device__ __forceinline__ void acc_into(Acc<VEC>& a, const Acc<VEC>& b) {
    if constexpr (VEC == 1) {
        a.v = racc<OP>(a.v, b.v);
    } else {
        a.v.x = racc<OP>(a.v.x, b.v.x);
        a.v.y = racc<OP>(a.v.y, b.v.y);
        a.v.z = racc<OP>(a.v.z, b.v.z);
        a.v.w = racc<OP>(a.v.w, b.v.w);
    }
}
template <int OP, int VEC>
__device__ __forceinline__ Acc<VEC> acc_ident() {
    Acc<VEC> r;
    const float i = rident<OP>();
    if constexpr (VEC == 1) r.v = i;
    else r.v = make_float4(i, i, i, i);
    return r;
}

// ---- reduce_outer: threads along the kept axis --------------------------------------------------
// grid.x tiles the kept items, grid.y the reduced range; block = (TX, TY), TX*TY = 256.
template <int OP, int VEC>
__global__ void __launch_bounds__(256) reduce_outer_kernel(const RedArgs a) {
    extern __shared__ float4 smem_raw[];
    Acc<VEC>* sm = reinterpret_cast<Acc<VEC>*>(smem_raw);
    // With a peer epilogue every rank starts at its OWN output segment and walks round from there, so that at any
    // moment the ranks store to different owners (no NVLink incast on one GPU).
    uint32_t bx = blockIdx.x + a.rot;
    if (bx >= gridDim.x) bx -= gridDim.x;
    const uint32_t g = bx * blockDim.x + threadIdx.x;
    const bool live = g < a.nk_items;
    const uint32_t r0 = blockIdx.y * a.r_per_split;
    const uint32_t r1 = min(r0 + a.r_per_split, a.nr_items);
    Acc<VEC> acc = acc_ident<OP, VEC>();
    if (live) {
        const float* base = a.src + kept_offset(a, g) * (a.nk ? 1 : 0);
        const uint32_t step = blockDim.y;
        uint32_t r = r0 + threadIdx.y;
        // 4 independent loads in flight per thread
        for (; r + 3 * step < r1; r += 4 * step) {
            Acc<VEC> x0 = ld_acc<VEC>(base + red_offset(a, r));
            Acc<VEC> x1 = ld_acc<VEC>(base + red_offset(a, r + step));
            Acc<VEC> x2 = ld_acc<VEC>(base + red_offset(a, r + 2 * step));
            Acc<VEC> x3 = ld_acc<VEC>(base + red_offset(a, r + 3 * step));
            acc_into<OP, VEC>(x0, x1);
            acc_into<OP, VEC>(x2, x3);
            acc_into<OP, VEC>(x0, x2);
            acc_into<OP, VEC>(acc, x0);
        }
        for (; r < r1; r += step) {
            Acc<VEC> x = ld_acc<VEC>(base + red_offset(a, r));
            acc_into<OP, VEC>(acc, x);
        }
    }
    if (blockDim.y > 1) {
        sm[threadIdx.y * blockDim.x + threadIdx.x] = acc;
        __syncthreads();
        if (threadIdx.y != 0) return;
        for (uint32_t y = 1; y < blockDim.y; y++) acc_into<OP, VEC>(acc, sm[y * blockDim.x + threadIdx.x]);
    }
    if (!live) return;
    float* dst = a.peer.world ? peer_dst(a.peer, g * VEC) : a.dst + (size_t)blockIdx.y * a.nout + (size_t)g * VEC;
    if constexpr (VEC == 1) {
        dst[0] = a.finalize ? rfinal<OP>(a.init, acc.v) : acc.v;
    } else {
        float4 o = acc.v;
        if (a.finalize) {
            o.x = rfinal<OP>(a.init, o.x);
            o.y = rfinal<OP>(a.init, o.y);
            o.z = rfinal<OP>(a.init, o.z);
            o.w = rfinal<OP>(a.init, o.w);
        }
        *reinterpret_cast<float4*>(dst) = o;
    }
}

// ---- reduce_inner: a team of lanes along the reduced axis ----------------------------------------
template <int OP, int VEC>
__device__ __forceinline__ float acc_collapse(const Acc<VEC>& a) {
    if constexpr (VEC == 1) return a.v;
    else return racc<OP>(racc<OP>(a.v.x, a.v.y), racc<OP>(a.v.z, a.v.w));
}

// TEAM in {1,2,4,8,16,32}: sub-warp teams, one output per team (grid.y splits the reduced range).
template <int OP, int VEC, int TEAM>
__global__ void __launch_bounds__(256) reduce_inner_kernel(const RedArgs a) {
    const uint32_t gt = (uint32_t)(((uint64_t)blockIdx.x * 256u + threadIdx.x) / TEAM);  // team id == output id (64-bit: nout*TEAM may pass 2^32)
    const uint32_t lane = threadIdx.x % TEAM;
    const bool live = gt < a.nout;
    Acc<VEC> acc = acc_ident<OP, VEC>();
    if (live) {
        const float* base = a.src + (a.nk ? kept_offset(a, gt) : 0);
        // Lanes run along the innermost reduced axis; the outer reduced axes are walked by a plain loop
        // whose offset is resolved once per row (not once per load). This kernel is never split
        // (gridDim.y == 1), so the range is always the whole reduced space.
        const uint32_t e0 = a.rext[0];
        const uint32_t n_outer = a.nr_items / e0;
        const long long s0 = a.rstride[0];
        for (uint32_t ro = 0; ro < n_outer; ro++) {
            long long obase = 0;
            uint32_t rem = ro;
            for (int d = 1; d < a.nr; d++) {
                uint32_t idx;
                if (d == a.nr - 1) {
                    idx = rem;
                } else {
                    const uint32_t q = fdiv(rem, a.rdiv[d]);
                    idx = rem - q * a.rext[d];
                    rem = q;
                }
                obase += (long long)idx * a.rstride[d];
            }
            const float* row = base + obase;
            uint32_t r = lane;
            for (; r + 3 * TEAM < e0; r += 4 * TEAM) {
                Acc<VEC> x0 = ld_acc<VEC>(row + (long long)r * s0);
                Acc<VEC> x1 = ld_acc<VEC>(row + (long long)(r + TEAM) * s0);
                Acc<VEC> x2 = ld_acc<VEC>(row + (long long)(r + 2 * TEAM) * s0);
                Acc<VEC> x3 = ld_acc<VEC>(row + (long long)(r + 3 * TEAM) * s0);
                acc_into<OP, VEC>(x0, x1);
                acc_into<OP, VEC>(x2, x3);
                acc_into<OP, VEC>(x0, x2);
                acc_into<OP, VEC>(acc, x0);
            }
            for (; r < e0; r += TEAM) {
                Acc<VEC> x = ld_acc<VEC>(row + (long long)r * s0);
                acc_into<OP, VEC>(acc, x);
            }
        }
    }
    float v = acc_collapse<OP, VEC>(acc);
#pragma unroll
    for (int o = TEAM / 2; o > 0; o >>= 1) v = racc<OP>(v, __shfl_down_sync(0xffffffffu, v, o, TEAM));
    if (live && lane == 0) {
        float* dst = a.dst + (size_t)blockIdx.y * a.nout + gt;
        *dst = a.finalize ? rfinal<OP>(a.init, v) : v;
    }
}

// The same for unit-stride rows that are NOT 16-byte aligned (odd row lengths such as 255 or 85, sliced
// views): every row is read as the aligned float4s that cover it, and the up-to-3 foreign elements in the
// first and last vector are replaced by the identity. The alignment is resolved per row on the device,
// so there is no constraint on any stride. (The over-read stays inside the 16-byte chunk that holds a
// row element, hence inside the allocation.) Scalar lanes left 2^20 rows of 255 floats at 3.7 TB/s.
template <int OP>
__device__ __forceinline__ float4 peel_mask(float4 x, int first, int e0) {
    const float id = rident<OP>();
    if (first < 0 || first + 3 >= e0) {
        if (first < 0 || first >= e0) x.x = id;
        if (first + 1 < 0 || first + 1 >= e0) x.y = id;
        if (first + 2 < 0 || first + 2 >= e0) x.z = id;
        if (first + 3 < 0 || first + 3 >= e0) x.w = id;
    }
    return x;
}
template <int OP, int TEAM>
__global__ void __launch_bounds__(256) reduce_inner_peel_kernel(const RedArgs a) {
    const uint32_t gt = (uint32_t)(((uint64_t)blockIdx.x * 256u + threadIdx.x) / TEAM);
    const uint32_t lane = threadIdx.x % TEAM;
    const bool live = gt < a.nout;
    Acc<4> acc = acc_ident<OP, 4>();
    if (live) {
        const float* base = a.src + (a.nk ? kept_offset(a, gt) : 0);
        const int e0 = (int)a.rext[0];
        const uint32_t n_outer = a.nr_items / a.rext[0];
        for (uint32_t ro = 0; ro < n_outer; ro++) {
            long long obase = 0;
            uint32_t rem = ro;
            for (int d = 1; d < a.nr; d++) {
                uint32_t idx;
                if (d == a.nr - 1) {
                    idx = rem;
                } else {
                    const uint32_t q = fdiv(rem, a.rdiv[d]);
                    idx = rem - q * a.rext[d];
                    rem = q;
                }
                obase += (long long)idx * a.rstride[d];
            }
            const float* row = base + obase;
            const int m = (int)((reinterpret_cast<uintptr_t>(row) >> 2) & 3);
            const float4* q = reinterpret_cast<const float4*>(row - m);
            const uint32_t nvec = (uint32_t)(m + e0 + 3) >> 2;  // vector j holds row elements 4j-m .. 4j-m+3
            uint32_t j = lane;
            for (; j + 3 * TEAM < nvec; j += 4 * TEAM) {
                Acc<4> x0, x1, x2, x3;
                x0.v = __ldcs(q + j);
                x1.v = __ldcs(q + j + TEAM);
                x2.v = __ldcs(q + j + 2 * TEAM);
                x3.v = __ldcs(q + j + 3 * TEAM);
                x0.v = peel_mask<OP>(x0.v, 4 * (int)j - m, e0);
                x1.v = peel_mask<OP>(x1.v, 4 * (int)(j + TEAM) - m, e0);
                x2.v = peel_mask<OP>(x2.v, 4 * (int)(j + 2 * TEAM) - m, e0);
                x3.v = peel_mask<OP>(x3.v, 4 * (int)(j + 3 * TEAM) - m, e0);
                acc_into<OP, 4>(x0, x1);
                acc_into<OP, 4>(x2, x3);
                acc_into<OP, 4>(x0, x2);
                acc_into<OP, 4>(acc, x0);
            }
            for (; j < nvec; j += TEAM) {
                Acc<4> x;
                x.v = peel_mask<OP>(__ldcs(q + j), 4 * (int)j - m, e0);
                acc_into<OP, 4>(acc, x);
            }
        }
    }
    float v = acc_collapse<OP, 4>(acc);
#pragma unroll
    for (int o = TEAM / 2; o > 0; o >>= 1) v = racc<OP>(v, __shfl_down_sync(0xffffffffu, v, o, TEAM));
    if (live && lane == 0) a.dst[gt] = a.finalize ? rfinal<OP>(a.init, v) : v;
}

// Packed short rows: source is a dense [nout, R] block (R = 2..8) reduced over its inner axis -- the
// component axis of RGB / xyz / complex data. One lane per output would issue R scalar loads at a stride
// of R floats; instead a thread takes FOUR rows = 4R consecutive floats = R aligned float4 loads, sums
// each row out of registers (all indices are compile-time) and stores one float4.
template <int OP, int R>
__global__ void __launch_bounds__(256)
reduce_rows_packed_kernel(const float* __restrict__ src, float* __restrict__ out, uint32_t nout4, float init) {
    const uint32_t t = blockIdx.x * 256u + threadIdx.x;
    if (t >= nout4) return;
    const float4* p = reinterpret_cast<const float4*>(src) + (size_t)t * R;
    float x[4 * R];
#pragma unroll
    for (int j = 0; j < R; j++) {
        const float4 v = __ldg(p + j);
        x[4 * j] = v.x, x[4 * j + 1] = v.y, x[4 * j + 2] = v.z, x[4 * j + 3] = v.w;
    }
    float o[4];
#pragma unroll
    for (int row = 0; row < 4; row++) {
        float acc = x[row * R];
#pragma unroll
        for (int k = 1; k < R; k++) acc = racc<OP>(acc, x[row * R + k]);
        o[row] = rfinal<OP>(init, acc);
    }
    reinterpret_cast<float4*>(out)[t] = make_float4(o[0], o[1], o[2], o[3]);
}

template <int OP, int R>
static bool launch_rows_packed(const float* src, float* out, uint64_t nout, float init) {
    const uint32_t nout4 = (uint32_t)(nout / 4);
    reduce_rows_packed_kernel<OP, R><<<(nout4 + 255) / 256, 256, 0, rt().stream>>>(src, out, nout4, init);
    note_launch("reduce_rows_packed");
    return check_launch("reduce_rows_packed");
}

// One CTA per (output, split): long reduced extents.
template <int OP, int VEC>
__global__ void __launch_bounds__(256) reduce_inner_cta_kernel(const RedArgs a) {
    __shared__ float warp_part[8];
    const uint32_t o = blockIdx.x;
    const uint32_t r0 = blockIdx.y * a.r_per_split;
    const uint32_t r1 = min(r0 + a.r_per_split, a.nr_items);
    const float* base = a.src + (a.nk ? kept_offset(a, o) : 0);
    Acc<VEC> acc = acc_ident<OP, VEC>();
    uint32_t r = r0 + threadIdx.x;
    for (; r + 3 * 256 < r1; r += 4 * 256) {
        Acc<VEC> x0 = ld_acc<VEC>(base + red_offset(a, r));
        Acc<VEC> x1 = ld_acc<VEC>(base + red_offset(a, r + 256));
        Acc<VEC> x2 = ld_acc<VEC>(base + red_offset(a, r + 512));
        Acc<VEC> x3 = ld_acc<VEC>(base + red_offset(a, r + 768));
        acc_into<OP, VEC>(x0, x1);
        acc_into<OP, VEC>(x2, x3);
        acc_into<OP, VEC>(x0, x2);
        acc_into<OP, VEC>(acc, x0);
    }
    for (; r < r1; r += 256) {
        Acc<VEC> x = ld_acc<VEC>(base + red_offset(a, r));
        acc_into<OP, VEC>(acc, x);
    }
    float v = acc_collapse<OP, VEC>(acc);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v = racc<OP>(v, __shfl_down_sync(0xffffffffu, v, s));
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) v = racc<OP>(v, warp_part[w]);
        float* dst = a.dst + (size_t)blockIdx.y * a.nout + o;
        *dst = a.finalize ? rfinal<OP>(a.init, v) : v;
    }
}

// ---- reduce_sequential: the reference's own summation order (verification mode) --------------------------
// One thread per output walks its reduced elements in ascending logical order and accumulates exactly
// like the serial reference build (arraylib.c:737-750): acc = init; acc = fn(acc, x) ...; out = fn(init, acc).
// Slow by construction (no parallelism inside an output); selected with al_set_tuning("reduce_sequential", 1)
// to show that the ONLY difference between reduce_sum here and in the reference is the order of the adds.
template <int OP>
__global__ void __launch_bounds__(256) reduce_sequential_kernel(const RedArgs a) {
    const uint32_t o = blockIdx.x * 256u + threadIdx.x;
    if (o >= a.nout) return;
    const float* base = a.src + (a.nk ? kept_offset(a, o) : 0);
    float acc = a.init;
    for (uint32_t r = 0; r < a.nr_items; r++) {
        const float x = base[red_offset(a, r)];
        if constexpr (OP == AL_RMAX) acc = (acc >= x || x != x) ? acc : x;       // glibc fmax, left operand wins ties
        else if constexpr (OP == AL_RMIN) acc = (acc <= x || x != x) ? acc : x;  // glibc fmin
        else acc = racc<OP>(acc, x);
    }
    float r2;
    if constexpr (OP == AL_RMAX) r2 = (a.init >= acc || acc != acc) ? a.init : acc;
    else if constexpr (OP == AL_RMIN) r2 = (a.init <= acc || acc != acc) ? a.init : acc;
    else r2 = racc<OP>(a.init, acc);
    a.dst[o] = r2;
}

// ---- reduce_finish: out[o] = final(sum_s partial[s][o]) ------------------------------------------
template <int OP>
__global__ void __launch_bounds__(256)
reduce_finish_kernel(const float* __restrict__ partial, float* __restrict__ out, uint32_t nout, uint32_t splits,
                     float init, const PeerMap peer) {
    if (nout <= 2048) {  // one warp per output, lanes stride over the splits
        const uint32_t o = (blockIdx.x * 256u + threadIdx.x) >> 5, lane = threadIdx.x & 31;
        float v = rident<OP>();
        if (o < nout)
            for (uint32_t s = lane; s < splits; s += 32) v = racc<OP>(v, partial[(size_t)s * nout + o]);
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) v = racc<OP>(v, __shfl_down_sync(0xffffffffu, v, k));
        if (o < nout && lane == 0) {
            if (peer.world) *peer_dst(peer, o) = v;
            else out[o] = rfinal<OP>(init, v);
        }
    } else {
        const uint32_t o = blockIdx.x * 256u + threadIdx.x;
        if (o >= nout) return;
        float v = rident<OP>();
        for (uint32_t s = 0; s < splits; s++) v = racc<OP>(v, partial[(size_t)s * nout + o]);
        if (peer.world) *peer_dst(peer, o) = v;
        else out[o] = rfinal<OP>(init, v);
    }
}

// ---- reduce_window: overlapping windows, out[i] = sum_{j<w} x[i+j] --------------------------------
// Each CTA stages `span` consecutive source elements in shared memory and emits span - w outputs.
// The staged span is cut into blocks of w; with P = inclusive prefix scan and S = inclusive suffix
// scan inside each block, the window that starts at offset k of block b is
//     S[b][k]  (rest of its own block)  (+)  P[b+1][k-1]  (head of the next block, k > 0)
// i.e. 2 shared reads + 1 add per output instead of w reads (van Herk / Gil-Werman), and the
// rounding behaviour is that of two sequential sums of <= w terms. Blocks are stored with a pitch
// of w+1 floats so that threads scanning different blocks hit different banks.

// Fixed-w variant (w = 8/16/32/64): the first half of the CTA's warps build the suffix scans, the
// second half the prefix scans, each thread owning 64/W blocks held in registers.
constexpr int kWinFixedThreads = 128; // fixed-w kernel: 128 threads x 32 elements = 4096-element span,
constexpr int kWinLoads = 32;         // all 32 loads of a thread in flight at once; ~5 CTAs per SM

// With the default init (the op's identity) the reference's double init fold collapses: for sum it
// only turns a -0 result into +0 (0 + (0 + v)), for max/min/prod it is the identity.
template <int OP>
__device__ __forceinline__ float rfinal_fast(float init, float v, bool init_is_identity) {
    if (!init_is_identity) return rfinal<OP>(init, v);
    if constexpr (OP == AL_RSUM) return __fadd_rn(v, 0.0f);
    else return racc<OP>(init, v);  // one fold is enough (idempotent); it maps an all-NaN max/min to the identity
}

template <int OP, int W>
__global__ void __launch_bounds__(kWinFixedThreads, 4)
reduce_window_fixed_kernel(const float* __restrict__ x, float* __restrict__ out, uint32_t nout, uint32_t n_tiles,
                           uint32_t tiles_per_row, long long row_stride, float init, int init_is_identity) {
    // `nout` windows per row; rows (a batch of independent signals) are `row_stride` source elements
    // apart and nout outputs apart; tiles are numbered row-major over (row, tile within row).
    extern __shared__ float wsm[];
    constexpr int T = kWinFixedThreads, SPAN = kWinFixedThreads * kWinLoads;
    constexpr int PITCH = W + 1, NBLK = SPAN / W, TILE = SPAN - W, PER_THREAD = 64 / W;
    constexpr int SHIFT = W == 64 ? 6 : W == 32 ? 5 : W == 16 ? 4 : 3;
    constexpr int BLK_PER_ROUND = T / W;  // blocks covered by one sweep of the CTA over consecutive elements
    float* P = wsm;
    float* S = wsm + NBLK * PITCH;
    const size_t n_in = (size_t)nout + W - 1;
    const int tid = threadIdx.x;
    // Element e = u*T + tid of a tile lives at P[(e / W) * PITCH + e % W]; since W divides T, e % W is
    // a per-thread constant and e / W advances by T/W per round, so every shared address below is
    // `lane base + compile-time offset`.
    const int kk = tid & (W - 1);
    float* const stage = P + (tid >> SHIFT) * PITCH + kk;
    const float* const s_lane = S + (tid >> SHIFT) * PITCH + kk;
    const float* const p_lane = P + ((tid >> SHIFT) + 1) * PITCH + kk - 1;
    const bool prefix_role = tid >= T / 2;  // warp-uniform
    const int owner = tid & (T / 2 - 1);
    const bool has_head = kk != 0;
    const bool ident = init_is_identity != 0;

    // Persistent CTA: the loads of tile k+1 are issued right after tile k has been parked in shared
    // memory, so they are in flight during the whole scan/emit work of tile k.
    float v[kWinLoads];
    auto fetch = [&](uint32_t tile) {
        const uint32_t row = tile / tiles_per_row;
        const size_t i0 = (size_t)(tile - row * tiles_per_row) * TILE;
        const float* src = x + (long long)row * row_stride + i0 + tid;
        if (i0 + SPAN <= n_in) {  // CTA-uniform fast path: no per-element bounds checks
#pragma unroll
            for (int u = 0; u < kWinLoads; u++) v[u] = __ldcs(src + u * T);
        } else {
#pragma unroll
            for (int u = 0; u < kWinLoads; u++) v[u] = (i0 + u * T + tid < n_in) ? __ldcs(src + u * T) : rident<OP>();
        }
    };
    uint32_t tile = blockIdx.x;
    if (tile < n_tiles) fetch(tile);
    for (; tile < n_tiles; tile += gridDim.x) {
#pragma unroll
        for (int u = 0; u < kWinLoads; u++) stage[u * BLK_PER_ROUND * PITCH] = v[u];
        __syncthreads();
        if (tile + gridDim.x < n_tiles) fetch(tile + gridDim.x);
        float xr[PER_THREAD][W];
#pragma unroll
        for (int m = 0; m < PER_THREAD; m++) {
            const float* p = P + (owner * PER_THREAD + m) * PITCH;
#pragma unroll
            for (int k = 0; k < W; k++) xr[m][k] = p[k];
        }
        __syncthreads();  // every block is in registers before P is overwritten in place
#pragma unroll
        for (int m = 0; m < PER_THREAD; m++) {
            const int b = owner * PER_THREAD + m;
            if (prefix_role) {
                float* p = P + b * PITCH;
                float run = xr[m][0];
#pragma unroll
                for (int k = 1; k < W; k++) {
                    run = racc<OP>(run, xr[m][k]);
                    p[k] = run;
                }
            } else {
                float* sfx = S + b * PITCH;
                float run = xr[m][W - 1];
                sfx[W - 1] = run;
#pragma unroll
                for (int k = W - 2; k >= 0; k--) {
                    run = racc<OP>(xr[m][k], run);
                    sfx[k] = run;
                }
            }
        }
        __syncthreads();
        // output e = u*T + tid: S[b][kk] (+) P[b+1][kk-1], b = u*T/W + tid/W
        const uint32_t row = tile / tiles_per_row;
        const size_t i0 = (size_t)(tile - row * tiles_per_row) * TILE;
        float* dst = out + (size_t)row * nout + i0 + tid;
        if (i0 + TILE <= nout) {
#pragma unroll
            for (int u = 0; u < kWinLoads; u++) {
                if (u * T + tid < TILE) {  // only the last round is partial (TILE = SPAN - W)
                    float r = s_lane[u * BLK_PER_ROUND * PITCH];
                    if (has_head) r = racc<OP>(r, p_lane[u * BLK_PER_ROUND * PITCH]);
                    __stcs(dst + u * T, rfinal_fast<OP>(init, r, ident));
                }
            }
        } else {
#pragma unroll
            for (int u = 0; u < kWinLoads; u++) {
                if (u * T + tid < TILE && i0 + u * T + tid < nout) {
                    float r = s_lane[u * BLK_PER_ROUND * PITCH];
                    if (has_head) r = racc<OP>(r, p_lane[u * BLK_PER_ROUND * PITCH]);
                    __stcs(dst + u * T, rfinal_fast<OP>(init, r, ident));
                }
            }
        }
        __syncthreads();  // P and S are free again before the next tile is parked
    }
}

// Half-block variant (default for w = 8/16/32/64; config C5a uses w = 64): like the kernel above, but a
// block is split between two adjacent lanes. Each lane reads only its w/2 elements (one shared read per
// element instead of two), builds the prefix and suffix scans of its half in registers, swaps half
// totals with its partner through one shuffle and folds the partner's total into the scan that crosses
// the middle of the block. That is 8 LSU operations per output instead of 9, on the pipe that bounds
// this kernel. Blocks are laid out with a pitch of w+2 floats and a one-float gap between the halves;
// consecutive lanes own consecutive half blocks, so a warp's 32 half blocks fall on 32 different banks
// for every w (the whole-block kernel strides lanes by 64/w blocks and conflicts for w < 64).
template <int OP, int W>
__global__ void __launch_bounds__(kWinFixedThreads, 4)
reduce_window_half_kernel(const float* __restrict__ x, float* __restrict__ out, uint32_t nout, uint32_t n_tiles,
                          uint32_t tiles_per_row, long long row_stride, float init, int init_is_identity) {
    extern __shared__ float wsm[];
    constexpr int H = W / 2, T = kWinFixedThreads, SPAN = kWinFixedThreads * kWinLoads;
    constexpr int PITCH = W + 2, NBLK = SPAN / W, TILE = SPAN - W;
    constexpr int UPT = 2 * NBLK / T;  // half blocks per lane: UPT * H == kWinLoads
    constexpr int SHIFT = W == 64 ? 6 : W == 32 ? 5 : W == 16 ? 4 : 3;
    constexpr int BLK_PER_ROUND = T / W;
    static_assert(UPT * H == kWinLoads && UPT >= 1, "a lane holds kWinLoads elements");
    float* P = wsm;
    float* S = wsm + NBLK * PITCH;
    const size_t n_in = (size_t)nout + W - 1;
    const int tid = threadIdx.x;
    // element e = u*T + tid of a tile: block u*T/W + tid/W, offset kk = tid%W, stored at kk + (kk >= H)
    const int kk = tid & (W - 1);
    const int slot = (tid >> SHIFT) * PITCH + kk + (kk >= H ? 1 : 0);
    float* const stage = P + slot;
    const float* const s_lane = S + slot;
    // Prefix values are stored one logical slot to the right (P[b][k] sits where element k+1 would), so
    // that the read of P[b+1][kk-1] uses the same lane->bank map as S[b][kk] (reading it in place would
    // straddle the gap between the halves and collide on one bank in every second warp).
    const float* const p_lane = P + slot + PITCH;  // P[b+1][kk-1], kk >= 1
    const int half = tid & 1;
    const int scan_base = (tid >> 1) * PITCH + half * (H + 1);  // unit m adds m * (T/2) blocks
    const bool has_head = kk != 0;
    const bool ident = init_is_identity != 0;

    float v[kWinLoads];
    auto fetch = [&](uint32_t tile) {
        const uint32_t row = tile / tiles_per_row;
        const size_t i0 = (size_t)(tile - row * tiles_per_row) * TILE;
        const float* src = x + (long long)row * row_stride + i0 + tid;
        if (i0 + SPAN <= n_in) {
#pragma unroll
            for (int u = 0; u < kWinLoads; u++) v[u] = __ldcs(src + u * T);
        } else {
#pragma unroll
            for (int u = 0; u < kWinLoads; u++) v[u] = (i0 + u * T + tid < n_in) ? __ldcs(src + u * T) : rident<OP>();
        }
    };
    uint32_t tile = blockIdx.x;
    if (tile < n_tiles) fetch(tile);
    for (; tile < n_tiles; tile += gridDim.x) {
#pragma unroll
        for (int u = 0; u < kWinLoads; u++) stage[u * BLK_PER_ROUND * PITCH] = v[u];
        __syncthreads();
        if (tile + gridDim.x < n_tiles) fetch(tile + gridDim.x);
        float xr[UPT][H];
#pragma unroll
        for (int m = 0; m < UPT; m++)
#pragma unroll
            for (int k = 0; k < H; k++) xr[m][k] = P[scan_base + m * (T / 2) * PITCH + k];
        __syncthreads();  // every half block is in registers before P is overwritten in place
#pragma unroll
        for (int m = 0; m < UPT; m++) {
            const int base = scan_base + m * (T / 2) * PITCH;
            float lp[H];
            lp[0] = xr[m][0];
#pragma unroll
            for (int k = 1; k < H; k++) lp[k] = racc<OP>(lp[k - 1], xr[m][k]);
            const float other = __shfl_xor_sync(0xffffffffu, lp[H - 1], 1);  // partner's half total
            const float before = half ? other : rident<OP>();  // what precedes this half inside the block
            const float after = half ? rident<OP>() : other;   // what follows it
            float* const pw = P + (base - half * (H + 1)) + (half ? H + 2 : 1);  // shifted home of this half's prefixes
#pragma unroll
            for (int k = 0; k < H; k++) pw[k + ((!half && k == H - 1) ? 1 : 0)] = racc<OP>(before, lp[k]);
            float run = xr[m][H - 1];
            S[base + H - 1] = racc<OP>(run, after);
#pragma unroll
            for (int k = H - 2; k >= 0; k--) {
                run = racc<OP>(xr[m][k], run);
                S[base + k] = racc<OP>(run, after);
            }
        }
        __syncthreads();
        const uint32_t row = tile / tiles_per_row;
        const size_t i0 = (size_t)(tile - row * tiles_per_row) * TILE;
        float* dst = out + (size_t)row * nout + i0 + tid;
        if (i0 + TILE <= nout) {
#pragma unroll
            for (int u = 0; u < kWinLoads; u++) {
                if (u * T + tid < TILE) {
                    float r = s_lane[u * BLK_PER_ROUND * PITCH];
                    if (has_head) r = racc<OP>(r, p_lane[u * BLK_PER_ROUND * PITCH]);
                    __stcs(dst + u * T, rfinal_fast<OP>(init, r, ident));
                }
            }
        } else {
#pragma unroll
            for (int u = 0; u < kWinLoads; u++) {
                if (u * T + tid < TILE && i0 + u * T + tid < nout) {
                    float r = s_lane[u * BLK_PER_ROUND * PITCH];
                    if (has_head) r = racc<OP>(r, p_lane[u * BLK_PER_ROUND * PITCH]);
                    __stcs(dst + u * T, rfinal_fast<OP>(init, r, ident));
                }
            }
        }
        __syncthreads();
    }
}

// Any-w variant (w not in {8,16,32,64}): the same prefix/suffix decomposition, but the block length is a
// run-time value, so blocks cannot be dealt out to lanes. Instead every lane owns 32 CONSECUTIVE staged
// elements and the two scans become SEGMENTED scans over the tile (segments = blocks of w):
//   1. a lane folds its chunk into two aggregates (the open segment at its right end / at its left end),
//   2. a segmented scan of the aggregates across the CTA (warp shuffles + one shared hop) gives every
//      lane the value carried into its chunk from the left and from the right,
//   3. the lane re-walks its chunk seeded with the carries and stores S (inclusive) and P (EXCLUSIVE:
//      what precedes the element inside its block, the identity at a block start),
//   4. outputs: S[i] (+) P[i + w] -- no special case for windows that coincide with a block, and both
//      shared addresses are a per-lane base plus a compile-time offset.
// Elements are stored with one pad float per 32 so that lane t's chunk starts at bank t. Tiles start on
// a block boundary, so the positions of the block boundaries inside a lane's chunk depend on (lane, w)
// only and are kept as two 32-bit masks. The old kernel (one lane scans one whole block serially) ran a
// 100-wide window at 1.8 TB/s because only nblk of its 256 lanes had work during the scans.
// L = 32 elements per lane needs 128 registers (4 CTAs of 128 threads per SM: ncu showed 24 % occupancy, the shared-memory
// round trips exposed as `short_scoreboard` / `wait` stalls, issue slots 64 % busy). The kernel is also instantiated with
// L = 16 (tile of 2048 elements): half the registers, twice the warps. Measured: +15 % at w = 12, nothing from w = 24 on
// (4.1-4.3 TB/s either way: the shared-memory traffic, not the occupancy, is what binds), so only w < 24 uses it.
constexpr int kAnyT = 128;

template <int OP, int L>
__global__ void __launch_bounds__(kAnyT, L == 32 ? 4 : 7)
reduce_window_any_kernel(const float* __restrict__ x, float* __restrict__ out, uint32_t nout, uint32_t w, uint32_t n_tiles,
                         uint32_t tiles_per_row, long long row_stride, float init, int init_is_identity) {
    extern __shared__ float wsm[];
    constexpr int T = kAnyT, SPAN = kAnyT * L, NW = kAnyT / 32, PADDED = SPAN + SPAN / 32;
    float* P = wsm;               // staged elements, overwritten in place by the prefix scans
    float* S = wsm + PADDED;      // suffix scans
    __shared__ float inc[2][T];   // inclusive cross-lane scan values: [0] left-to-right, [1] right-to-left
    __shared__ float wtot[2][NW];
    __shared__ int wflag[2][NW];
    const uint32_t TILE = SPAN - w;  // outputs per tile
    const size_t n_in = (size_t)nout + w - 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool ident = init_is_identity != 0;
    const float id = rident<OP>();
    uint32_t starts = 0, ends = 0;  // bit k: chunk element k opens / closes a block
    {
        uint32_t pos = (uint32_t)(tid * L) % w;
        for (int k = 0; k < L; k++) {
            if (pos == 0) starts |= 1u << k;
            if (pos == w - 1) ends |= 1u << k;
            if (++pos == w) pos = 0;
        }
    }
    // output e = u*T + tid reads S at e + e/32 = u*(T + NW) + tid + warp and P at the padded index of e + w
    const float* const s_lane = S + tid + warp;
    const float* const p_lane = P + (tid + w) + ((tid + w) >> 5);

    float v[L];
    auto fetch = [&](uint32_t tile) {
        const uint32_t row = tile / tiles_per_row;
        const size_t i0 = (size_t)(tile - row * tiles_per_row) * TILE;
        const float* src = x + (long long)row * row_stride + i0 + tid;
        if (i0 + SPAN <= n_in) {
#pragma unroll
            for (int u = 0; u < L; u++) v[u] = __ldcs(src + u * T);
        } else {
#pragma unroll
            for (int u = 0; u < L; u++) v[u] = (i0 + u * T + tid < n_in) ? __ldcs(src + u * T) : id;
        }
    };
    uint32_t tile = blockIdx.x;
    if (tile < n_tiles) fetch(tile);
    for (; tile < n_tiles; tile += gridDim.x) {
        // element e = u*T + tid lives at e + e/32 = u*(T + NW) + tid + warp
#pragma unroll
        for (int u = 0; u < L; u++) P[u * (T + NW) + tid + warp] = v[u];
        __syncthreads();
        if (tile + gridDim.x < n_tiles) fetch(tile + gridDim.x);
        float xr[L];
        float* const mine = P + tid * L + (tid * L) / 32;  // one pad float per 32 elements: a chunk never straddles a pad
#pragma unroll
        for (int k = 0; k < L; k++) xr[k] = mine[k];
        // 1. aggregates: al = fold of the elements after the last block start (all of them if none),
        //                ar = fold of the elements before the first block end (all of them if none)
        float al = xr[0], ar = xr[L - 1];
#pragma unroll
        for (int k = 1; k < L; k++) {
            al = ((starts >> k) & 1u) ? xr[k] : racc<OP>(al, xr[k]);
            ar = ((ends >> (L - 1 - k)) & 1u) ? xr[L - 1 - k] : racc<OP>(xr[L - 1 - k], ar);
        }
        // 2. segmented inclusive scans of (aggregate, "chunk holds a boundary") across the CTA
        float sl = al, sr = ar;
        int fl = starts != 0, fr = ends != 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const float ol = __shfl_up_sync(0xffffffffu, sl, d), orr = __shfl_down_sync(0xffffffffu, sr, d);
            const int ofl = __shfl_up_sync(0xffffffffu, fl, d), ofr = __shfl_down_sync(0xffffffffu, fr, d);
            if (lane >= d) {
                if (!fl) sl = racc<OP>(ol, sl);
                fl |= ofl;
            }
            if (lane + d < 32) {
                if (!fr) sr = racc<OP>(sr, orr);
                fr |= ofr;
            }
        }
        if (lane == 31) wtot[0][warp] = sl, wflag[0][warp] = fl;
        if (lane == 0) wtot[1][warp] = sr, wflag[1][warp] = fr;
        __syncthreads();
        if (!fl && warp > 0) {  // still open towards the left: take what the warps before this one carry
            float c = wtot[0][0];
            for (int j = 1; j < warp; j++) c = wflag[0][j] ? wtot[0][j] : racc<OP>(c, wtot[0][j]);
            sl = racc<OP>(c, sl);
        }
        if (!fr && warp < NW - 1) {
            float c = wtot[1][NW - 1];
            for (int j = NW - 2; j > warp; j--) c = wflag[1][j] ? wtot[1][j] : racc<OP>(wtot[1][j], c);
            sr = racc<OP>(sr, c);
        }
        inc[0][tid] = sl;
        inc[1][tid] = sr;
        __syncthreads();
        // 3. the scans proper, seeded with the carries (lane 0 starts a block; the carry into the last
        //    lane from beyond the tile does not exist, and the S values that would need it are never read)
        float runp = tid > 0 ? inc[0][tid - 1] : id;
        float runs = tid < T - 1 ? inc[1][tid + 1] : id;
        float* const sfx = S + tid * L + (tid * L) / 32;
#pragma unroll
        for (int k = 0; k < L; k++) {
            const float before = ((starts >> k) & 1u) ? id : runp;  // exclusive prefix of element k
            mine[k] = before;
            runp = racc<OP>(before, xr[k]);
            const int kr = L - 1 - k;
            const float after = ((ends >> kr) & 1u) ? id : runs;
            runs = racc<OP>(xr[kr], after);
            sfx[kr] = runs;
        }
        __syncthreads();
        // 4. output e = u*T + tid: S[e] (+) Pex[e + w]
        const uint32_t row = tile / tiles_per_row;
        const size_t i0 = (size_t)(tile - row * tiles_per_row) * TILE;
        float* dst = out + (size_t)row * nout + i0 + tid;
        const uint32_t live = (uint32_t)min((size_t)TILE, (size_t)nout - i0);  // outputs of this tile that exist
#pragma unroll
        for (int u = 0; u < L; u++) {
            if (u * T + tid < live) {
                const float r = racc<OP>(s_lane[u * (T + NW)], p_lane[u * (T + NW)]);
                __stcs(dst + u * T, rfinal_fast<OP>(init, r, ident));
            }
        }
        __syncthreads();
    }
}

// Register/shuffle variant (default for w = 8/16/32/64/128 when source and output rows are 16-byte aligned;
// config C5a). The kernels above push every element through shared memory six times (stage, read, P and
// S write, two output reads): ~8 LSU operations per output on a pipe that saturates at SM-clock rate, so
// under a power cap (lower SM clock) they fall below the HBM roofline (ncu: L1/TEX 83 %, DRAM 75 %). Here
// nothing touches shared memory and there is no block barrier:
//   * a warp owns R consecutive "rows" of 128 source elements; lane l holds elements 4l..4l+3 of a row as
//     ONE 128-bit load, so a block of w elements is G = w/4 adjacent lanes;
//   * lane-local prefix/suffix sums (6 adds), then the block-wide EXCLUSIVE prefix E and suffix F of the
//     lane totals by Kogge-Stone over the G lanes: log2(G) shfl.up + log2(G) shfl.down with predicated adds; four
//     rows are scanned together so that their (independent) shuffle chains hide each other's latency;
//   * out[k] = S[b][k] (+) Pex[b+1][k]: the exclusive prefix of the NEXT block lives G lanes to the right
//     (in the next row for the last block of a row), fetched with 4 index shuffles per row; the publisher
//     selects which row it publishes, so one shuffle serves both cases;
//   * results leave as one 128-bit store per lane.
// Per 128 outputs: 1 LDG.128 + 1 STG.128 + 2*log2(G)+4 SHFL (12 for w = 64) and ~40 FADD -- about a third of
// the LSU work of the shared-memory kernels. Loads are double buffered in registers (B rows in flight while
// B rows are being scanned). Each segment of R = B*NB-1 output rows re-scans one halo row (only its first G
// lanes load anything).

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gmem_src, uint32_t src_bytes) {
    // 16 bytes global -> shared without passing through registers; bytes beyond src_bytes are zero-filled and not read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// B rows per batch, S batches per warp in a shared-memory ring (S - 1 of them in flight). The ring exists only to keep
// bytes in flight WITHOUT holding registers: a register double buffer gave 64 KB per SM in flight and the warps spent
// 39 % of their time waiting for it (ncu, round 2: DRAM 73 % busy, issue slots 50 %). Every lane reads back exactly
// the 16-byte slots it copied itself (cp.async + wait_group is per thread), so there is no barrier, no mbarrier and no
// cross-proxy ordering anywhere.
template <int OP, int W, int B, int S, int NB, int WPC>
__global__ void __launch_bounds__(WPC * 32, 512 / (WPC * 32))
reduce_window_shfl_kernel(const float* __restrict__ x, float* __restrict__ out, uint32_t nout, uint32_t n_segs,
                          uint32_t segs_per_row, long long row_stride, float init, int init_is_identity) {
    extern __shared__ float4 ring_raw[];
    constexpr int G = W / 4;         // lanes per block of W elements
    constexpr int R = B * NB - 1;    // output rows (of 128) per segment of NB batches; the last loaded row is the halo
    constexpr int IL = 4;            // rows scanned together: their shuffle chains are independent and interleave
    constexpr int STEPS = G == 2 ? 1 : G == 4 ? 2 : G == 8 ? 3 : G == 16 ? 4 : 5;
    static_assert(G >= 2 && G <= 32 && (G & (G - 1)) == 0 && B % IL == 0, "w/4 must be a power of two up to 32");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t gw = blockIdx.x * WPC + warp, nw = gridDim.x * WPC;
    const uint32_t n_in = nout + W - 1;  // < 2^31 + 128: element indices inside a signal fit 32 bits
    const bool ident = init_is_identity != 0;
    const float id = rident<OP>();
    const int partner = (lane + G) & 31;
    const bool low = lane < G;  // this lane's consumer sits in the PREVIOUS row
    bool pu[STEPS], pd[STEPS];  // does the Kogge-Stone source lane of step k lie inside this lane's block?
#pragma unroll
    for (int k = 0; k < STEPS; k++) {
        pu[k] = (lane & (G - 1)) >= (1 << k);
        pd[k] = (lane & (G - 1)) + (1 << k) < G;
    }
    if (gw >= n_segs) return;
    const uint32_t n_it = ((n_segs - gw + nw - 1) / nw) * NB;  // batches this warp walks through

    float4* const ring = ring_raw + (size_t)warp * (S * B * 32) + lane;            // [stage][row][lane]
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);

    // first element (inside its signal) of batch `it`, and the signal it belongs to
    auto locate = [&](uint32_t it, uint32_t& row, uint32_t& r0) {
        const uint32_t seg = gw + (it / NB) * nw;
        row = segs_per_row == n_segs ? 0u : seg / segs_per_row;  // one signal (C5a): no division
        r0 = (seg - row * segs_per_row) * R + (it % NB) * B;
    };
    auto issue = [&](uint32_t it) {  // batch `it` -> ring stage it % S
        if (it < n_it) {
            uint32_t row, r0;
            locate(it, row, r0);
            const uint32_t e0 = r0 * 128 + 4 * lane;  // this lane's first element
            const float* src = x + (long long)row * row_stride + e0;
            const uint32_t dst = ring_s + (it % S) * (B * 32 * 16);
            const bool halo_idle = it % NB == NB - 1 && !low;  // last row of the segment: only the first block is needed
            if ((r0 + B) * 128 <= n_in) {  // warp-uniform: the whole batch is in range
#pragma unroll
                for (int u = 0; u < B - 1; u++) cp_async16(dst + u * (32 * 16), src + u * 128, 16u);
                cp_async16(dst + (B - 1) * (32 * 16), halo_idle ? x : src + (B - 1) * 128, halo_idle ? 0u : 16u);
            } else {
#pragma unroll
                for (int u = 0; u < B; u++) {
                    const uint32_t e = e0 + u * 128;
                    uint32_t bytes = e + 4 <= n_in ? 16u : e < n_in ? (n_in - e) * 4u : 0u;
                    if (u == B - 1 && halo_idle) bytes = 0;
                    cp_async16(dst + u * (32 * 16), bytes ? (const void*)(src + u * 128) : (const void*)x, bytes);
                }
            }
        }
        cp_async_commit();  // (an empty group when there is nothing left, so that the wait counts stay aligned)
    };

#pragma unroll
    for (int k = 0; k < S - 1; k++) issue(k);
    float cs0 = id, cs1 = id, cs2 = id, cs3 = id, cx0 = id, cx1 = id, cx2 = id, cx3 = id;
    for (uint32_t it = 0; it < n_it; it++) {
        cp_async_wait<S - 2>();  // this lane's copies of batch `it` have landed
        const float4* const stage = ring + (it % S) * (B * 32);
        uint32_t row, r0;
        locate(it, row, r0);  // r0 = 128-element row index of the batch's first row
        const uint32_t batch = it % NB;
        // The stage that batch it + S - 1 lands in is the one read in the PREVIOUS iteration; those reads have been
        // consumed by now (in-order issue), so it is free.
        issue(it + S - 1);
        // rows r0-1 .. r0+B-2 are emitted by this batch (row r needs the prefixes of row r+1); `drow` points at this
        // lane's vector of row r0-1
        float* const drow = out + (size_t)row * nout + 4 * lane + ((long long)r0 - 1) * 128;
        const bool interior = (r0 + B - 1) * 128 <= nout && (OP == AL_RSUM || ((r0 + B) * 128 <= n_in && batch != NB - 1));

        auto body = [&](auto interior_tag) {
            constexpr bool INTERIOR = decltype(interior_tag)::value;  // whole rows everywhere, nothing to mask
#pragma unroll
            for (int g = 0; g < B; g += IL) {
                float p2[IL], p3[IL], s0[IL], s1[IL], s2[IL], vu[IL], vd[IL], E[IL], F[IL];
                float4 cur[IL];
#pragma unroll
                for (int i = 0; i < IL; i++) cur[i] = stage[(g + i) * 32];
                if constexpr (!INTERIOR && OP != AL_RSUM) {  // zero fill is the identity of a sum only
#pragma unroll
                    for (int i = 0; i < IL; i++) {
                        const uint32_t e = (r0 + g + i) * 128 + 4 * lane;
                        const bool idle = g + i == B - 1 && batch == NB - 1 && !low;
                        if (idle || e + 0 >= n_in) cur[i].x = id;
                        if (idle || e + 1 >= n_in) cur[i].y = id;
                        if (idle || e + 2 >= n_in) cur[i].z = id;
                        if (idle || e + 3 >= n_in) cur[i].w = id;
                    }
                }
#pragma unroll
                for (int i = 0; i < IL; i++) {
                    const float4 xv = cur[i];
                    p2[i] = racc<OP>(xv.x, xv.y), p3[i] = racc<OP>(p2[i], xv.z);
                    s2[i] = racc<OP>(xv.z, xv.w), s1[i] = racc<OP>(xv.y, s2[i]), s0[i] = racc<OP>(xv.x, s1[i]);
                    vu[i] = vd[i] = racc<OP>(p3[i], xv.w);
                    E[i] = F[i] = id;
                }
                // exclusive prefix E / suffix F of the lane totals inside each block of G lanes (Kogge-Stone; the
                // inclusive values vu / vd only feed the next step)
#pragma unroll
                for (int k = 0; k < STEPS; k++) {
                    float ou[IL], od[IL];
#pragma unroll
                    for (int i = 0; i < IL; i++) {
                        ou[i] = __shfl_up_sync(0xffffffffu, vu[i], 1 << k, G);
                        od[i] = __shfl_down_sync(0xffffffffu, vd[i], 1 << k, G);
                    }
#pragma unroll
                    for (int i = 0; i < IL; i++) {
                        if (pu[k]) vu[i] = racc<OP>(ou[i], vu[i]), E[i] = racc<OP>(ou[i], E[i]);
                        if (pd[k]) vd[i] = racc<OP>(vd[i], od[i]), F[i] = racc<OP>(F[i], od[i]);
                    }
                }
#pragma unroll
                for (int i = 0; i < IL; i++) {
                    const float4 xv = cur[i];
                    const float nx0 = E[i], nx1 = racc<OP>(E[i], xv.x), nx2 = racc<OP>(E[i], p2[i]), nx3 = racc<OP>(E[i], p3[i]);
                    if (g + i > 0 || batch > 0) {  // emit the previous row: its own suffixes (+) the next block's exclusive prefixes
                        float4 o;
                        if constexpr (G == 32) {
                            o.x = nx0, o.y = nx1, o.z = nx2, o.w = nx3;
                        } else {
                            o.x = __shfl_sync(0xffffffffu, low ? nx0 : cx0, partner);
                            o.y = __shfl_sync(0xffffffffu, low ? nx1 : cx1, partner);
                            o.z = __shfl_sync(0xffffffffu, low ? nx2 : cx2, partner);
                            o.w = __shfl_sync(0xffffffffu, low ? nx3 : cx3, partner);
                        }
                        o.x = racc<OP>(cs0, o.x), o.y = racc<OP>(cs1, o.y), o.z = racc<OP>(cs2, o.z), o.w = racc<OP>(cs3, o.w);
                        if (!ident) {  // with the identity as init the fold changes nothing: no sum here is -0, no max/min is NaN
                            o.x = rfinal<OP>(init, o.x), o.y = rfinal<OP>(init, o.y), o.z = rfinal<OP>(init, o.z), o.w = rfinal<OP>(init, o.w);
                        }
                        float* const d = drow + (g + i) * 128;
                        if constexpr (INTERIOR) {
                            __stcs(reinterpret_cast<float4*>(d), o);
                        } else {
                            const uint32_t el = (r0 + g + i - 1) * 128 + 4 * lane;  // this lane's first output of the emitted row
                            if (el + 4 <= nout) {
                                __stcs(reinterpret_cast<float4*>(d), o);
                            } else {
                                if (el + 0 < nout) __stcs(d + 0, o.x);
                                if (el + 1 < nout) __stcs(d + 1, o.y);
                                if (el + 2 < nout) __stcs(d + 2, o.z);
                            }
                        }
                    }
                    cs0 = racc<OP>(s0[i], F[i]), cs1 = racc<OP>(s1[i], F[i]), cs2 = racc<OP>(s2[i], F[i]), cs3 = racc<OP>(xv.w, F[i]);
                    cx0 = nx0, cx1 = nx1, cx2 = nx2, cx3 = nx3;
                }
            }
        };
        if (interior) body(std::true_type{});
        else body(std::false_type{});
    }
    cp_async_wait<0>();
}

template <int OP, int W, int B, int S, int NB, int WPC>
static bool launch_window_shfl_cfg(const float* src, float* out, uint32_t nout, uint32_t rows, long long row_stride, float init, long per_sm) {
    const float ident_tab[AL_REDOP_COUNT] = {0.0f, -INFINITY, INFINITY, 1.0f};
    const int init_is_identity = memcmp(&init, &ident_tab[OP], sizeof(float)) == 0;
    constexpr size_t smem = (size_t)WPC * S * B * 512;
    constexpr int R = B * NB - 1;
    const uint64_t rows128 = ((uint64_t)nout + 127) / 128;
    const uint64_t segs_per_row = (rows128 + R - 1) / R;
    const uint64_t all_segs = segs_per_row * rows;
    if (all_segs >= (1ull << 31)) return set_error("NotImplementedError: too many window segments"), false;
    static bool configured = false;
    if (!configured) {
        if (!cuda_ok(cudaFuncSetAttribute(reduce_window_shfl_kernel<OP, W, B, S, NB, WPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute"))
            return false;
        configured = true;
    }
    uint64_t blocks = (all_segs + WPC - 1) / WPC;
    if (blocks > (uint64_t)kNumSMs * per_sm) blocks = (uint64_t)kNumSMs * per_sm;
    reduce_window_shfl_kernel<OP, W, B, S, NB, WPC><<<(unsigned)blocks, WPC * 32, smem, rt().stream>>>(src, out, nout, (uint32_t)all_segs, (uint32_t)segs_per_row, row_stride, init, init_is_identity);
    note_launch("reduce_window_shfl");
    return check_launch("reduce_window_shfl");
}

template <int OP, int W>
static bool launch_window_shfl(const float* src, float* out, uint32_t nout, uint32_t rows, long long row_stride, float init) {
    // Ring shape (rows per batch, stages, batches per segment, warps per CTA) and CTAs per SM. Measured on C5a inside a
    // busy stretch (power cap active, tools/microbench.py c5a_attrib): 16 x 2 x 1 with 6 resident warps per SM and
    // 12 x 2 x 1 with 8 both run at 0.335 ms (6.4 TB/s); more resident warps are SLOWER (10-12 warps: 0.36-0.40 ms).
    const long v = rt().window_shfl;
    const long per_sm = rt().window_ctas_per_sm > 0 ? rt().window_ctas_per_sm : (v == 1 ? 3 : 2);
    switch (v) {
        case 2: return launch_window_shfl_cfg<OP, W, 12, 2, 1, 4>(src, out, nout, rows, row_stride, init, per_sm);
        case 3: return launch_window_shfl_cfg<OP, W, 16, 2, 2, 4>(src, out, nout, rows, row_stride, init, per_sm);
        default: return launch_window_shfl_cfg<OP, W, 16, 2, 1, 2>(src, out, nout, rows, row_stride, init, per_sm);
    }
}

template <int OP, int W>
static bool launch_window_fixed(const float* src, float* out, uint32_t nout, uint32_t rows, long long row_stride, float init) {
    const float ident_tab[AL_REDOP_COUNT] = {0.0f, -INFINITY, INFINITY, 1.0f};
    const int init_is_identity = memcmp(&init, &ident_tab[OP], sizeof(float)) == 0;
    constexpr int kWinSpan = kWinFixedThreads * kWinLoads;
    constexpr int TILE = kWinSpan - W;
    const bool split_halves = !rt().window_whole_blocks;
    const size_t smem = 2ull * (kWinSpan / W) * (split_halves ? W + 2 : W + 1) * sizeof(float);
    static bool configured = false;
    if (!configured) {
        if (!cuda_ok(cudaFuncSetAttribute(reduce_window_fixed_kernel<OP, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem + 8192),
                     "cudaFuncSetAttribute") ||
            !cuda_ok(cudaFuncSetAttribute(reduce_window_half_kernel<OP, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem + 8192),
                     "cudaFuncSetAttribute"))
            return false;
        configured = true;
    }
    const unsigned tiles_per_row = (unsigned)(((uint64_t)nout + TILE - 1) / TILE);
    const uint64_t all_tiles = (uint64_t)tiles_per_row * rows;
    if (all_tiles >= (1ull << 31)) return set_error("NotImplementedError: too many window tiles"), false;
    const unsigned n_tiles = (unsigned)all_tiles;
    const long per_sm = rt().window_ctas_per_sm > 0 ? rt().window_ctas_per_sm : 32;  // 4 resident, the rest queue (tail balance)
    unsigned blocks = (unsigned)(kNumSMs * per_sm);
    if (blocks > n_tiles) blocks = n_tiles;
    if (split_halves)
        reduce_window_half_kernel<OP, W><<<blocks, kWinFixedThreads, smem, rt().stream>>>(src, out, nout, n_tiles, tiles_per_row, row_stride, init, init_is_identity);
    else
        reduce_window_fixed_kernel<OP, W><<<blocks, kWinFixedThreads, smem, rt().stream>>>(src, out, nout, n_tiles, tiles_per_row, row_stride, init, init_is_identity);
    note_launch("reduce_window");
    return check_launch("reduce_window");
}

template <int OP>
static bool launch_window_any(const float* src, float* out, uint32_t nout, uint32_t w, uint32_t rows, long long row_stride, float init) {
    const float ident_tab[AL_REDOP_COUNT] = {0.0f, -INFINITY, INFINITY, 1.0f};
    const int init_is_identity = memcmp(&init, &ident_tab[OP], sizeof(float)) == 0;
    const bool small = w < 24 && rt().window_any_l16;  // L = 16 elements per lane (tile 2048), else 32 (tile 4096)
    const uint32_t span = (uint32_t)kAnyT * (small ? 16u : 32u);
    const uint32_t tile = span - w;
    const unsigned tiles_per_row = (unsigned)(((uint64_t)nout + tile - 1) / tile);
    const uint64_t all_tiles = (uint64_t)tiles_per_row * rows;
    if (all_tiles >= (1ull << 31)) return set_error("NotImplementedError: too many window tiles"), false;
    const unsigned n_tiles = (unsigned)all_tiles;
    const long per_sm = rt().window_ctas_per_sm > 0 ? rt().window_ctas_per_sm : 32;
    unsigned blocks = (unsigned)(kNumSMs * per_sm);
    if (blocks > n_tiles) blocks = n_tiles;
    const size_t smem = 2ull * (span + span / 32) * sizeof(float);
    if (small) reduce_window_any_kernel<OP, 16><<<blocks, kAnyT, smem, rt().stream>>>(src, out, nout, w, n_tiles, tiles_per_row, row_stride, init, init_is_identity);
    else reduce_window_any_kernel<OP, 32><<<blocks, kAnyT, smem, rt().stream>>>(src, out, nout, w, n_tiles, tiles_per_row, row_stride, init, init_is_identity);
    note_launch("reduce_window_any");
    return check_launch("reduce_window_any");
}

// ---- host planning ------------------------------------------------------------------------------------
struct AxisList {
    int n = 0;
    uint64_t extent[kMaxDims];
    int64_t stride[kMaxDims];
    bool push(uint64_t e, int64_t s) {  // called innermost first; merges when contiguous
        if (n > 0 && s == stride[n - 1] * (int64_t)extent[n - 1]) {
            extent[n - 1] *= e;
            return true;
        }
        if (n == kMaxDims) return false;
        extent[n] = e;
        stride[n] = s;
        n++;
        return true;
    }
    uint64_t count() const {
        uint64_t c = 1;
        for (int i = 0; i < n; i++) c *= extent[i];
        return c;
    }
};

static uint32_t pow2_ceil(uint32_t v) {
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

template <int OP>
static bool run_reduce(const float* src, float* out, const AxisList& K, const AxisList& R, float init,
                       const PeerMap* peer = nullptr, bool* peer_used = nullptr) {
    Runtime& r = rt();
    if (peer) {
        // Only the output-contiguous plan carries the peer epilogue (a reduction over the sharded leading
        // axis keeps a unit-stride inner axis); anything else is reduced locally and pushed afterwards.
        *peer_used = false;
        const bool outer_plan = K.n > 0 && K.stride[0] == 1 && !(R.n > 0 && R.stride[0] == 1);
        if (r.reduce_sequential || !outer_plan) return true;
    }
    cudaStream_t st = r.stream;
    const uint64_t nout = K.count(), rtot = R.count();
    if (nout >= (1ull << 31) || rtot >= (1ull << 31)) {
        set_error("NotImplementedError: reduce with >= 2^31 outputs or reduced elements per output");
        return false;
    }
    RedArgs a;
    memset(&a, 0, sizeof a);
    a.src = src;
    a.nout = (uint32_t)nout;
    a.init = init;
    a.nk = K.n;
    a.nr = R.n;
    const bool general = r.force_general;
    if (r.reduce_sequential) {
        for (int d = 0; d < K.n; d++) {
            a.kext[d] = (uint32_t)K.extent[d];
            a.kstride[d] = K.stride[d];
            a.kdiv[d] = make_fastdiv(a.kext[d]);
        }
        for (int d = 0; d < R.n; d++) {
            a.rext[d] = (uint32_t)R.extent[d];
            a.rstride[d] = R.stride[d];
            a.rdiv[d] = make_fastdiv(a.rext[d]);
        }
        if (R.n == 0) {
            a.nr = 1;
            a.rext[0] = 1;
            a.rstride[0] = 0;
            a.rdiv[0] = make_fastdiv(1);
        }
        a.nr_items = (uint32_t)rtot;
        a.dst = out;
        reduce_sequential_kernel<OP><<<(unsigned)((nout + 255) / 256), 256, 0, st>>>(a);
        note_launch("reduce_sequential");
        return check_launch("reduce_sequential");
    }

    // -- overlapping windows (C5a): kept stride == reduced stride == 1, optionally a batch of rows --
    const bool win_axes = R.n == 1 && R.stride[0] == 1 && (K.n == 1 || K.n == 2) && K.stride[0] == 1;
    if (!general && !r.disable_window && win_axes && R.extent[0] >= 8 && R.extent[0] <= 1024 && K.extent[0] >= 4 * R.extent[0]) {
        const uint32_t w = (uint32_t)R.extent[0];
        const uint32_t per_row = (uint32_t)K.extent[0];
        const uint32_t rows = K.n == 2 ? (uint32_t)K.extent[1] : 1;
        const long long row_stride = K.n == 2 ? K.stride[1] : 0;
        // 128-bit rows on both sides: the register/shuffle kernel (no shared memory)
        const bool rows_aligned = (uintptr_t)src % 16 == 0 && (uintptr_t)out % 16 == 0 &&
                                  (rows == 1 || (per_row % 4 == 0 && row_stride % 4 == 0));
        if (r.window_shfl && rows_aligned) {
            switch (w) {
                case 8: return launch_window_shfl<OP, 8>(src, out, per_row, rows, row_stride, init);
                case 16: return launch_window_shfl<OP, 16>(src, out, per_row, rows, row_stride, init);
                case 32: return launch_window_shfl<OP, 32>(src, out, per_row, rows, row_stride, init);
                case 64: return launch_window_shfl<OP, 64>(src, out, per_row, rows, row_stride, init);
                case 128: return launch_window_shfl<OP, 128>(src, out, per_row, rows, row_stride, init);
                default: break;
            }
        }
        switch (w) {
            case 8: return launch_window_fixed<OP, 8>(src, out, per_row, rows, row_stride, init);
            case 16: return launch_window_fixed<OP, 16>(src, out, per_row, rows, row_stride, init);
            case 32: return launch_window_fixed<OP, 32>(src, out, per_row, rows, row_stride, init);
            case 64: return launch_window_fixed<OP, 64>(src, out, per_row, rows, row_stride, init);
            default: return launch_window_any<OP>(src, out, per_row, w, rows, row_stride, init);
        }
    }

    // -- dense [nout, R] with a short reduced inner axis --
    if (!general && K.n == 1 && R.n == 1 && R.stride[0] == 1 && R.extent[0] >= 2 && R.extent[0] <= 8 &&
        K.stride[0] == (int64_t)R.extent[0] && nout % 4 == 0 && (uintptr_t)src % 16 == 0 && (uintptr_t)out % 16 == 0) {
        switch (R.extent[0]) {
            case 2: return launch_rows_packed<OP, 2>(src, out, nout, init);
            case 3: return launch_rows_packed<OP, 3>(src, out, nout, init);
            case 4: return launch_rows_packed<OP, 4>(src, out, nout, init);
            case 5: return launch_rows_packed<OP, 5>(src, out, nout, init);
            case 6: return launch_rows_packed<OP, 6>(src, out, nout, init);
            case 7: return launch_rows_packed<OP, 7>(src, out, nout, init);
            default: return launch_rows_packed<OP, 8>(src, out, nout, init);
        }
    }

    const bool outer = K.n > 0 && K.stride[0] == 1 && !(R.n > 0 && R.stride[0] == 1 && K.extent[0] < 8);
    const bool aligned = (uintptr_t)src % 16 == 0;
    for (int d = 0; d < K.n; d++) {
        a.kext[d] = (uint32_t)K.extent[d];
        a.kstride[d] = K.stride[d];
    }
    for (int d = 0; d < R.n; d++) {
        a.rext[d] = (uint32_t)R.extent[d];
        a.rstride[d] = R.stride[d];
    }
    if (R.n == 0) {  // nothing to reduce: one pass that only applies the init fold
        a.nr = 1;
        a.rext[0] = 1;
        a.rstride[0] = 0;
    }
    auto all_mult4 = [](const AxisList& L, int from) {
        for (int d = from; d < L.n; d++)
            if (L.stride[d] % 4 != 0) return false;
        return true;
    };
    const int target_ctas = kNumSMs * 16;  // two waves of 8 CTAs per SM: measured 7.05 TB/s vs 6.55 for exactly one wave (C4 dims (1,2))
    float* partial = nullptr;

    if (outer) {
        const bool vec = !general && aligned && K.extent[0] % 4 == 0 && all_mult4(K, 1) && all_mult4(R, 0);
        const int VEC = vec ? 4 : 1;
        a.kext[0] = (uint32_t)(K.extent[0] / VEC);
        a.kstride[0] = VEC;
        a.nk_items = (uint32_t)(nout / VEC);
        a.nr_items = (uint32_t)rtot;
        for (int d = 0; d < kMaxDims; d++) {
            a.kdiv[d] = make_fastdiv(a.kext[d] ? a.kext[d] : 1);
            a.rdiv[d] = make_fastdiv(a.rext[d] ? a.rext[d] : 1);
        }
        // TX lanes along the kept axis (no wider than the kept extent), the rest of the CTA along
        // the reduced axis
        unsigned TX = 256, TY = 1;
        if (rtot >= 64) {
            TX = pow2_ceil(a.nk_items);
            if (TX > 32) TX = 32;
            TY = 256 / TX;
        }
        const unsigned gx = (a.nk_items + TX - 1) / TX;
        unsigned splits = 1;
        if (r.reduce_splits > 0) splits = (unsigned)r.reduce_splits;
        else if (gx < (unsigned)target_ctas) splits = (target_ctas + gx - 1) / gx;
        unsigned max_splits = (unsigned)((rtot + TY * 8 - 1) / (TY * 8));
        if (splits > max_splits) splits = max_splits ? max_splits : 1;
        if (splits > 65535) splits = 65535;
        a.r_per_split = (uint32_t)((rtot + splits - 1) / splits);
        splits = (unsigned)((rtot + a.r_per_split - 1) / a.r_per_split);
        a.finalize = splits == 1 && !peer;
        PeerMap finish_peer;
        memset(&finish_peer, 0, sizeof finish_peer);
        if (splits > 1) {
            partial = (float*)device_alloc((size_t)splits * nout * sizeof(float));
            if (!partial) return false;
            a.dst = partial;
            if (peer) finish_peer = *peer;
        } else {
            a.dst = out;
            if (peer) {
                a.peer = *peer;
                a.rot = (uint32_t)(((uint64_t)peer->rank * gx) / (uint64_t)peer->world);
            }
        }
        if (peer) *peer_used = true;
        size_t smem = TY > 1 ? (size_t)256 * sizeof(float) * VEC : 0;
        dim3 grid(gx, splits), block(TX, TY);
        if (vec) reduce_outer_kernel<OP, 4><<<grid, block, smem, st>>>(a);
        else reduce_outer_kernel<OP, 1><<<grid, block, smem, st>>>(a);
        note_launch(vec ? "reduce_outer_vec4" : "reduce_outer_scalar");
        if (!check_launch("reduce_outer")) return false;
        if (splits > 1) {
            unsigned fb = nout <= 2048 ? (unsigned)((nout * 32 + 255) / 256) : (unsigned)((nout + 255) / 256);
            reduce_finish_kernel<OP><<<fb, 256, 0, st>>>(partial, out, (uint32_t)nout, splits, init, finish_peer);
            note_launch("reduce_finish");
            device_free(partial);
            return check_launch("reduce_finish");
        }
        return true;
    }

    // inner (team) path; also the generic fallback when nothing is unit-stride
    const bool vec = !general && aligned && R.n > 0 && R.stride[0] == 1 && R.extent[0] % 4 == 0 && all_mult4(R, 1) &&
                     all_mult4(K, 0);
    const int VEC = vec ? 4 : 1;
    if (vec) {
        a.rext[0] = (uint32_t)(R.extent[0] / 4);
        a.rstride[0] = 4;
    }
    a.nk_items = (uint32_t)nout;
    a.nr_items = (uint32_t)(rtot / VEC);
    for (int d = 0; d < kMaxDims; d++) {
        a.kdiv[d] = make_fastdiv(a.kext[d] ? a.kext[d] : 1);
        a.rdiv[d] = make_fastdiv(a.rext[d] ? a.rext[d] : 1);
    }
    const bool cta_team = a.nr_items >= 4096 && nout * 32 < (uint64_t)target_ctas * 256;
    unsigned splits = 1;
    if (cta_team) {
        if (r.reduce_splits > 0) splits = (unsigned)r.reduce_splits;
        else if (nout < (uint64_t)target_ctas) splits = (unsigned)((target_ctas + nout - 1) / nout);
        unsigned max_splits = (a.nr_items + 2047) / 2048;
        if (splits > max_splits) splits = max_splits;
        if (splits > 65535) splits = 65535;
        if (splits < 1) splits = 1;
    }
    a.r_per_split = (a.nr_items + splits - 1) / splits;
    splits = (a.nr_items + a.r_per_split - 1) / a.r_per_split;
    a.finalize = splits == 1;
    if (splits > 1) {
        partial = (float*)device_alloc((size_t)splits * nout * sizeof(float));
        if (!partial) return false;
        a.dst = partial;
    } else {
        a.dst = out;
    }
    if (cta_team) {
        dim3 grid((unsigned)nout, splits);
        if (vec) reduce_inner_cta_kernel<OP, 4><<<grid, 256, 0, st>>>(a);
        else reduce_inner_cta_kernel<OP, 1><<<grid, 256, 0, st>>>(a);
        note_launch("reduce_inner_cta");
    } else if (!vec && !general && r.reduce_peel && R.n > 0 && R.stride[0] == 1 && R.extent[0] >= 16 && R.extent[0] < (1u << 29)) {
        // unit-stride rows the aligned 128-bit path cannot take: masked aligned vectors (see the kernel)
        const uint32_t nvec = (uint32_t)(R.extent[0] / 4 + 1);
        const uint32_t per_lane = r.reduce_loads_per_lane > 0 ? (uint32_t)r.reduce_loads_per_lane : 4;
        uint32_t team = pow2_ceil((nvec + per_lane - 1) / per_lane);
        if (team > 32) team = 32;
        unsigned gx = (unsigned)(((uint64_t)nout * team + 255) / 256);
        switch (team) {
            case 1: reduce_inner_peel_kernel<OP, 1><<<gx, 256, 0, st>>>(a); break;
            case 2: reduce_inner_peel_kernel<OP, 2><<<gx, 256, 0, st>>>(a); break;
            case 4: reduce_inner_peel_kernel<OP, 4><<<gx, 256, 0, st>>>(a); break;
            case 8: reduce_inner_peel_kernel<OP, 8><<<gx, 256, 0, st>>>(a); break;
            case 16: reduce_inner_peel_kernel<OP, 16><<<gx, 256, 0, st>>>(a); break;
            default: reduce_inner_peel_kernel<OP, 32><<<gx, 256, 0, st>>>(a); break;
        }
        note_launch("reduce_inner_peel");
    } else {
        // lanes per output: ~4 loads per lane for 128-bit loads, ~32 for 4-byte loads (measured: 2^20 rows of
        // 255 floats 3.35 -> 4.48 TB/s; 2^24 rows of 64 floats best at 4), never wider than the inner axis
        const uint32_t per_lane = r.reduce_loads_per_lane > 0 ? (uint32_t)r.reduce_loads_per_lane : (vec ? 4 : 32);
        uint32_t team = pow2_ceil((a.nr_items + per_lane - 1) / per_lane);
        if (team > 32) team = 32;
        while (team > 1 && team > pow2_ceil(a.rext[0])) team >>= 1;
        unsigned gx = (unsigned)(((uint64_t)nout * team + 255) / 256);
#define AL_TEAM_CASE(T)                                                               \
    case T:                                                                           \
        if (vec) reduce_inner_kernel<OP, 4, T><<<dim3(gx, 1), 256, 0, st>>>(a);       \
        else reduce_inner_kernel<OP, 1, T><<<dim3(gx, 1), 256, 0, st>>>(a);           \
        break;
        switch (team) {
            AL_TEAM_CASE(1)
            AL_TEAM_CASE(2)
            AL_TEAM_CASE(4)
            AL_TEAM_CASE(8)
            AL_TEAM_CASE(16)
            AL_TEAM_CASE(32)
        }
#undef AL_TEAM_CASE
        note_launch(vec ? "reduce_inner_vec4" : "reduce_inner_scalar");
    }
    if (!check_launch("reduce_inner")) return false;
    if (splits > 1) {
        unsigned fb = nout <= 2048 ? (unsigned)((nout * 32 + 255) / 256) : (unsigned)((nout + 255) / 256);
        PeerMap no_peer;
        memset(&no_peer, 0, sizeof no_peer);
        reduce_finish_kernel<OP><<<fb, 256, 0, st>>>(partial, out, (uint32_t)nout, splits, init, no_peer);
        note_launch("reduce_finish");
        device_free(partial);
        return check_launch("reduce_finish");
    }
    return true;
}

// Split the source axes into the kept list K and the reduced list R (each innermost first, merged
// where contiguous) and produce the keepdims output shape.
static bool build_axes(int redop, const NDArray* src, const size_t* dims, size_t ndims, AxisList& K, AxisList& R,
                       size_t* dshape, bool* red) {
    AL_REQUIRE(redop >= 0 && redop < AL_REDOP_COUNT, "ValueError: unknown reduce op %d", redop);
    const size_t nd = src->lay->ndim;
    AL_REQUIRE(nd <= 64, "ValueError: too many dimensions");
    for (size_t i = 0; i < 64; i++) red[i] = false;
    for (size_t i = 0; i < ndims; i++) {
        AL_REQUIRE(dims[i] < nd, "ValueError: out of bounds");
        red[dims[i]] = true;
    }
    for (size_t i = 0; i < nd; i++) dshape[i] = red[i] ? 1 : src->lay->shape[i];
    // Kept axes stay in logical order (they define the output layout). The order of the REDUCED axes is
    // free for the tree kernels, so they are visited by ascending source stride: a transposed or permuted
    // source then still merges into long unit-stride runs (a full reduce of X.T is one contiguous stream;
    // in logical order it was a stride-N gather at 0.25 TB/s). The sequential verification mode keeps the
    // logical order, because there the order IS the point.
    std::vector<size_t> red_axes;
    for (size_t ax = nd; ax-- > 0;) {  // innermost first
        uint64_t e = src->lay->shape[ax];
        if (e == 1) continue;
        if (red[ax]) {
            red_axes.push_back(ax);
            continue;
        }
        AL_REQUIRE(K.push(e, (int64_t)src->lay->stride[ax]), "NotImplementedError: reduce layout needs more than %d non-mergeable axes", kMaxDims);
    }
    if (!rt().reduce_sequential)
        std::stable_sort(red_axes.begin(), red_axes.end(),
                         [&](size_t x, size_t y) { return src->lay->stride[x] < src->lay->stride[y]; });
    for (size_t ax : red_axes)
        AL_REQUIRE(R.push(src->lay->shape[ax], (int64_t)src->lay->stride[ax]), "NotImplementedError: reduce layout needs more than %d non-mergeable axes", kMaxDims);
    return true;
}

// ---- sign of a zero max/min result (arraylib.c:18-19, 737-744) -----------------------------------------------
// fmax/fmin keep their LEFT operand when both compare equal, so the reference's sequential fold returns the FIRST
// zero in logical order when the extreme value of a reduced set is zero and the set holds both +0 and -0 (checked
// against the compiled reference: max of [-0, +0] is -0, of [+0, -0] is +0; same for min). The tree kernels use the
// hardware FMNMX (max -> +0, min -> -0 whatever the order). This pass repairs exactly those outputs: one warp per
// output reads the result; only when it is a zero does the warp walk the reduced elements in the reference's logical
// order, 32 at a time, and stop at the first zero. Non-zero outputs cost one 4-byte read each.
constexpr int kZeroFixDims = 16;
struct ZeroFixArgs {
    int nkept, nred;
    uint64_t kext[kZeroFixDims], rext[kZeroFixDims];      // logical order, OUTERMOST first
    int64_t kstride[kZeroFixDims], rstride[kZeroFixDims];
    uint64_t nout, count;
};

// Zero-sign repair. A warp looks at 256 outputs per trip (two 128-bit loads per lane, both in flight) and only for a ZERO
// result walks that output's reduced elements, 32 at a time, up to the first zero. (The first version gave every output a
// whole warp, i.e. one dependent 4-byte load per warp and trip: 26 ms behind a 0.5 ms window reduce_max with 2^28 outputs.)
__device__ __forceinline__ void zero_sign_fix_one(const float* __restrict__ src, float* __restrict__ out, const ZeroFixArgs& a,
                                                  uint64_t o, float init, int init_is_zero, unsigned lane) {
    if (init_is_zero) {  // a zero init is the first zero of the fold
        if (lane == 0) out[o] = init;
        return;
    }
    int64_t base = 0;
    uint64_t rem = o;
    for (int d = a.nkept - 1; d >= 0; d--) {
        base += (int64_t)(rem % a.kext[d]) * a.kstride[d];
        rem /= a.kext[d];
    }
    for (uint64_t r0 = 0; r0 < a.count; r0 += 32) {
        const uint64_t r = r0 + lane;
        float x = 1.0f;
        if (r < a.count) {
            int64_t off = base;
            uint64_t q = r;
            for (int d = a.nred - 1; d >= 0; d--) {
                off += (int64_t)(q % a.rext[d]) * a.rstride[d];
                q /= a.rext[d];
            }
            x = src[off];
        }
        const unsigned hit = __ballot_sync(0xffffffffu, x == 0.0f);
        if (hit) {
            const float first = __shfl_sync(0xffffffffu, x, __ffs(hit) - 1);
            if (lane == 0) out[o] = first;
            return;
        }
    }
}

__global__ void __launch_bounds__(256)
reduce_zero_sign_kernel(const float* __restrict__ src, float* __restrict__ out, const ZeroFixArgs a, float init, int init_is_zero) {
    const unsigned lane = threadIdx.x & 31;
    const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    const uint64_t warp = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    // whole groups of 256 outputs through 128-bit loads (the output block is 16-byte aligned: a fresh allocation)
    const uint64_t groups = (reinterpret_cast<uintptr_t>(out) & 15) == 0 ? a.nout / 256 : 0;
    for (uint64_t g = warp; g < groups; g += warps) {
        const uint64_t o0 = g * 256;
        float4 q[2];
#pragma unroll
        for (int u = 0; u < 2; u++) q[u] = *reinterpret_cast<const float4*>(out + o0 + (u * 32 + lane) * 4);
        const bool mine = q[0].x == 0.0f || q[0].y == 0.0f || q[0].z == 0.0f || q[0].w == 0.0f ||
                          q[1].x == 0.0f || q[1].y == 0.0f || q[1].z == 0.0f || q[1].w == 0.0f;  // (NaN compares unequal too)
        unsigned owners = __ballot_sync(0xffffffffu, mine);
        while (owners) {  // rare: some lane holds a zero result
            const int l = __ffs(owners) - 1;
            owners &= owners - 1;
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const float c[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (__shfl_sync(0xffffffffu, c[k], l) == 0.0f) zero_sign_fix_one(src, out, a, o0 + (u * 32 + l) * 4 + k, init, init_is_zero, lane);
            }
        }
    }
    // the rest (and unaligned outputs): 32 outputs per trip
    for (uint64_t o0 = groups * 256 + warp * 32; o0 < a.nout; o0 += warps * 32) {
        const uint64_t o = o0 + lane;
        const float v = o < a.nout ? out[o] : 1.0f;
        unsigned zeros = __ballot_sync(0xffffffffu, v == 0.0f);
        while (zeros) {
            const int l = __ffs(zeros) - 1;
            zeros &= zeros - 1;
            zero_sign_fix_one(src, out, a, o0 + l, init, init_is_zero, lane);
        }
    }
}

// Returns false only on a launch failure. Layouts with more than kZeroFixDims non-unit axes keep the tree's sign.
static bool fix_zero_sign(const NDArray* src, float* out, const bool* red, float init, bool custom_init) {
    ZeroFixArgs a;
    memset(&a, 0, sizeof a);
    a.nout = a.count = 1;
    for (size_t i = 0; i < src->lay->ndim; i++) {
        const uint64_t e = src->lay->shape[i];
        if (e == 1) continue;
        if (red[i]) {
            if (a.nred == kZeroFixDims) return true;
            a.rext[a.nred] = e;
            a.rstride[a.nred++] = (int64_t)src->lay->stride[i];
            a.count *= e;
        } else {
            if (a.nkept == kZeroFixDims) return true;
            a.kext[a.nkept] = e;
            a.kstride[a.nkept++] = (int64_t)src->lay->stride[i];
            a.nout *= e;
        }
    }
    const uint64_t blocks = (a.nout + 8 * 256 - 1) / (8 * 256);  // 8 warps per CTA, 256 outputs per warp and trip
    const unsigned grid = (unsigned)std::min<uint64_t>(blocks, (uint64_t)kNumSMs * 8);
    reduce_zero_sign_kernel<<<grid, 256, 0, rt().stream>>>(src->ptr, out, a, init, custom_init && init == 0.0f);
    rt().launches++;  // counted, but al_last_kernel() keeps naming the reduction kernel the plan chose
    return check_launch("reduce_zero_sign");
}

NDArray* reduce_impl(int redop, const NDArray* src, const size_t* dims, size_t ndims, float init, bool custom_init) {
    if (!valid(src, "array_reduce")) return nullptr;
    const size_t nd = src->lay->ndim;
    bool red[64];
    size_t dshape[64];
    AxisList K, R;
    if (!build_axes(redop, src, dims, ndims, K, R, dshape, red)) return nullptr;
    if (R.count() >= (1ull << 31)) {
        // More than 2^31-1 reduced elements per output (e.g. a full reduce of config C3's 2^32-element
        // result): halve the widest reduced axis, reduce both halves and combine them elementwise.
        AL_REQUIRE(!custom_init, "NotImplementedError: reduce over >= 2^31 elements with a custom init");
        size_t ax = nd;
        for (size_t i = 0; i < nd; i++)
            if (red[i] && src->lay->shape[i] > 1 && (ax == nd || src->lay->shape[i] > src->lay->shape[ax])) ax = i;
        AL_REQUIRE(ax < nd, "NotImplementedError: cannot split this reduction");
        size_t lo_shape[64], hi_shape[64];
        for (size_t i = 0; i < nd; i++) lo_shape[i] = hi_shape[i] = src->lay->shape[i];
        lo_shape[ax] = src->lay->shape[ax] / 2;
        hi_shape[ax] = src->lay->shape[ax] - lo_shape[ax];
        NDArray* lo = new_view(src, lo_shape, src->lay->stride, nd, 0);
        NDArray* hi = new_view(src, hi_shape, src->lay->stride, nd, lo_shape[ax] * src->lay->stride[ax]);
        NDArray* rlo = reduce_impl(redop, lo, dims, ndims, init, false);
        NDArray* rhi = rlo ? reduce_impl(redop, hi, dims, ndims, init, false) : nullptr;
        NDArray* both = nullptr;
        if (rlo && rhi) {
            const int combine[AL_REDOP_COUNT] = {AL_SUM, AL_MAX, AL_MIN, AL_MUL};
            both = array_array_op_named(combine[redop], rlo, rhi);
        }
        std::string keep = last_error_string();
        for (NDArray* t : {lo, hi, rlo, rhi})
            if (t) release_array(t);
        restore_error(keep);
        return both;
    }
    NDArray* out = new_array(dshape, nd);
    if (!out) return nullptr;
    if (!custom_init) {
        const float ident[AL_REDOP_COUNT] = {0.0f, -INFINITY, INFINITY, 1.0f};
        init = ident[redop];
    }
    bool ok = false;
    switch (redop) {
        case AL_RSUM: ok = run_reduce<AL_RSUM>(src->ptr, out->ptr, K, R, init); break;
        case AL_RMAX: ok = run_reduce<AL_RMAX>(src->ptr, out->ptr, K, R, init); break;
        case AL_RMIN: ok = run_reduce<AL_RMIN>(src->ptr, out->ptr, K, R, init); break;
        case AL_RPROD: ok = run_reduce<AL_RPROD>(src->ptr, out->ptr, K, R, init); break;
    }
    if (ok && (redop == AL_RMAX || redop == AL_RMIN) && !rt().reduce_sequential && !rt().skip_zero_sign_fix)
        ok = fix_zero_sign(src, out->ptr, red, init, custom_init);
    if (!ok) {
        std::string keep = last_error_string();
        release_array(out);
        restore_error(keep);
        return nullptr;
    }
    return out;
}

// Local stage of a cross-GPU reduction: the same planning as reduce_impl, but the final stores of the
// output-contiguous kernel go straight into the owners' peer slots, un-finalised (peer.cuh).
bool reduce_into_peers(int redop, const NDArray* src, const size_t* dims, size_t ndims, const PeerMap& map, bool* used) {
    *used = false;
    if (!valid(src, "al_comm_reduce")) return false;
    bool red[64];
    size_t dshape[64];
    AxisList K, R;
    if (!build_axes(redop, src, dims, ndims, K, R, dshape, red)) return false;
    if (K.count() >= (1ull << 31) || R.count() >= (1ull << 31)) return true;  // caller takes the two-step route
    const float ident[AL_REDOP_COUNT] = {0.0f, -INFINITY, INFINITY, 1.0f};
    switch (redop) {
        case AL_RSUM: return run_reduce<AL_RSUM>(src->ptr, nullptr, K, R, ident[redop], &map, used);
        case AL_RMAX: return run_reduce<AL_RMAX>(src->ptr, nullptr, K, R, ident[redop], &map, used);
        case AL_RMIN: return run_reduce<AL_RMIN>(src->ptr, nullptr, K, R, ident[redop], &map, used);
        default: return run_reduce<AL_RPROD>(src->ptr, nullptr, K, R, ident[redop], &map, used);
    }
}

}  // namespace al

using namespace al;

extern "C" {

NDArray* array_reduce_named(int redop, const NDArray* src, const size_t* dims, size_t ndims, f32 init) {
    AL_ENTER;
    return reduce_impl(redop, src, dims, ndims, init, true);
}
NDArray* array_reduce_sum(const NDArray* src, const size_t* dims, size_t ndims) {  // arraylib.c:758-760
    AL_ENTER;
    return reduce_impl(AL_RSUM, src, dims, ndims, 0.f, false);
}
NDArray* array_reduce_max(const NDArray* src, const size_t* dims, size_t ndims) {  // arraylib.c:762-764
    AL_ENTER;
    return reduce_impl(AL_RMAX, src, dims, ndims, 0.f, false);
}
NDArray* array_reduce_min(const NDArray* src, const size_t* dims, size_t ndims) {  // arraylib.c:766-768
    AL_ENTER;
    return reduce_impl(AL_RMIN, src, dims, ndims, 0.f, false);
}
NDArray* array_reduce(binop, const NDArray*, const size_t*, size_t, f32) {
    AL_ENTER;
    set_error("NotImplementedError: host callbacks cannot run on the device; use array_reduce_named with an al_redop entry");
    return nullptr;
}

}  // extern "C"
