// Multi-GPU plumbing: one process per GPU, NCCL resolved at run time (dlopen) so that a
// single-GPU user needs no NCCL at all. The only collectives the array core needs are the
// partial-result allreduce when a reduction spans the sharded leading axis, and an allgather to
// materialise a sharded array.
#include <dlfcn.h>
#include <nccl.h>

#include "al_internal.h"
#include "peer.cuh"

namespace al {

struct Nccl {
    void* handle = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static Nccl g_nccl;

static bool load_nccl() {
    if (g_nccl.handle) return true;
    const char* names[] = {getenv("ARRAYLIB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n) continue;
        g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) {
        set_error("RuntimeError: cannot load libnccl.so.2 (%s)", dlerror());
        return false;
    }
#define AL_SYM(field, sym)                                                        \
    g_nccl.field = (decltype(g_nccl.field))dlsym(g_nccl.handle, sym);             \
    if (!g_nccl.field) {                                                          \
        set_error("RuntimeError: NCCL symbol %s not found", sym);                 \
        return false;                                                             \
    }
    AL_SYM(GetUniqueId, "ncclGetUniqueId")
    AL_SYM(CommInitRank, "ncclCommInitRank")
    AL_SYM(CommDestroy, "ncclCommDestroy")
    AL_SYM(AllReduce, "ncclAllReduce")
    AL_SYM(AllGather, "ncclAllGather")
    AL_SYM(ReduceScatter, "ncclReduceScatter")
    AL_SYM(Send, "ncclSend")
    AL_SYM(Recv, "ncclRecv")
    AL_SYM(GroupStart, "ncclGroupStart")
    AL_SYM(GroupEnd, "ncclGroupEnd")
    AL_SYM(GetErrorString, "ncclGetErrorString")
#undef AL_SYM
    return true;
}

static bool nccl_ok(ncclResult_t r, const char* what) {
    if (r == ncclSuccess) return true;
    set_error("RuntimeError: NCCL failure %s in %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?", what);
    return false;
}


// ---- peer-memory windows (see peer.cuh for the protocol) -------------------------------------------------
constexpr size_t kFlagBytes = 4096;            // 2 x [kMaxPeers] u32 arrival counters, padded
constexpr long long kPeerTimeoutNs = 180ll * 1000 * 1000 * 1000;  // a rank may legitimately lag (host work between collectives)

struct PeerWindows {
    bool tried = false, ok = false;
    std::string why = "not initialised";
    size_t cap_elems = 0;                       // capacity of `slots` and of `result`, in floats
    char* base[kMaxPeers] = {nullptr};          // base[rank] = own allocation, others = IPC mappings
    uint32_t epoch = 0;
    int parity = 0;                             // slot buffer of the collective in flight (epoch parity)
    int* host_error = nullptr;                  // mapped pinned int the kernels raise on a timeout
    int* dev_error = nullptr;
    uint32_t seg = 0;                           // segment size of the operation in flight
};
static PeerWindows g_peer;

static uint32_t* win_flags(int r) { return reinterpret_cast<uint32_t*>(g_peer.base[r]); }
// Two slot buffers, used alternately: a rank that has finished collective e may already push its partials of
// collective e+1 while a slower rank still combines the slots of e (the scatter variant has no second
// hand-over that would hold it back). Buffer e%2 is reused by e+2, and nobody can be in e+2 before every rank
// has announced e+1, i.e. has finished combining e.
static float* win_slots(int r, int parity) { return reinterpret_cast<float*>(g_peer.base[r] + kFlagBytes) + (size_t)parity * g_peer.cap_elems; }
static float* win_result(int r) { return reinterpret_cast<float*>(g_peer.base[r] + kFlagBytes) + 2 * g_peer.cap_elems; }

static void close_windows() {
    for (int r = 0; r < g_nccl.world && r < kMaxPeers; r++) {
        if (!g_peer.base[r]) continue;
        if (r == g_nccl.rank) cudaFree(g_peer.base[r]);
        else cudaIpcCloseMemHandle(g_peer.base[r]);
        g_peer.base[r] = nullptr;
    }
    g_peer.cap_elems = 0;
    g_peer.ok = false;
}

// all ranks agree (min over ranks) on whether a local step worked
static bool all_ranks_ok(bool mine) {
    float* flag = (float*)device_alloc(sizeof(float));
    if (!flag) return false;
    float h = mine ? 1.0f : 0.0f;
    bool ok = cudaMemcpyAsync(flag, &h, sizeof h, cudaMemcpyHostToDevice, rt().stream) == cudaSuccess &&
              g_nccl.AllReduce(flag, flag, 1, ncclFloat32, ncclMin, g_nccl.comm, rt().stream) == ncclSuccess &&
              cudaMemcpyAsync(&h, flag, sizeof h, cudaMemcpyDeviceToHost, rt().stream) == cudaSuccess &&
              cudaStreamSynchronize(rt().stream) == cudaSuccess;
    device_free(flag);
    return ok && h == 1.0f;
}

// (Re)create the windows with room for `elems` floats in slots and in result. Collective: every rank
// calls it with the same size at the same point of its stream.
static bool open_windows(size_t elems) {
    const int world = g_nccl.world, rank = g_nccl.rank;
    g_peer.tried = true;
    if (world > kMaxPeers) {
        g_peer.why = "more than 8 ranks";
        return false;
    }
    // quiesce: nobody may still be using the old windows
    if (!all_ranks_ok(true)) {
        g_peer.why = "ranks did not agree before remapping";
        return false;
    }
    close_windows();
    cudaGetLastError();
    const size_t bytes = kFlagBytes + 3 * elems * sizeof(float);
    bool mine = true;
    char* own = nullptr;
    cudaIpcMemHandle_t handle;
    memset(&handle, 0, sizeof handle);
    if (cudaMalloc(&own, bytes) != cudaSuccess || cudaMemsetAsync(own, 0, kFlagBytes, rt().stream) != cudaSuccess ||
        cudaIpcGetMemHandle(&handle, own) != cudaSuccess) {
        mine = false;
        g_peer.why = std::string("cudaMalloc / cudaIpcGetMemHandle: ") + cudaGetErrorString(cudaGetLastError());
    }
    if (!g_peer.host_error) {
        if (cudaHostAlloc(&g_peer.host_error, sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
            cudaHostGetDevicePointer(&g_peer.dev_error, g_peer.host_error, 0) != cudaSuccess) {
            mine = false;
            g_peer.why = "cudaHostAlloc (error flag)";
            g_peer.host_error = nullptr;
        } else {
            *g_peer.host_error = 0;
        }
    }
    // exchange the handles: one in-place NCCL allgather of world x 64 bytes
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::vector<cudaIpcMemHandle_t> all(world);
    char* xchg = (char*)device_alloc(64 * (size_t)world);
    bool sent = xchg && cudaMemcpyAsync(xchg + 64 * rank, &handle, 64, cudaMemcpyHostToDevice, rt().stream) == cudaSuccess &&
                g_nccl.AllGather(xchg + 64 * rank, xchg, 64, ncclChar, g_nccl.comm, rt().stream) == ncclSuccess &&
                cudaMemcpyAsync(all.data(), xchg, 64 * (size_t)world, cudaMemcpyDeviceToHost, rt().stream) == cudaSuccess &&
                cudaStreamSynchronize(rt().stream) == cudaSuccess;
    if (xchg) device_free(xchg);
    if (!sent) {
        mine = false;
        g_peer.why = "handle exchange failed";
    }
    if (!all_ranks_ok(mine)) {  // somebody could not even allocate: nobody maps anything
        if (own) cudaFree(own);
        if (mine) g_peer.why = "another rank could not create its window";
        cudaGetLastError();
        return false;
    }
    g_peer.base[rank] = own;
    for (int r = 0; r < world && mine; r++) {
        if (r == rank) continue;
        void* mapped = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&mapped, all[r], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            mine = false;
            g_peer.why = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e);
            cudaGetLastError();
        } else {
            g_peer.base[r] = (char*)mapped;
        }
    }
    g_peer.cap_elems = elems;
    if (!all_ranks_ok(mine)) {
        if (mine) g_peer.why = "another rank could not map the windows";
        close_windows();
        return false;
    }
    g_peer.ok = true;
    g_peer.why = "ok";
    return true;
}

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ long long global_ns() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

struct HandOver {
    uint32_t* announce[kMaxPeers];  // &flags_of_rank_r[phase * kMaxPeers + my_rank]
    const uint32_t* arrivals;       // &my_flags[phase * kMaxPeers]
    uint32_t epoch;
    int world;
    int* error;
};
// Block 0 announces to every rank; every block then waits until all `world` announcements for this
// epoch have landed in local memory. Bounded: after kPeerTimeoutNs the error flag is raised and the
// kernel carries on (the result is garbage, the host reports the failure).
__device__ __forceinline__ void hand_over(const HandOver& h) {
    if (blockIdx.x == 0 && threadIdx.x < h.world) st_release_sys(h.announce[threadIdx.x], h.epoch);
    if (threadIdx.x < h.world) {
        const long long t0 = global_ns();
        while ((int32_t)(ld_acquire_sys(h.arrivals + threadIdx.x) - h.epoch) < 0) {
            if (global_ns() - t0 > kPeerTimeoutNs) {
                *(volatile int*)h.error = 1;
                break;
            }
            __nanosleep(200);
        }
    }
    __syncthreads();
}

// step 1 for a partial that already sits in local memory: scatter it into the owners' slots
__global__ void __launch_bounds__(256) peer_push_kernel(const float* __restrict__ partial, uint32_t nout, const PeerMap map, uint32_t rot) {
    // `rot` (a multiple of 4) rotates the walk so that rank r starts with the segment of owner r: at any moment
    // the ranks write to DIFFERENT owners instead of all converging on one GPU's NVLink ingress
    const uint32_t stride = gridDim.x * blockDim.x * 4;
    const uint32_t padded = (nout + 3) & ~3u;
    for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) * 4; i < padded; i += stride) {
        uint32_t e = i + rot;
        if (e >= padded) e -= padded;
        float* d = peer_dst(map, e);
        if (e + 4 <= nout) {
            float4 v = make_float4(partial[e], partial[e + 1], partial[e + 2], partial[e + 3]);
            *reinterpret_cast<float4*>(d) = v;
        } else {
            for (uint32_t k = 0; e + k < nout; k++) d[k] = partial[e + k];
        }
    }
}

template <int OP>
__device__ __forceinline__ float comb(float a, float b) {
    if constexpr (OP == AL_RSUM) return __fadd_rn(a, b);
    // max / min like the reference's fold (gcc's inline fmax / fmin, arraylib.c:18-19): NaNs are ignored and a tie --
    // +0 against -0 -- keeps the LEFT operand. Partials are combined in rank order = logical order of the leading axis,
    // and every partial already carries the sign of ITS first zero (reduce_zero_sign_kernel), so the result has the sign
    // of the first zero of the whole reduced set, exactly as on one GPU.
    else if constexpr (OP == AL_RMAX) return (b != b) ? a : ((a != a) ? b : (a >= b ? a : b));
    else if constexpr (OP == AL_RMIN) return (b != b) ? a : ((a != a) ? b : (a <= b ? a : b));
    else return __fmul_rn(a, b);
}

struct CombineArgs {
    HandOver h;
    const float* slots;             // own window: [world][seg]
    float* result[kMaxPeers];       // &result_of_rank_r[my_rank * seg]
    uint32_t seg, count;            // count = elements of the own segment that exist (<= seg)
    float init;
    int nres;                       // destinations: world (allreduce: every rank's result window) or 1 (scatter: local output)
};
// step 2: own segment = init (+) (init (+) (p0 (+) p1 (+) ... in rank order)), stored into every rank's result
template <int OP>
__global__ void __launch_bounds__(256) peer_combine_kernel(const CombineArgs a) {
    hand_over(a.h);
    const int world = a.h.world;
    const uint32_t stride = gridDim.x * blockDim.x * 4;
    for (uint32_t e = (blockIdx.x * blockDim.x + threadIdx.x) * 4; e < a.count; e += stride) {
        float4 v = __ldcg(reinterpret_cast<const float4*>(a.slots + e));
        for (int r = 1; r < world; r++) {
            const float4 w = __ldcg(reinterpret_cast<const float4*>(a.slots + (size_t)r * a.seg + e));
            v.x = comb<OP>(v.x, w.x);
            v.y = comb<OP>(v.y, w.y);
            v.z = comb<OP>(v.z, w.z);
            v.w = comb<OP>(v.w, w.w);
        }
        v.x = comb<OP>(a.init, comb<OP>(a.init, v.x));
        v.y = comb<OP>(a.init, comb<OP>(a.init, v.y));
        v.z = comb<OP>(a.init, comb<OP>(a.init, v.z));
        v.w = comb<OP>(a.init, comb<OP>(a.init, v.w));
        if (e + 4 <= a.count) {
            for (int r = 0; r < a.nres; r++) *reinterpret_cast<float4*>(a.result[r] + e) = v;
        } else {  // ragged end of the last segment (only the scatter variant writes into an exactly sized array)
            const float t[4] = {v.x, v.y, v.z, v.w};
            for (int r = 0; r < a.nres; r++)
                for (uint32_t k = 0; e + k < a.count; k++) a.result[r][e + k] = t[k];
        }
    }
}

// step 3: after every owner delivered its segment, copy the finished result out of the window
__global__ void __launch_bounds__(256) peer_collect_kernel(const HandOver h, const float* __restrict__ result, float* __restrict__ out, uint32_t nout) {
    hand_over(h);
    const uint32_t stride = gridDim.x * blockDim.x * 4;
    const bool aligned = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    for (uint32_t e = (blockIdx.x * blockDim.x + threadIdx.x) * 4; e < nout; e += stride) {
        if (aligned && e + 4 <= nout) {
            *reinterpret_cast<float4*>(out + e) = __ldcg(reinterpret_cast<const float4*>(result + e));
        } else {
            for (uint32_t k = 0; k < 4 && e + k < nout; k++) out[e + k] = __ldcg(result + e + k);
        }
    }
}

static HandOver make_hand_over(int phase) {
    HandOver h;
    memset(&h, 0, sizeof h);
    for (int r = 0; r < g_nccl.world; r++) h.announce[r] = win_flags(r) + phase * kMaxPeers + g_nccl.rank;
    h.arrivals = win_flags(g_nccl.rank) + phase * kMaxPeers;
    h.epoch = g_peer.epoch;
    h.world = g_nccl.world;
    h.error = g_peer.dev_error;
    return h;
}

static bool peer_error_raised() {
    if (g_peer.host_error && *(volatile int*)g_peer.host_error) {
        set_error("RuntimeError: peer collective timed out waiting for another rank");
        return true;
    }
    return false;
}

bool comm_error_pending() { return peer_error_raised(); }

bool peer_available() {
    return g_nccl.comm && g_nccl.world > 1 && rt().peer_collective && (!g_peer.tried || g_peer.ok);
}

bool peer_begin(uint64_t nout, PeerMap* map) {
    if (!peer_available() || nout == 0 || nout >= (1ull << 31) - 64) return false;
    const int world = g_nccl.world;
    uint64_t seg = (nout + world - 1) / world;
    seg = (seg + 3) & ~3ull;
    const size_t need = (size_t)seg * world;
    if (need > g_peer.cap_elems) {
        size_t want = need < (16u << 20) ? (16u << 20) : need;  // at least 2 x 64 MiB: C4's 2^24 outputs fit without a remap
        if (!open_windows(want)) return false;
    }
    if (peer_error_raised()) return false;
    g_peer.seg = (uint32_t)seg;
    memset(map, 0, sizeof *map);
    g_peer.parity = (int)((g_peer.epoch + 1) & 1);
    for (int o = 0; o < world; o++) map->slot[o] = win_slots(o, g_peer.parity) + (size_t)g_nccl.rank * seg;
    map->seg = (uint32_t)seg;
    map->segdiv = make_fastdiv((uint32_t)seg);
    map->world = world;
    map->rank = g_nccl.rank;
    return true;
}

bool peer_push(const float* partial, uint64_t nout, const PeerMap& map) {
    unsigned blocks = (unsigned)((nout / 4 + 255) / 256 + 1);
    if (blocks > (unsigned)kNumSMs * 8) blocks = kNumSMs * 8;
    const uint32_t rot = (uint32_t)(((uint64_t)g_nccl.rank * map.seg) % ((nout + 3) & ~3ull));
    peer_push_kernel<<<blocks, 256, 0, rt().stream>>>(partial, (uint32_t)nout, map, rot);
    note_launch("peer_push");
    return check_launch("peer_push");
}

bool peer_finish(int redop, float* out, uint64_t nout, float init, bool scatter) {
    const int world = g_nccl.world, rank = g_nccl.rank;
    const uint32_t seg = g_peer.seg;
    g_peer.epoch++;
    CombineArgs c;
    memset(&c, 0, sizeof c);
    c.h = make_hand_over(0);
    c.slots = win_slots(rank, g_peer.parity);
    if (scatter) {  // the own segment is the whole local result: one hand-over, no broadcast, no collect
        c.nres = 1;
        c.result[0] = out;
    } else {
        c.nres = world;
        for (int r = 0; r < world; r++) c.result[r] = win_result(r) + (size_t)rank * seg;
    }
    c.seg = seg;
    const uint64_t lo = (uint64_t)rank * seg;
    c.count = lo >= nout ? 0u : (uint32_t)(nout - lo < seg ? nout - lo : seg);
    c.init = init;
    unsigned blocks = (unsigned)((c.count / 4 + 255) / 256 + 1);
    if (blocks > (unsigned)kNumSMs * 4) blocks = kNumSMs * 4;
    cudaStream_t st = rt().stream;
    switch (redop) {
        case AL_RSUM: peer_combine_kernel<AL_RSUM><<<blocks, 256, 0, st>>>(c); break;
        case AL_RMAX: peer_combine_kernel<AL_RMAX><<<blocks, 256, 0, st>>>(c); break;
        case AL_RMIN: peer_combine_kernel<AL_RMIN><<<blocks, 256, 0, st>>>(c); break;
        default: peer_combine_kernel<AL_RPROD><<<blocks, 256, 0, st>>>(c); break;
    }
    note_launch("peer_combine");
    if (!check_launch("peer_combine")) return false;
    if (scatter) return true;
    blocks = (unsigned)((nout / 4 + 255) / 256 + 1);
    if (blocks > (unsigned)kNumSMs * 8) blocks = kNumSMs * 8;
    peer_collect_kernel<<<blocks, 256, 0, st>>>(make_hand_over(1), win_result(rank), out, (uint32_t)nout);
    note_launch("peer_collect");
    return check_launch("peer_collect");
}

}  // namespace al

using namespace al;

extern "C" {

int al_comm_unique_id(char out[128]) {
    AL_ENTER;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    if (!ensure_init() || !load_nccl()) return -1;
    ncclUniqueId id;
    if (!nccl_ok(g_nccl.GetUniqueId(&id), "ncclGetUniqueId")) return -1;
    memcpy(out, &id, 128);
    return 0;
}

int al_comm_init(const char id_bytes[128], int rank, int world) {
    AL_ENTER;
    if (!ensure_init() || !load_nccl()) return -1;
    if (g_nccl.comm) return 0;
    ncclUniqueId id;
    memcpy(&id, id_bytes, 128);
    if (!nccl_ok(g_nccl.CommInitRank(&g_nccl.comm, world, id, rank), "ncclCommInitRank")) return -1;
    g_nccl.rank = rank;
    g_nccl.world = world;
    return 0;
}

int al_comm_allreduce(NDArray* a, int redop) {
    AL_ENTER;
    if (!valid_mut(a, "al_comm_allreduce")) return -1;
    if (!g_nccl.comm) return set_error("RuntimeError: al_comm_init has not been called"), -1;
    if (!is_contiguous(a)) return set_error("ValueError: allreduce needs a contiguous array"), -1;
    const ncclRedOp_t ops[AL_REDOP_COUNT] = {ncclSum, ncclMax, ncclMin, ncclProd};
    if (redop < 0 || redop >= AL_REDOP_COUNT) return set_error("ValueError: unknown reduce op %d", redop), -1;
    if (peer_error_raised()) return -1;
    size_t n = count_of(a->lay->shape, a->lay->ndim);
    PeerMap map;
    // Small messages are barriers and host scalars (dist.barrier, allreduce_scalar): ranks may legitimately arrive
    // minutes apart there (one of them is still on the host), which the bounded wait of the peer kernels would report
    // as a failure. They go through NCCL, which simply waits.
    if (n * sizeof(float) >= (64u << 10) && peer_begin(n, &map)) {  // same windows as al_comm_reduce: push, rank-ordered combine, collect (in place)
        const float ident[AL_REDOP_COUNT] = {0.0f, -INFINITY, INFINITY, 1.0f};
        return peer_push(a->ptr, n, map) && peer_finish(redop, a->ptr, n, ident[redop], false) ? 0 : -1;
    }
    clear_error();
    if (!nccl_ok(g_nccl.AllReduce(a->ptr, a->ptr, n, ncclFloat32, ops[redop], g_nccl.comm, rt().stream), "ncclAllReduce"))
        return -1;
    note_launch("nccl_allreduce");
    return 0;
}

int al_comm_allgather(const NDArray* shard, NDArray* full) {
    AL_ENTER;
    if (!valid(shard, "al_comm_allgather") || !valid_mut(full, "al_comm_allgather")) return -1;
    if (!g_nccl.comm) return set_error("RuntimeError: al_comm_init has not been called"), -1;
    if (!is_contiguous(shard) || !is_contiguous(full)) return set_error("ValueError: allgather needs contiguous arrays"), -1;
    size_t n = count_of(shard->lay->shape, shard->lay->ndim);
    if (count_of(full->lay->shape, full->lay->ndim) != n * (size_t)g_nccl.world)
        return set_error("ValueError: allgather output must hold world * shard elements"), -1;
    if (!nccl_ok(g_nccl.AllGather(shard->ptr, full->ptr, n, ncclFloat32, g_nccl.comm, rt().stream), "ncclAllGather")) return -1;
    note_launch("nccl_allgather");
    return 0;
}

int al_comm_alltoall(const NDArray* send, NDArray* recv) {
    AL_ENTER;
    if (!valid(send, "al_comm_alltoall") || !valid_mut(recv, "al_comm_alltoall")) return -1;
    if (!g_nccl.comm) return set_error("RuntimeError: al_comm_init has not been called"), -1;
    if (!is_contiguous(send) || !is_contiguous(recv)) return set_error("ValueError: alltoall needs contiguous arrays"), -1;
    const size_t n = count_of(send->lay->shape, send->lay->ndim);
    const size_t world = (size_t)g_nccl.world;
    if (count_of(recv->lay->shape, recv->lay->ndim) != n || n % world != 0)
        return set_error("ValueError: alltoall needs equal-sized arrays whose element count is a multiple of the world size"), -1;
    const size_t chunk = n / world;
    // block r of `send` goes to rank r; block r of `recv` comes from rank r (grouped point-to-point, one NCCL launch)
    bool ok = nccl_ok(g_nccl.GroupStart(), "ncclGroupStart");
    for (size_t r = 0; ok && r < world; r++) {
        ok = nccl_ok(g_nccl.Send(send->ptr + r * chunk, chunk, ncclFloat32, (int)r, g_nccl.comm, rt().stream), "ncclSend") &&
             nccl_ok(g_nccl.Recv(recv->ptr + r * chunk, chunk, ncclFloat32, (int)r, g_nccl.comm, rt().stream), "ncclRecv");
    }
    if (!nccl_ok(g_nccl.GroupEnd(), "ncclGroupEnd") || !ok) return -1;
    note_launch("nccl_alltoall");
    return 0;
}

// Where the flat output of a cross-GPU reduction is cut when it stays SHARDED: along its first axis with an
// extent > 1 (the leading non-reduced axis), in `world` equal blocks that are also the flat peer segments.
static bool scatter_axis(const size_t* dshape, size_t nd, int world, size_t* axis) {
    for (size_t i = 0; i < nd; i++) {
        if (dshape[i] == 1) continue;
        const uint64_t nout = count_of(dshape, nd);
        if (dshape[i] % (size_t)world != 0 || (nout / world) % 4 != 0) return false;
        *axis = i;
        return true;
    }
    return false;  // a full reduce has a single output: nothing to shard
}

static NDArray* comm_reduce(int redop, const NDArray* shard, const size_t* dims, size_t ndims, size_t shard_ax, bool scatter) {
    if (!valid(shard, "al_comm_reduce")) return nullptr;
    AL_REQUIRE(redop >= 0 && redop < AL_REDOP_COUNT, "ValueError: unknown reduce op %d", redop);
    const size_t nd = shard->lay->ndim;
    AL_REQUIRE(nd >= 1 && nd <= 64, "ValueError: al_comm_reduce needs 1..64 dimensions");
    AL_REQUIRE(shard_ax < nd, "ValueError: shard axis out of bounds");
    for (size_t i = 0; i < shard_ax; i++)
        AL_REQUIRE(shard->lay->shape[i] == 1, "ValueError: the shard axis must be the leading non-unit axis");
    bool crosses = false;
    size_t dshape[64];
    for (size_t i = 0; i < nd; i++) dshape[i] = shard->lay->shape[i];
    for (size_t i = 0; i < ndims; i++) {
        AL_REQUIRE(dims[i] < nd, "ValueError: out of bounds");
        dshape[dims[i]] = 1;
        crosses |= dims[i] == shard_ax;
    }
    if (!crosses || !g_nccl.comm || g_nccl.world == 1) {
        AL_REQUIRE(!scatter, "ValueError: al_comm_reduce_scatter needs a communicator and a reduction over axis 0");
        return reduce_impl(redop, shard, dims, ndims, 0.0f, false);
    }
    if (peer_error_raised()) return nullptr;
    const float ident[AL_REDOP_COUNT] = {0.0f, -INFINITY, INFINITY, 1.0f};
    const uint64_t nout = count_of(dshape, nd);
    size_t lshape[64], ax = 0;
    for (size_t i = 0; i < nd; i++) lshape[i] = dshape[i];
    if (scatter) {
        AL_REQUIRE(scatter_axis(dshape, nd, g_nccl.world, &ax),
                   "ValueError: the reduced shape cannot be cut into %d equal 16-byte aligned blocks along its leading axis", g_nccl.world);
        lshape[ax] = dshape[ax] / (size_t)g_nccl.world;
    }
    PeerMap map;
    if (peer_begin(nout, &map)) {
        // local kernel -> owners' slots over NVLink -> rank-ordered combine -> result on every rank (or, scatter: the
        // own segment only)
        // (max / min go through the ordinary local reduce: its zero-sign repair pass must run before the partial leaves)
        bool fused = false;
        if (redop != AL_RMAX && redop != AL_RMIN && !reduce_into_peers(redop, shard, dims, ndims, map, &fused)) return nullptr;
        if (!fused) {
            NDArray* part = reduce_impl(redop, shard, dims, ndims, 0.0f, false);
            if (!part) return nullptr;
            bool ok = peer_push(part->ptr, nout, map);
            std::string keep = last_error_string();
            release_array(part);
            restore_error(keep);
            if (!ok) return nullptr;
        }
        NDArray* out = new_array(scatter ? lshape : dshape, nd);
        if (!out) return nullptr;
        if (!peer_finish(redop, out->ptr, nout, ident[redop], scatter)) {
            std::string keep = last_error_string();
            release_array(out);
            restore_error(keep);
            return nullptr;
        }
        return out;
    }
    clear_error();
    NDArray* part = reduce_impl(redop, shard, dims, ndims, 0.0f, false);
    if (!part) return nullptr;
    const ncclRedOp_t ops[AL_REDOP_COUNT] = {ncclSum, ncclMax, ncclMin, ncclProd};
    if (scatter) {
        NDArray* out = new_array(lshape, nd);
        bool ok = out && nccl_ok(g_nccl.ReduceScatter(part->ptr, out->ptr, nout / g_nccl.world, ncclFloat32, ops[redop], g_nccl.comm, rt().stream), "ncclReduceScatter");
        std::string keep = last_error_string();
        release_array(part);
        if (!ok && out) release_array(out);
        restore_error(keep);
        if (!ok) return nullptr;
        note_launch("nccl_reduce_scatter");
        return out;
    }
    if (!nccl_ok(g_nccl.AllReduce(part->ptr, part->ptr, nout, ncclFloat32, ops[redop], g_nccl.comm, rt().stream), "ncclAllReduce")) {
        std::string keep = last_error_string();
        release_array(part);
        restore_error(keep);
        return nullptr;
    }
    note_launch("nccl_allreduce");
    return part;
}

// Result REPLICATED on every rank (allreduce semantics).
NDArray* al_comm_reduce(int redop, const NDArray* shard, const size_t* dims, size_t ndims) {
    AL_ENTER;
    return comm_reduce(redop, shard, dims, ndims, 0, false);
}

// Result SHARDED along its leading non-reduced axis (reduce-scatter semantics): rank r gets block r.
NDArray* al_comm_reduce_scatter(int redop, const NDArray* shard, const size_t* dims, size_t ndims) {
    AL_ENTER;
    return comm_reduce(redop, shard, dims, ndims, 0, true);
}

// Both of the above for a shard that is cut along `shard_axis` (all axes before it have extent 1).
NDArray* al_comm_reduce_ex(int redop, const NDArray* shard, const size_t* dims, size_t ndims, size_t shard_axis, int scatter) {
    AL_ENTER;
    return comm_reduce(redop, shard, dims, ndims, shard_axis, scatter != 0);
}

// 1 when al_comm_reduce_scatter can shard the result of reducing `dims` of an array of this shape.
int al_comm_can_scatter(const size_t* shape, size_t nd, const size_t* dims, size_t ndims) {
    AL_ENTER;
    if (!g_nccl.comm || g_nccl.world == 1 || nd == 0 || nd > 64) return 0;
    size_t dshape[64], ax;
    bool crosses = false;
    for (size_t i = 0; i < nd; i++) dshape[i] = shape[i];
    for (size_t i = 0; i < ndims; i++) {
        if (dims[i] >= nd) return 0;
        dshape[dims[i]] = 1;
        crosses |= dims[i] == 0;
    }
    return crosses && scatter_axis(dshape, nd, g_nccl.world, &ax) ? 1 : 0;
}

const char* al_comm_transport(void) {
    AL_ENTER;
    static thread_local std::string text;
    if (!g_nccl.comm || g_nccl.world == 1) text = "single process";
    else if (!rt().peer_collective) text = "nccl (peer_collective=0)";
    else if (!g_peer.tried) text = "peer-memory (windows not mapped yet)";
    else if (g_peer.ok) text = "peer-memory";
    else text = "nccl (" + g_peer.why + ")";
    return text.c_str();
}

int al_comm_destroy(void) {
    AL_ENTER;
    if (g_nccl.comm) {
        cudaStreamSynchronize(rt().stream);
        if (g_peer.cap_elems) {
            all_ranks_ok(true);  // nobody is still writing into a window that is about to go away
            close_windows();
        }
        g_peer.tried = false;
        g_nccl.CommDestroy(g_nccl.comm);
        g_nccl.comm = nullptr;
    }
    return 0;
}

}  // extern "C"
