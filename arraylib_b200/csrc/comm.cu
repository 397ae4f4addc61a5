// Multi-GPU plumbing: one process per GPU, NCCL resolved at run time (dlopen) so that a
// single-GPU user needs no NCCL at all. The only collectives the array core needs are the
// partial-result allreduce when a reduction spans the sharded leading axis, and an allgather to
// materialise a sharded array.
#include <dlfcn.h>
#include <nccl.h>

#include "al_internal.h"

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
    AL_SYM(GetErrorString, "ncclGetErrorString")
#undef AL_SYM
    return true;
}

static bool nccl_ok(ncclResult_t r, const char* what) {
    if (r == ncclSuccess) return true;
    set_error("RuntimeError: NCCL failure %s in %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?", what);
    return false;
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
    if (!valid(a, "al_comm_allreduce")) return -1;
    if (!g_nccl.comm) return set_error("RuntimeError: al_comm_init has not been called"), -1;
    if (!is_contiguous(a)) return set_error("ValueError: allreduce needs a contiguous array"), -1;
    const ncclRedOp_t ops[AL_REDOP_COUNT] = {ncclSum, ncclMax, ncclMin, ncclProd};
    if (redop < 0 || redop >= AL_REDOP_COUNT) return set_error("ValueError: unknown reduce op %d", redop), -1;
    size_t n = count_of(a->lay->shape, a->lay->ndim);
    if (!nccl_ok(g_nccl.AllReduce(a->ptr, a->ptr, n, ncclFloat32, ops[redop], g_nccl.comm, rt().stream), "ncclAllReduce"))
        return -1;
    note_launch("nccl_allreduce");
    return 0;
}

int al_comm_allgather(const NDArray* shard, NDArray* full) {
    AL_ENTER;
    if (!valid(shard, "al_comm_allgather") || !valid(full, "al_comm_allgather")) return -1;
    if (!g_nccl.comm) return set_error("RuntimeError: al_comm_init has not been called"), -1;
    if (!is_contiguous(shard) || !is_contiguous(full)) return set_error("ValueError: allgather needs contiguous arrays"), -1;
    size_t n = count_of(shard->lay->shape, shard->lay->ndim);
    if (count_of(full->lay->shape, full->lay->ndim) != n * (size_t)g_nccl.world)
        return set_error("ValueError: allgather output must hold world * shard elements"), -1;
    if (!nccl_ok(g_nccl.AllGather(shard->ptr, full->ptr, n, ncclFloat32, g_nccl.comm, rt().stream), "ncclAllGather")) return -1;
    note_launch("nccl_allgather");
    return 0;
}

int al_comm_destroy(void) {
    AL_ENTER;
    if (g_nccl.comm) {
        cudaStreamSynchronize(rt().stream);
        g_nccl.CommDestroy(g_nccl.comm);
        g_nccl.comm = nullptr;
    }
    return 0;
}

}  // extern "C"
