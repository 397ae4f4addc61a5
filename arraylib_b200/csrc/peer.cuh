// Peer-memory collective for reductions that cross the sharded leading axis ("fuse the
// final cross-GPU add into the partial-reduce epilogue ... peer loads over NVLink from the partials in
// fixed order").
//
// Every rank owns one window of device memory that all other ranks of the node map through CUDA IPC
// (NVLink 5 / NVSwitch peer access). The window holds
//     slots  [world][seg]   slot r of rank o = rank r's partial of the output segment rank o owns
//     result [world * seg]  the finished output, every segment written by its owner
//     flags  2 x [world]    arrival counters (monotonic epochs) for the two hand-over points
// and one reduction is three launches per rank, no host synchronisation, no NCCL call:
//     1. the LOCAL reduce kernel stores its (un-finalised) outputs straight into the owners' slots
//        (PeerMap below; st.global over NVLink, overlapped with its own HBM reads),
//     2. peer_combine: announce "my partials are delivered", wait for every rank's announcement,
//        add the world partials of the own segment in RANK ORDER (bitwise identical on every rank
//        and run to run), apply the init fold and store the segment into every rank's result,
//     3. peer_collect: announce "my segment is delivered", wait for all, copy result -> output.
// Kernel boundaries order the data stores before the announcements (a launch starts after the
// previous launch's stores are performed at system scope); announcements are st.release.sys and the
// waiters poll their OWN memory with ld.acquire.sys, bounded by a wall-clock timeout that raises an
// error flag instead of hanging the GPU.
#pragma once
#include "al_internal.h"
#include "device_utils.cuh"

namespace al {

constexpr int kMaxPeers = 8;

// Where the local reduce kernel puts output element e: owner o = e / seg, this rank's slot there.
struct PeerMap {
    float* slot[kMaxPeers];
    FastDiv segdiv;
    uint32_t seg;  // elements per segment, a multiple of 4 (a float4 never straddles two owners)
    int world;     // 0: no peer target, store locally
    int rank;      // this rank (the fused kernels start their walk at their own segment, see reduce_outer)
};
__device__ __forceinline__ float* peer_dst(const PeerMap& m, uint32_t e) {
    const uint32_t o = fdiv(e, m.segdiv);
    return m.slot[o] + (e - o * m.seg);
}

// host side (comm.cu)
bool peer_available();                           // communicator up and the IPC windows mapped
bool peer_begin(uint64_t nout, PeerMap* map);    // grow the windows if needed (collective) and fill `map`
bool peer_push(const float* partial, uint64_t nout, const PeerMap& map);  // step 1 for an already materialised partial
bool peer_finish(int redop, float* out, uint64_t nout, float init, bool scatter);  // steps 2 and 3 (scatter: step 2 only, own segment -> out)
// local reduce with the peer epilogue (reduce.cu); *used = false when the chosen plan has none
bool reduce_into_peers(int redop, const NDArray* src, const size_t* dims, size_t ndims, const PeerMap& map, bool* used);

}  // namespace al
