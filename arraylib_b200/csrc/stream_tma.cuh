// EXPERIMENT (off by default, al_set_tuning("stream_tma", 1)): the contiguous binary stream driven
// entirely by the bulk-async engine: cp.async.bulk global->shared with mbarrier completion, compute
// from shared memory, cp.async.bulk shared->global. Kept as the measured answer to "would TMA-style
// staging beat plain 128-bit loads on this path?" -- see DESIGN.md (it does not: both sit at the DRAM
// ceiling, which is why the default kernels use LDG.128/STG.128).
// Caution: thread 0 issues the bulk copies right after __syncthreads(); a later TMA experiment on the window
// kernel found that BAR.SYNC.DEFER_BLOCKING does not hold back a following bulk-async
// instruction. No mismatch has been observed with this kernel, but it is opt-in for that reason too.
#pragma once
#include "elementwise.cuh"

namespace al {

constexpr int kTmaChunk = 2048;  // floats per operand per stage (8 KiB)
constexpr int kTmaStages = 3;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}

template <typename F>
__global__ void __launch_bounds__(256) ew_stream_tma_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                             float* __restrict__ out, size_t n, const F f) {
    extern __shared__ __align__(128) unsigned char tma_smem[];
    float* sa = reinterpret_cast<float*>(tma_smem);
    float* sb = sa + kTmaStages * kTmaChunk;
    float* so = sb + kTmaStages * kTmaChunk;
    uint64_t* bar = reinterpret_cast<uint64_t*>(so + kTmaStages * kTmaChunk);
    const size_t n_chunks = (n + kTmaChunk - 1) / kTmaChunk;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < kTmaStages; s++) mbar_init(bar + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](size_t it) {  // thread 0 only: start the loads of this CTA's it-th chunk
        const size_t chunk = blockIdx.x + it * gridDim.x;
        if (chunk >= n_chunks) return;
        const int s = (int)(it % kTmaStages);
        const size_t e0 = chunk * kTmaChunk;
        const uint32_t bytes = (uint32_t)(min((size_t)kTmaChunk, n - e0) * sizeof(float));
        mbar_expect_tx(bar + s, 2 * bytes);
        bulk_g2s(sa + s * kTmaChunk, a + e0, bytes, bar + s);
        bulk_g2s(sb + s * kTmaChunk, b + e0, bytes, bar + s);
    };
    if (tid == 0)
        for (int it = 0; it < kTmaStages - 1; it++) issue(it);
    for (size_t it = 0;; it++) {
        const size_t chunk = blockIdx.x + it * gridDim.x;
        if (chunk >= n_chunks) break;
        const int s = (int)(it % kTmaStages);
        // the output buffer of this stage was handed to a bulk store kTmaStages iterations ago
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kTmaStages - 1) : "memory");
        __syncthreads();  // also: everyone is done reading the stage that the next load will overwrite
        if (tid == 0) issue(it + kTmaStages - 1);
        mbar_wait(bar + s, (uint32_t)((it / kTmaStages) & 1));
        const size_t e0 = chunk * kTmaChunk;
        const uint32_t len4 = (uint32_t)(min((size_t)kTmaChunk, n - e0) / 4);
        const float4* pa = reinterpret_cast<const float4*>(sa + s * kTmaChunk);
        const float4* pb = reinterpret_cast<const float4*>(sb + s * kTmaChunk);
        float4* po = reinterpret_cast<float4*>(so + s * kTmaChunk);
        for (uint32_t i = tid; i < len4; i += 256) {
            const float4 x = pa[i], y = pb[i];
            float4 r;
            float e[2];
            e[0] = x.x, e[1] = y.x;
            r.x = f(e);
            e[0] = x.y, e[1] = y.y;
            r.y = f(e);
            e[0] = x.z, e[1] = y.z;
            r.z = f(e);
            e[0] = x.w, e[1] = y.w;
            r.w = f(e);
            po[i] = r;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the bulk store
        __syncthreads();
        if (tid == 0) {
            bulk_s2g(out + e0, po, len4 * 16);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// Returns 0 = not applicable, 1 = launched, -1 = error.
template <typename F>
static int try_launch_stream_tma(const F& f, const Folded& fo, const float* const* in, float* out) {
    if constexpr (F::NIN != 2) {
        return 0;
    } else {
        if (!rt().stream_tma || fo.ndim != 1 || fo.in_stride[0][0] != 1 || fo.in_stride[1][0] != 1) return 0;
        const size_t n = fo.extent[0];
        if (n % 4 != 0 || n < (size_t)kTmaChunk * kNumSMs) return 0;
        if ((uintptr_t)in[0] % 16 || (uintptr_t)in[1] % 16 || (uintptr_t)out % 16) return 0;
        const size_t smem = (size_t)3 * kTmaStages * kTmaChunk * sizeof(float) + kTmaStages * sizeof(uint64_t);
        static bool configured = false;
        if (!configured) {
            if (!cuda_ok(cudaFuncSetAttribute(ew_stream_tma_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                         "cudaFuncSetAttribute"))
                return -1;
            configured = true;
        }
        ew_stream_tma_kernel<F><<<kNumSMs * 3, 256, smem, rt().stream>>>(in[0], in[1], out, n, f);
        note_launch("ew_stream_tma");
        return check_launch("ew_stream_tma") ? 1 : -1;
    }
}

}  // namespace al
