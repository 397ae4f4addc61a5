// Compatibility-only kernels (out of the measured hot path): matmul and dot.
#include "al_internal.h"

namespace al {

// arraylib/src/arraylib.c:522-564 -- for a fixed (i, j) the k terms are accumulated in ascending
// order into an f32 starting at 0, product fused into the add like the reference build.
__global__ void __launch_bounds__(256)
matmul_kernel(const float* __restrict__ a, long long as0, long long as1, const float* __restrict__ b, long long bs0,
              long long bs1, float* __restrict__ out, uint32_t M, uint32_t K, uint32_t N) {
    const uint32_t j = blockIdx.x * 16 + threadIdx.x % 16, i = blockIdx.y * 16 + threadIdx.x / 16;
    if (i >= M || j >= N) return;
    float acc = 0.0f;
    for (uint32_t k = 0; k < K; k++) acc = fmaf(a[i * as0 + k * as1], b[k * bs0 + j * bs1], acc);
    out[(size_t)i * N + j] = acc;
}

}  // namespace al

using namespace al;

extern "C" {

NDArray* array_array_matmul(const NDArray* lhs, const NDArray* rhs) {
    AL_ENTER;
    if (!valid(lhs, "matmul") || !valid(rhs, "matmul")) return nullptr;
    AL_REQUIRE(lhs->lay->ndim == 2 && rhs->lay->ndim == 2, "ValueError: matmul needs 2-D operands");
    AL_REQUIRE(lhs->lay->shape[1] == rhs->lay->shape[0], "ValueError: matmul inner extents differ");
    size_t M = lhs->lay->shape[0], K = lhs->lay->shape[1], N = rhs->lay->shape[1];
    AL_REQUIRE(M < (1u << 20) && N < (1u << 20), "NotImplementedError: matmul extent too large for the compat kernel");
    size_t shape[2] = {M, N};
    NDArray* out = new_array(shape, 2);
    if (!out) return nullptr;
    dim3 grid((unsigned)((N + 15) / 16), (unsigned)((M + 15) / 16));
    matmul_kernel<<<grid, 256, 0, rt().stream>>>(lhs->ptr, lhs->lay->stride[0], lhs->lay->stride[1], rhs->ptr,
                                                 rhs->lay->stride[0], rhs->lay->stride[1], out->ptr, (uint32_t)M,
                                                 (uint32_t)K, (uint32_t)N);
    note_launch("matmul_compat");
    if (!check_launch("matmul")) {
        release_array(out);
        return nullptr;
    }
    return out;
}

// arraylib/src/arraylib.c:671-688 -- 1-D dot -> shape [1]; built from the hot-path kernels
// (elementwise multiply, then a full reduce_sum).
NDArray* array_array_dot(const NDArray* lhs, const NDArray* rhs) {
    AL_ENTER;
    if (!valid(lhs, "dot") || !valid(rhs, "dot")) return nullptr;
    AL_REQUIRE(lhs->lay->ndim == 1, "ValueError: lhs dimension != 1 for dot.");
    AL_REQUIRE(rhs->lay->ndim == 1, "ValueError: rhs dimension != 1 for dot.");
    AL_REQUIRE(lhs->lay->shape[0] == rhs->lay->shape[0], "ValueError: dot shape mismatch.");
    NDArray* prod = array_array_mul(lhs, rhs);
    if (!prod) return nullptr;
    size_t dims[1] = {0};
    NDArray* out = array_reduce_sum(prod, dims, 1);
    std::string keep = last_error_string();
    release_array(prod);
    restore_error(keep);
    return out;
}

}  // extern "C"
