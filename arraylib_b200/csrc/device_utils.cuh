// Device-side helpers shared by the kernel translation units.
#pragma once
#include <cstdint>

namespace al {

// Division by a runtime-constant divisor via one mul-hi and a shift (valid for numerators below
// 2^31). This is what replaces the reference's per-element 64-bit `%` and `/` in offset_indexer
// (arraylib/src/arraylib.c:116-123).
struct FastDiv {
    uint32_t d, mul, shr;
};
inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f{d, 0, 0};
    if (d <= 1) {
        f.d = 1;
        return f;
    }
    uint32_t k = 0;
    while ((1ull << k) < d) k++;
    uint32_t p = 31 + k;
    f.mul = (uint32_t)(((1ull << p) + d - 1) / d);
    f.shr = p - 32;
    return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) {
    return f.d == 1 ? n : (__umulhi(n, f.mul) >> f.shr);
}

}  // namespace al
