// Host-side memcpy for the pinned staging ring (runtime.cu): non-temporal stores.
// glibc's memcpy only switches to streaming stores above ~3/4 of the shared cache size; the ring hands every helper
// thread a 3-4 MB piece, so each destination line was first READ into the cache (write-allocate) and then written back:
// three bytes of memory traffic per byte copied, on a path (pageable from_numpy / to_numpy) that is bound by the host's
// memory bandwidth. Streaming stores make it two. Plain C++ (built with g++ -mavx2, no CUDA in this file).
#include <immintrin.h>

#include <cstddef>
#include <cstdint>
#include <cstring>

extern "C" __attribute__((visibility("hidden"))) void al_host_copy_stream(char* dst, const char* src, size_t n) {
    if (n < ((size_t)256 << 10) || !__builtin_cpu_supports("avx2")) {
        memcpy(dst, src, n);
        return;
    }
    size_t head = (32 - ((uintptr_t)dst & 31)) & 31;
    memcpy(dst, src, head);
    dst += head, src += head, n -= head;
    const size_t body = n & ~(size_t)127;
    for (size_t i = 0; i < body; i += 128) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 64));
        const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), a);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 96), d);
    }
    _mm_sfence();
    memcpy(dst + body, src + body, n - body);
}
