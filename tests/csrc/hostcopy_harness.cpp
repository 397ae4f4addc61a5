// Checks arraylib_b200/csrc/hostcopy.cpp (streaming-store copy of the pinned staging ring) against memcpy on random sizes and
// alignments, including the bytes around the destination. Built and run by tests/test_host_logic.py (CPU only).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
extern "C" void al_host_copy_stream(char* dst, const char* src, size_t n);
int main() {
    std::vector<char> a(40 << 20), b(40 << 20), c(40 << 20);
    for (size_t i = 0; i < a.size(); i++) a[i] = (char)(i * 2654435761u >> 13);
    srand(1);
    for (int t = 0; t < 200; t++) {
        size_t so = rand() % 97, dofs = rand() % 101, n = (t % 3 == 0) ? rand() % 4096 : (size_t)rand() % (33u << 20);
        memset(b.data(), 0, n + 256); memset(c.data(), 0, n + 256);
        al_host_copy_stream(b.data() + dofs, a.data() + so, n);
        memcpy(c.data() + dofs, a.data() + so, n);
        if (memcmp(b.data(), c.data(), n + 256)) { printf("MISMATCH t=%d n=%zu\n", t, n); return 1; }
    }
    printf("hostcopy ok\n");
    return 0;
}
