// Compares every result that arraylib_b200/csrc/fast_ufunc.h ACCEPTS with glibc's (float)f((double)x), over all 2^32
// float bit patterns (stride 1) or a strided subset. Usage: fast_ufunc_harness [stride] [first] ; exit code 1 on a mismatch.
// Build twice: plain, and with -DFU_PERTURB_SEEDS (reciprocal / rsqrt seeds wobbled by +-2^-21 like a hardware MUFU).
#include <stdio.h>
#include <stdlib.h>
#include "../../arraylib_b200/csrc/fast_ufunc.h"

typedef int (*fast_fn)(float, float*);
typedef double (*ref_fn)(double);
static const struct { const char* name; fast_fn fast; ref_fn ref; } table[] = {
    {"exp", fu_fast_exp, exp},       {"log", fu_fast_log, log},       {"sin", fu_fast_sin, sin},
    {"cos", fu_fast_cos, cos},       {"tan", fu_fast_tan, tan},       {"sinh", fu_fast_sinh, sinh},
    {"cosh", fu_fast_cosh, cosh},    {"tanh", fu_fast_tanh, tanh},    {"asinh", fu_fast_asinh, asinh},
    {"acosh", fu_fast_acosh, acosh}, {"atanh", fu_fast_atanh, atanh},
    {"atan", fu_fast_atan, atan},    {"asin", fu_fast_asin, asin},    {"acos", fu_fast_acos, acos},
};

int main(int argc, char** argv) {
    const unsigned long stride = argc > 1 ? strtoul(argv[1], 0, 10) : 1;
    const unsigned long first = argc > 2 ? strtoul(argv[2], 0, 10) : 0;
    const char* only = argc > 3 ? argv[3] : 0;
    long total_bad = 0;
    for (unsigned f = 0; f < sizeof table / sizeof table[0]; f++) {
        if (only && strcmp(only, table[f].name)) continue;
        long acc = 0, rej = 0, bad = 0, rej_fin = 0;
#pragma omp parallel for reduction(+ : acc, rej, bad, rej_fin) schedule(static, 1 << 16)
        for (unsigned long i = first; i < (1ul << 32); i += stride) {
            const uint32_t b = (uint32_t)i;
            float x, got = 0.0f;
            memcpy(&x, &b, 4);
            if (table[f].fast(x, &got)) {
                acc++;
                const float want = (float)table[f].ref((double)x);
                uint32_t g, w;
                memcpy(&g, &got, 4), memcpy(&w, &want, 4);
                if (g != w) {
                    bad++;
                    if (bad < 5) printf("%s MISMATCH x=%a (%08x) got=%a want=%a\n", table[f].name, x, b, got, want);
                }
            } else {
                rej++;
                rej_fin += (b & 0x7fffffffu) > 0x38800000u && (b & 0x7fffffffu) < 0x42800000u;  // 2^-14 < |x| < 64
            }
        }
        printf("%-6s accepted %ld rejected %ld (of those with 2^-14 < |x| < 64: %ld) mismatches %ld\n", table[f].name, acc, rej, rej_fin, bad);
        total_bad += bad;
    }
    printf("TOTAL mismatches %ld\n", total_bad);
    return total_bad != 0;
}
