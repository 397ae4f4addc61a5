#include <stdio.h>
#include <stdlib.h>
#include "../../arraylib_b200/csrc/fast_pow.h"
static uint64_t s = 88172645463325252ull;
static uint64_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static float rfloat(int mode) {
    uint32_t b;
    if (mode == 0) { b = (uint32_t)rnd(); }                               // any bit pattern
    else if (mode == 1) { b = 0x3f000000u + (uint32_t)(rnd() % 0x01800000u); }  // [0.5, 4)
    else if (mode == 2) { float f = (float)((double)(rnd() % 2000001) / 1000.0 - 1000.0); memcpy(&b, &f, 4); }
    else if (mode == 3) { float f = (float)(int)(rnd() % 41) - 20.0f; memcpy(&b, &f, 4); }   // small integers
    else { b = 0x3f800000u + (uint32_t)(rnd() % 2048) - 1024; }   // near 1
    float f; memcpy(&f, &b, 4); return f;
}
int main(int argc, char** argv) {
    long n = argc > 1 ? atol(argv[1]) : 100000000;
    if (argc > 2) s ^= (uint64_t)atol(argv[2]) * 0x9E3779B97F4A7C15ull;
    long fast = 0, slow = 0, bad = 0, amb = 0;
    for (int mx = 0; mx < 5; mx++) for (int my = 0; my < 5; my++) {
        long f0 = 0, s0 = 0, b0 = 0;
        for (long i = 0; i < n / 25; i++) {
            float x = rfloat(mx), y = rfloat(my), got;
            float want = (float)pow((double)x, (double)y);
            if (fp_fast_pow(x, y, &got)) {
                f0++;
                uint32_t a, b; memcpy(&a, &got, 4); memcpy(&b, &want, 4);
                if (a != b && !(got != got && want != want)) { b0++; if (bad + b0 < 20) printf("MISMATCH x=%a y=%a got=%a want=%a\n", x, y, got, want); }
            } else s0++;
        }
        printf("modes x%d y%d: fast %ld slow %ld (%.4f%%) bad %ld\n", mx, my, f0, s0, 100.0 * s0 / (f0 + s0), b0);
        fast += f0; slow += s0; bad += b0;
    }
    long lf = 0, ls = 0, lb = 0;
    for (int mx = 0; mx < 5; mx++) for (long i = 0; i < n / 25; i++) {
        float x = rfloat(mx), got;
        if (mx == 0) x = fabsf(x);
        float want = (float)log((double)x);
        if (fp_fast_log(x, &got)) {
            lf++;
            uint32_t a, b; memcpy(&a, &got, 4); memcpy(&b, &want, 4);
            if (a != b && !(got != got && want != want)) { lb++; if (lb < 20) printf("LOG MISMATCH x=%a got=%a want=%a\n", x, got, want); }
        } else ls++;
    }
    printf("log: fast %ld slow %ld bad %ld\n", lf, ls, lb);
    bad += lb;
    printf("TOTAL fast %ld slow %ld bad %ld\n", fast, slow, bad);
    return bad != 0;
}
