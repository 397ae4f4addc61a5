/* A plain-C caller of libarraylib_b200.so: the C ABI has no Python or C++ types in it.
 * Built and run by tests/test_gpu_c_caller.py:  gcc abi_smoke.c -I include -L ... -larraylib_b200 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "arraylib_b200.h"

#define CHECK(cond)                                                                       \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            fprintf(stderr, "FAILED %s (line %d): %s\n", #cond, __LINE__, al_last_error()); \
            return 1;                                                                     \
        }                                                                                 \
    } while (0)

int main(void) {
    CHECK(al_init(0) == 0);
    /* arange(1,361).reshape(3,4,5,6).reduce_sum(dims=(1,2)) -> sqrt -> reshape(3,6) -> transpose (README) */
    NDArray* a = array_arange(1, 361, 1);
    CHECK(a != NULL && a->lay->ndim == 1 && a->lay->shape[0] == 360 && a->data->refs == 1);
    size_t shape4[4] = {3, 4, 5, 6}, dims[2] = {1, 2}, shape2[2] = {3, 6}, perm[2] = {1, 0};
    NDArray* r = array_reshape(a, shape4, 4);
    CHECK(r != NULL && r->view && a->data->refs == 2);
    NDArray* s = array_reduce_sum(r, dims, 2);
    CHECK(s != NULL && s->lay->shape[0] == 3 && s->lay->shape[1] == 1 && s->lay->shape[3] == 6 && s->lay->stride[0] == 6);
    NDArray* q = array_sqrt(s);
    NDArray* m = q ? array_reshape(q, shape2, 2) : NULL;
    NDArray* t = m ? array_transpose(m, perm) : NULL;
    CHECK(t != NULL && t->lay->shape[0] == 6 && t->lay->stride[0] == 1 && t->lay->stride[1] == 6);
    float host[18];
    CHECK(array_to_host(t, host) == 0);
    CHECK(host[0] == sqrtf(1160.0f) && host[1] == sqrtf(3560.0f) && host[17] == sqrtf(6060.0f));
    /* broadcasting + scalar: ones[3,6] + ones[6] + 1.0 */
    size_t s36[2] = {3, 6}, s6[1] = {6};
    NDArray* x = array_ones(s36, 2);
    NDArray* y = array_ones(s6, 1);
    NDArray* z = array_array_sum(x, y);
    NDArray* w = z ? array_scalar_sum(z, 1.0f) : NULL;
    CHECK(w != NULL);
    size_t idx[2] = {2, 5};
    CHECK(array_get_elem(w, idx) == 3.0f);
    /* errors come back as NULL + message, never abort() */
    size_t s4[1] = {4};
    NDArray* bad = array_ones(s4, 1);
    CHECK(array_array_sum(x, bad) == NULL);
    printf("error text: %s\n", al_last_error());
    CHECK(array_op(NULL, x) == NULL);
    NDArray* all[] = {a, r, s, q, m, t, x, y, z, w, bad};
    for (size_t i = 0; i < sizeof all / sizeof all[0]; i++) array_free(all[i]);
    CHECK(al_bytes_in_use() == 0);
    printf("C ABI smoke ok, %llu kernels launched\n", (unsigned long long)al_launch_count());
    return 0;
}
