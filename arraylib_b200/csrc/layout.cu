// Host-side layout folding: turns (shape, per-operand strides) into the smallest equivalent
// iteration space. This replaces the reference's per-element offset_indexer walk
// (arraylib/src/arraylib.c:116-123) with work done once per call.
#include "al_internal.h"

namespace al {

bool fold_layout(size_t ndim, const size_t* shape, const size_t* out_stride, int nin,
                 const size_t* const* in_stride, Folded* f) {
    // Walk logical axes from innermost to outermost, appending or merging.
    int n = 0;
    for (size_t ax = ndim; ax-- > 0;) {
        uint64_t e = shape[ax];
        if (e == 1) continue;  // contributes nothing to any offset
        int64_t so = out_stride ? (int64_t)out_stride[ax] : 0;
        if (n > 0) {
            // `ax` is the next-outer neighbour of folded axis n-1: mergeable when stepping it once
            // equals stepping the inner axis `extent` times, in the output and in every operand.
            uint64_t inner = f->extent[n - 1];
            bool merge = !out_stride || so == f->out_stride[n - 1] * (int64_t)inner;
            for (int i = 0; i < nin && merge; i++)
                merge = (int64_t)in_stride[i][ax] == f->in_stride[i][n - 1] * (int64_t)inner;
            if (merge) {
                f->extent[n - 1] = inner * e;
                continue;
            }
        }
        if (n == kMaxDims) {
            set_error("NotImplementedError: layout needs more than %d non-mergeable axes", kMaxDims);
            return false;
        }
        f->extent[n] = e;
        f->out_stride[n] = so;
        for (int i = 0; i < nin; i++) f->in_stride[i][n] = (int64_t)in_stride[i][ax];
        n++;
    }
    if (n == 0) {  // every extent was 1: a single element
        f->extent[0] = 1;
        f->out_stride[0] = 1;
        for (int i = 0; i < nin; i++) f->in_stride[i][0] = 1;
        n = 1;
    }
    f->ndim = n;
    return true;
}

}  // namespace al
