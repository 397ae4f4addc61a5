// One slice of the elementwise engine's template instantiations (split so the slices compile in parallel).
#include "elementwise.cuh"
#include "stream_tma.cuh"

namespace al {

template <template <int> class FT, int N, typename Make>
static bool dispatch_op(int op, const EwCall& c, const Folded& fo, Make make) {
    if constexpr (N == 0) {
        set_error("ValueError: unknown op code %d", op);
        return false;
    } else {
        if (op == N - 1) return run_functor(make(FT<N - 1>{}), c, fo);
        return dispatch_op<FT, N - 1>(op, c, fo, make);
    }
}

bool launch_ew_binary(const EwCall& c, const Folded& fo) {
    if (rt().stream_tma && (c.op == AL_SUM || c.op == AL_MUL)) {  // experiment, see stream_tma.cuh
        const float* in[2] = {c.in[0], c.in[1]};
        const int r = c.op == AL_SUM ? try_launch_stream_tma(BinF<AL_SUM>{}, fo, in, c.out)
                                     : try_launch_stream_tma(BinF<AL_MUL>{}, fo, in, c.out);
        if (r != 0) return r > 0;
    }
    return dispatch_op<BinF, AL_BINOP_COUNT>(c.op, c, fo, [](auto f) { return f; });
}

}  // namespace al
