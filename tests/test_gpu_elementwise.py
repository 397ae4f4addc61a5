"""GPU parity: broadcast binary / scalar / unary / copy / where against the CPU oracle.

All elementwise, broadcast, view and copy results must be BIT-EXACT. The
checker is oracle/oracle.c (itself pinned against the compiled reference in test_oracle.py) and,
when oracle/_ref is present, the reference's own machine code.
"""
import math

import numpy as np
import pytest

from tests.conftest import assert_bit_equal

pytestmark = pytest.mark.gpu

BINARY = ["sum", "sub", "mul", "div", "mod", "pow", "max", "min", "eq", "ne", "gt", "ge", "lt", "le"]
AL_NAME = {"sum": "add"}
EXACT_UNARY = ["neg", "abs", "sqrt", "ceil", "floor"]
LIBM_UNARY = ["exp", "log", "sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "asinh", "acosh",
              "atanh"]

SHAPE_PAIRS = [
    ((7,), (7,)), ((1,), (5,)), ((4, 1), (1, 6)), ((1, 2, 3), (1, 1, 2, 3)), ((1, 2, 3), (2, 1)),
    ((5, 1, 7), (3, 1)), ((8, 1, 64), (1, 8, 64)), ((8, 8, 64), (64,)), ((3, 4, 5, 6), (5, 1)),
    ((33, 1, 1, 36), (1, 7, 1, 36)), ((2, 3, 4, 5, 6), (4, 1, 6)), ((1024,), (1024,)), ((1031,), (1031,)),
    ((257, 129), (129,)), ((16, 4096), (16, 1)),
    # outer-broadcast (register-tiled) plan: ragged block edges, batch axes, inner broadcast
    ((6, 1, 5, 8), (1, 9, 1, 8)), ((4, 5, 1, 16), (4, 1, 7, 16)), ((9, 1, 1), (1, 10, 8)), ((1, 13, 12), (11, 1, 12)),
    ((2, 3, 1, 6, 1, 4), (1, 1, 5, 1, 7, 4)), ((64, 1, 256), (1, 48, 256)),
]


def special_values(rng, shape):
    """Random normals salted with the awkward floats: +-0, +-inf, nan, denormals, huge, tiny."""
    x = rng.standard_normal(shape).astype(np.float32)
    flat = x.reshape(-1)
    specials = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-42, -1e-42, 3.4e38, -3.4e38, 1.0, -1.0, 2.0, 0.5],
                        dtype=np.float32)
    n = min(flat.size // 3, specials.size)
    if n:
        pos = rng.choice(flat.size, size=n, replace=False)
        flat[pos] = specials[:n]
    return x


def fn(al, op):
    return getattr(al, AL_NAME.get(op, op))


@pytest.mark.parametrize("op", BINARY)
@pytest.mark.parametrize("sa,sb", SHAPE_PAIRS)
def test_binary_broadcast_bit_exact(al, orc, rng, op, sa, sb):
    a, b = special_values(rng, sa), special_values(rng, sb)
    got = al.to_numpy(fn(al, op)(al.from_numpy(a), al.from_numpy(b)))
    assert_bit_equal(got, orc.binary(op, a, b), f"{op} {sa}x{sb}")


@pytest.mark.parametrize("op", BINARY)
def test_binary_against_compiled_reference(al, ref, rng, op):
    if ref is None:
        pytest.skip("oracle/_ref not built")
    a, b = special_values(rng, (6, 1, 40)), special_values(rng, (5, 40))
    want = ref.binary(op, ref.from_numpy(a), ref.from_numpy(b)).numpy()
    got = al.to_numpy(fn(al, op)(al.from_numpy(a), al.from_numpy(b)))
    assert_bit_equal(got, want, op)
    want = ref.binary(op, ref.from_numpy(a), 0.75).numpy()
    assert_bit_equal(al.to_numpy(fn(al, op)(al.from_numpy(a), 0.75)), want, op + " scalar")


@pytest.mark.parametrize("op", BINARY)
@pytest.mark.parametrize("scalar", [1.0, -2.5, 0.0, 3.0])
def test_scalar_ops_bit_exact(al, orc, rng, op, scalar):
    a = special_values(rng, (5, 37))
    got = al.to_numpy(fn(al, op)(al.from_numpy(a), scalar))
    assert_bit_equal(got, orc.binary(op, a, scalar), f"{op} scalar {scalar}")


def test_signed_zero_and_nan_tie_rules(al, orc):
    """max/min: NaN-ignoring and the LEFT operand wins a +-0 tie."""
    a = np.array([0.0, -0.0, np.nan, 1.0, np.nan, -0.0, 0.0], dtype=np.float32)
    b = np.array([-0.0, 0.0, 1.0, np.nan, np.nan, -0.0, 0.0], dtype=np.float32)
    for op in ("max", "min"):
        got = al.to_numpy(fn(al, op)(al.from_numpy(a), al.from_numpy(b)))
        assert_bit_equal(got, orc.binary(op, a, b), op)


@pytest.mark.parametrize("n", [1, 3, 4, 5, 1023, 4096, (1 << 20) + 3, 1 << 22])
def test_contiguous_stream_sizes(al, orc, rng, n):
    a = rng.standard_normal(n).astype(np.float32)
    b = rng.standard_normal(n).astype(np.float32)
    A, B = al.from_numpy(a), al.from_numpy(b)
    assert_bit_equal(al.to_numpy(A + B), a + b, "add")
    assert_bit_equal(al.to_numpy(A * B), a * b, "mul")
    assert_bit_equal(al.to_numpy(A + 1.5), a + np.float32(1.5), "add scalar")


def test_strided_operands(al, orc, rng):
    base = rng.standard_normal((12, 20)).astype(np.float32)
    other = rng.standard_normal((20, 12)).astype(np.float32)
    A, O = al.from_numpy(base), al.from_numpy(other)
    # transposed view (a real view: source is contiguous) + contiguous
    assert_bit_equal(al.to_numpy(A.transpose((1, 0)) + O), base.T + other, "T + c")
    assert_bit_equal(al.to_numpy(O + A.transpose((1, 0))), other + base.T, "c + T")
    assert_bit_equal(al.to_numpy(A.transpose((1, 0)) * O.transpose((1, 0)).transpose((1, 0))), base.T * other, "T*T")
    # slices with steps
    v = A[1:11:3, 2:20:5]
    assert v.view and v.shape == (4, 4)
    assert_bit_equal(al.to_numpy(v), base[1:11:3, 2:20:5], "slice")
    assert_bit_equal(al.to_numpy(v - 2.0), base[1:11:3, 2:20:5] - np.float32(2), "slice - s")
    # overlapping as_strided windows
    flat = al.from_numpy(base.reshape(-1))
    w = flat.as_strided(shape=(50, 7), stride=(3, 1))
    wn = np.lib.stride_tricks.as_strided(base.reshape(-1), shape=(50, 7), strides=(12, 4))
    assert_bit_equal(al.to_numpy(w), wn, "as_strided")
    assert_bit_equal(al.to_numpy(w + w), wn + wn, "as_strided add")


@pytest.mark.parametrize("shape,perm", [((64, 96), (1, 0)), ((5, 33, 70), (2, 1, 0)), ((5, 33, 70), (0, 2, 1)),
                                        ((3, 40, 2, 50), (3, 2, 1, 0)), ((130, 257), (1, 0))])
def test_transposed_tile_path(al, orc, rng, shape, perm):
    x = rng.standard_normal(shape).astype(np.float32)
    y = rng.standard_normal(tuple(shape[p] for p in perm)).astype(np.float32)
    X, Y = al.from_numpy(x), al.from_numpy(y)
    xt = X.transpose(perm)
    assert xt.view
    want = np.transpose(x, perm)
    assert_bit_equal(al.to_numpy(xt + Y), want + y, "T + Y")
    assert_bit_equal(al.to_numpy(al.copy(xt, deep=True)), want, "deep copy of T")
    assert_bit_equal(al.to_numpy(xt.ravel()), want.reshape(-1), "ravel of T")
    assert_bit_equal(al.to_numpy(al.sqrt(al.abs(xt))), np.sqrt(np.abs(want)), "unary of T")


@pytest.mark.parametrize("shape,perm", [((128, 192), (1, 0)), ((3, 64, 2, 52), (0, 3, 2, 1)), ((8, 260, 68), (0, 2, 1)),
                                        ((4, 100, 36), (0, 2, 1))])
def test_tile64_equals_tile32(al, rng, shape, perm):
    """Shapes whose extents are multiples of 4 take the 128-bit transpose tile; the scalar 32x32
    tile must give the same bits."""
    from arraylib_b200._lib import lib
    x = rng.standard_normal(shape).astype(np.float32)
    y = rng.standard_normal(tuple(shape[p] for p in perm)).astype(np.float32)
    X, Y = al.from_numpy(x), al.from_numpy(y)
    wide = [al.to_numpy(X.transpose(perm) - Y), al.to_numpy(al.copy(X.transpose(perm), deep=True))]
    assert lib.al_last_kernel() == b"ew_tile64_transpose"
    lib.al_set_tuning(b"tile64", 0)
    try:
        narrow = [al.to_numpy(X.transpose(perm) - Y), al.to_numpy(al.copy(X.transpose(perm), deep=True))]
        assert lib.al_last_kernel() == b"ew_tile_transpose"
    finally:
        lib.al_set_tuning(b"tile64", 1)
    for wv, nv, ref_ in zip(wide, narrow, [np.transpose(x, perm) - y, np.transpose(x, perm)]):
        assert_bit_equal(wv, nv, "tile64 vs tile32")
        assert_bit_equal(wv, ref_, "tile64 vs numpy")
    for tq in (32, 128):  # the other tile aspect ratios
        lib.al_set_tuning(b"tile_tq", tq)
        try:
            assert_bit_equal(al.to_numpy(X.transpose(perm) - Y), np.transpose(x, perm) - y, f"tile_tq={tq}")
            assert_bit_equal(al.to_numpy(al.copy(X.transpose(perm), deep=True)), np.transpose(x, perm), f"copy tile_tq={tq}")
        finally:
            lib.al_set_tuning(b"tile_tq", 64)


def test_outer_broadcast_plan_is_used_and_exact(al, orc, rng):
    from arraylib_b200._lib import lib
    a = rng.standard_normal((37, 1, 64)).astype(np.float32)
    b = rng.standard_normal((1, 29, 64)).astype(np.float32)
    A, B = al.from_numpy(a), al.from_numpy(b)
    for op in ("sub", "div", "pow", "gt"):
        got = al.to_numpy(getattr(al, op)(A, B))
        assert lib.al_last_kernel() == b"ew_outer_bcast"
        assert_bit_equal(got, orc.binary(op, a, b), op)
        got = al.to_numpy(getattr(al, op)(B, A))  # operand roles swapped
        assert_bit_equal(got, orc.binary(op, b, a), op + " swapped")
    lib.al_set_tuning(b"outer_tile", 0)
    try:
        plain = al.to_numpy(A - B)
        assert lib.al_last_kernel() == b"ew_nd_vec4"
    finally:
        lib.al_set_tuning(b"outer_tile", 1)
    assert_bit_equal(plain, a - b, "nd kernel")


@pytest.mark.parametrize("shape,perm", [((5000, 3), (1, 0)), ((3, 5000), (1, 0)), ((4096, 4), (1, 0)), ((2, 70000), (1, 0)),
                                        ((64, 500, 2), (0, 2, 1)), ((9, 7, 301), (0, 2, 1)), ((31, 33), (1, 0)), ((6, 4, 5, 400), (0, 2, 3, 1)),
                                        ((1027, 5), (1, 0)), ((16, 1 << 16), (1, 0))])
def test_short_axis_transposes(al, rng, shape, perm):
    """Interleaved <-> planar layouts ([N,3] <-> [3,N]): the packed interleave kernels (128-bit, when the
    long extent is a multiple of 4) and the rectangular tile kernel must both be bit-exact."""
    from arraylib_b200._lib import lib
    x = rng.standard_normal(shape).astype(np.float32)
    want = np.transpose(x, perm)
    y = rng.standard_normal(want.shape).astype(np.float32)
    X, Y = al.from_numpy(x), al.from_numpy(y)
    seen = set()
    for ilv in (1, 0):
        lib.al_set_tuning(b"interleave", ilv)
        try:
            got = al.to_numpy(al.copy(X.transpose(perm), deep=True))
            seen.add(lib.al_last_kernel())
            assert_bit_equal(got, want, f"copy interleave={ilv}")
            assert_bit_equal(al.to_numpy(X.transpose(perm) * Y), want * y, f"mul interleave={ilv}")
            assert_bit_equal(al.to_numpy(Y - X.transpose(perm)), y - want, f"sub, transposed rhs, interleave={ilv}")
            assert_bit_equal(al.to_numpy(X.transpose(perm) + 2.5), want + np.float32(2.5), f"scalar interleave={ilv}")
        finally:
            lib.al_set_tuning(b"interleave", 1)
    assert b"ew_tile_rect_transpose" in seen, seen
    if shape in ((5000, 3), (3, 5000), (4096, 4), (2, 70000), (64, 500, 2), (16, 1 << 16)):
        assert seen & {b"ew_interleave", b"ew_deinterleave"}, seen


@pytest.mark.parametrize("n", [4096 * 148, 4096 * 148 * 3 + 4096 + 8, (1 << 23) + 4, 1 << 24])
def test_bulk_async_stream_experiment(al, rng, n):
    """The cp.async.bulk + mbarrier streaming kernel (off by default) computes the same bits."""
    from arraylib_b200._lib import lib
    a, b = rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)
    A, B = al.from_numpy(a), al.from_numpy(b)
    lib.al_set_tuning(b"stream_tma", 1)
    try:
        s = al.to_numpy(A + B)
        assert lib.al_last_kernel() == b"ew_stream_tma"
        m = al.to_numpy(A * B)
    finally:
        lib.al_set_tuning(b"stream_tma", 0)
    assert_bit_equal(s, a + b, "tma add")
    assert_bit_equal(m, a * b, "tma mul")


def test_general_fallback_equals_fast_paths(al, rng):
    """force_general=1 routes everything through the scalar gather kernel; results must not move."""
    from arraylib_b200._lib import lib
    a = rng.standard_normal((9, 1, 64)).astype(np.float32)
    b = rng.standard_normal((1, 11, 64)).astype(np.float32)
    t = rng.standard_normal((64, 99)).astype(np.float32)
    A, B, T = al.from_numpy(a), al.from_numpy(b), al.from_numpy(t)
    fast = [al.to_numpy(A + B), al.to_numpy(al.copy(T.transpose((1, 0)), deep=True))]
    assert lib.al_set_tuning(b"force_general", 1) == 0
    try:
        slow = [al.to_numpy(A + B), al.to_numpy(al.copy(T.transpose((1, 0)), deep=True))]
        assert lib.al_last_kernel() == b"ew_nd_scalar"
    finally:
        lib.al_set_tuning(b"force_general", 0)
    for f, s in zip(fast, slow):
        assert_bit_equal(f, s, "fast vs general")


@pytest.mark.parametrize("op", EXACT_UNARY)
def test_unary_exact(al, orc, rng, op):
    x = special_values(rng, (7, 129))
    got = al.to_numpy(getattr(al, op)(al.from_numpy(x)))
    assert_bit_equal(got, orc.unary(op, x), op)


@pytest.mark.parametrize("op", LIBM_UNARY)
def test_unary_libm_within_one_ulp(al, orc, rng, op):
    """The reference rounds a double libm result to f32. Device FP64 libm is within
    1-2 double ulp of glibc's, so the f32 results may differ by one f32 ulp on a ~2^-28 fraction of
    inputs. Stated tolerance: <= 1 f32 ulp, and <= 1e-4 of the elements differ at all."""
    x = (rng.standard_normal(1 << 16) * 3).astype(np.float32)
    if op in ("asin", "acos", "atanh"):
        x = np.tanh(x).astype(np.float32)
    if op == "acosh":
        x = (np.abs(x) + 1).astype(np.float32)
    if op == "log":
        x = np.abs(x).astype(np.float32) + np.float32(1e-3)
    got = al.to_numpy(getattr(al, op)(al.from_numpy(x)))
    want = orc.unary(op, x)
    finite = np.isfinite(want)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ulp = np.abs(got[finite].view(np.int32).astype(np.int64) - want[finite].view(np.int32).astype(np.int64))
    assert ulp.max() <= 1, f"{op}: max ulp {ulp.max()}"
    assert (ulp != 0).mean() <= 1e-4, f"{op}: {(ulp != 0).mean():.2e} of elements differ"


def test_where(al, orc, rng):
    c = special_values(rng, (6, 50))
    l, r = rng.standard_normal((6, 50)).astype(np.float32), rng.standard_normal((6, 50)).astype(np.float32)
    C, L, R = al.from_numpy(c), al.from_numpy(l), al.from_numpy(r)
    assert_bit_equal(al.to_numpy(al.where(C, L, R)), orc.where(c, l, r), "aaa")
    assert_bit_equal(al.to_numpy(al.where(C, L, 2.0)), orc.where(c, l, 2.0), "aas")
    assert_bit_equal(al.to_numpy(al.where(C, -1.0, R)), orc.where(c, -1.0, r), "asa")


def test_reversed_scalar_operands(al, rng):
    a = rng.standard_normal(100).astype(np.float32)
    A = al.from_numpy(a)
    assert_bit_equal(al.to_numpy(1.5 + A), np.float32(1.5) + a)
    assert_bit_equal(al.to_numpy(1.5 - A), np.float32(1.5) - a)
    assert_bit_equal(al.to_numpy(2.0 * A), np.float32(2) * a)
    assert_bit_equal(al.to_numpy(2.0 / A), np.float32(2) / a)


def test_errors_raise_instead_of_aborting(al):
    a, b = al.ones((2, 3)), al.ones((4,))
    with pytest.raises(AssertionError):
        a + b
    from arraylib_b200._lib import lib
    assert not lib.array_array_sum(a.buffer, b.buffer)
    assert lib.al_last_error().startswith(b"ValueError")
    with pytest.raises(ValueError):
        al.arange(10).as_strided(shape=(8, 4), stride=(1, 1))  # would read past the buffer
    with pytest.raises(NotImplementedError):
        al.apply(lambda v: math.erf(v), a)
    assert not lib.array_op(None, a.buffer)
    assert lib.al_last_error().startswith(b"NotImplementedError")


@pytest.mark.parametrize("sa,sb", [((50001, 3), (3,)), ((3,), (50001, 3)), ((50001, 3), (50001, 1)), ((4001, 1, 7), (1, 9, 7)),
                                   ((301, 1, 3), (1, 257, 3)), ((2, 33, 1, 1, 5), (1, 1, 7, 3, 5)), ((33, 1, 5, 3), (1, 29, 1, 3)),
                                   ((301, 1, 3), (1, 256, 3)), ((1000, 24), (24,)), ((1000, 3, 8), (3, 8)), ((77, 1, 6), (1, 40, 6)),
                                   ((1000, 5, 3), (5, 1)), ((333, 3), (333, 3)), ((2, 50001), (2, 1))])
@pytest.mark.parametrize("op", ["sum", "sub", "div", "max", "lt"])
def test_odd_inner_extents_congruent_operand_takes_128_bit_loads(al, orc, rng, sa, sb, op):
    """The 4-outputs-per-thread fallback: an operand laid out exactly like the output is read with one float4
    per thread, the broadcast operand through stepped scalar loads; also with misaligned operands and outputs
    whose element count is not a multiple of 4."""
    a, b = rng.standard_normal(sa).astype(np.float32), rng.standard_normal(sb).astype(np.float32)
    got = getattr(al, "add" if op == "sum" else op)(al.from_numpy(a), al.from_numpy(b))
    assert_bit_equal(al.to_numpy(got), orc.binary(op, a, b), f"{op} {sa} {sb}")
    # the same operands as misaligned views of a padded base (4-byte shifted pointers)
    pa, pb = np.zeros(a.size + 1, np.float32), np.zeros(b.size + 1, np.float32)
    pa[1:], pb[1:] = a.reshape(-1), b.reshape(-1)
    def row_major(shape):
        return tuple(int(np.prod(shape[i + 1:])) for i in range(len(shape)))

    A = al.from_numpy(pa)[1:a.size + 1].as_strided(shape=sa, stride=row_major(sa))
    B = al.from_numpy(pb)[1:b.size + 1].as_strided(shape=sb, stride=row_major(sb))
    got = getattr(al, "add" if op == "sum" else op)(A, B)
    assert_bit_equal(al.to_numpy(got), orc.binary(op, a, b), f"{op} {sa} {sb} shifted")


def _special_values():
    bits = [0x7FC00000, 0xFFC00000, 0x7FC12345, 0xFFD54321, 0x7F812345, 0xFF800001,  # +-qNaN, payloads, signalling NaNs
            0x7F800000, 0xFF800000, 0x00000000, 0x80000000, 0x3F800000, 0xBF800000, 0x40000000, 0xC1000000,
            0x3F000000, 0x40600000, 0x00012345, 0x7F61B1E6, 0xC0490FDB]
    return np.array(bits, dtype=np.uint32).view(np.float32)


def _assert_nan_strict(got, want, what, ulp=0):
    """Bit-exact INCLUDING the sign and payload of NaNs; finite results within `ulp` units in the last place."""
    got, want = np.ascontiguousarray(got, np.float32), np.ascontiguousarray(want, np.float32)
    gb, wb = got.view(np.uint32), want.view(np.uint32)
    nan = np.isnan(want) | np.isnan(got)
    bad = nan & (gb != wb)
    assert not bad.any(), f"{what}: NaN bits differ at {np.argwhere(bad)[:4].tolist()}: got {[hex(v) for v in gb[bad][:4]]} want {[hex(v) for v in wb[bad][:4]]}"
    fin = ~nan
    if ulp == 0:
        assert np.array_equal(gb[fin], wb[fin]), f"{what}: finite results differ"
    else:
        same_sign = (gb[fin] >> 31) == (wb[fin] >> 31)
        dist = np.abs(gb[fin].astype(np.int64) - wb[fin].astype(np.int64))
        assert np.all((dist <= ulp) & (same_sign | (dist == 0)) | (got[fin] == want[fin])), f"{what}: more than {ulp} ulp"


@pytest.mark.parametrize("op", ["sum", "sub", "mul", "div", "mod", "max", "min", "pow", "eq", "lt", "ge"])
def test_nan_sign_and_payload_follow_the_reference_binary(al, orc, op):
    """The reference is an x86-64 / glibc build: SSE ops return their first NaN operand (quieted) and create the
    NEGATIVE default NaN, fmod follows fprem, gcc's inline fmax/fmin return the right operand when the left one is
    NaN. The full grid of special values against the oracle, bits and all."""
    s = _special_values()
    l, r = np.repeat(s, s.size).reshape(s.size, s.size), np.tile(s, s.size).reshape(s.size, s.size)
    fn = getattr(al, "add" if op == "sum" else op)
    with np.errstate(all="ignore"):
        want = orc.binary(op, l, r)
    _assert_nan_strict(al.to_numpy(fn(al.from_numpy(l), al.from_numpy(r))), want, op, ulp=1 if op == "pow" else 0)
    for scalar in (0.0, 1.0, -2.5, float("inf"), float("nan")):
        with np.errstate(all="ignore"):
            want = orc.binary(op, s, scalar)
        _assert_nan_strict(al.to_numpy(fn(al.from_numpy(s), scalar)), want, f"{op} scalar {scalar}", ulp=1 if op == "pow" else 0)
    # fused chains run the bare arithmetic (they are issue-bound): same values, NaNs wherever the eager path has NaNs,
    # but the GPU's canonical NaN instead of the x86 sign / payload
    if op in ("sum", "sub", "mul", "div"):
        lazy = {"sum": lambda a, b: a + b, "sub": lambda a, b: a - b, "mul": lambda a, b: a * b, "div": lambda a, b: a / b}[op]
        got = lazy(al.lazy(al.from_numpy(l)), al.from_numpy(r)).eval()
        with np.errstate(all="ignore"):
            assert_bit_equal(al.to_numpy(got), orc.binary(op, l, r), f"{op} fused")


def test_nan_sign_and_payload_follow_the_reference_unary(al, orc):
    s = np.concatenate([_special_values(), np.float32([2.0, -2.0, 0.25, -0.25, 1e30, -1e30, 100.0, -100.0])])
    for op in orc.UNARY:
        with np.errstate(all="ignore"):
            want = orc.unary(op, s)
        exact = op in ("neg", "abs", "sqrt", "ceil", "floor")
        _assert_nan_strict(al.to_numpy(getattr(al, op)(al.from_numpy(s))), want, op, ulp=0 if exact else 1)


def test_pow_and_log_fast_paths_match_glibc_bit_for_bit(al, orc, rng):
    """csrc/fast_pow.h answers most pow / log elements with ~40 FP64 operations and hands the rest to the full
    double-precision routine. Against the oracle (glibc pow / log in double, rounded once: arraylib.c:17, 35) on 2^21
    operands per mix: positive bases with real exponents, negative bases with integer and non-integer exponents,
    squares, values next to 1, huge and tiny results, every special value."""
    n = 1 << 21
    def near_one(k):
        return (np.float32(1.0) + (rng.integers(-1024, 1025, size=k).astype(np.float32) * np.float32(2.0 ** -23))).astype(np.float32)
    any_bits = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32).view(np.float32)
    mixes = {
        "pos^real": (rng.uniform(0.5, 4.0, n).astype(np.float32), rng.uniform(-4.0, 4.0, n).astype(np.float32)),
        "normal^normal": (rng.standard_normal(n).astype(np.float32) * 3, rng.standard_normal(n).astype(np.float32) * 3),
        "any^smallint": (rng.uniform(-30.0, 30.0, n).astype(np.float32), rng.integers(-20, 21, size=n).astype(np.float32)),
        "near1^big": (near_one(n), rng.uniform(-1e6, 1e6, n).astype(np.float32)),
        "big^big": (rng.uniform(0.0, 1e6, n).astype(np.float32), rng.uniform(-40.0, 40.0, n).astype(np.float32)),
        "bits^bits": (any_bits, np.roll(any_bits, 12345)),
        "x^2": (rng.standard_normal(n).astype(np.float32) * np.float32(1e3), np.full(n, 2.0, np.float32)),
    }
    with np.errstate(all="ignore"):
        for name, (x, y) in mixes.items():
            got = al.to_numpy(al.pow(al.from_numpy(x), al.from_numpy(y)))
            assert_bit_equal(got, orc.binary("pow", x, y), f"pow {name}")
        got = al.to_numpy(al.from_numpy(mixes["x^2"][0]) ** 2.0)  # scalar exponent
        assert_bit_equal(got, orc.binary("pow", mixes["x^2"][0], 2.0), "pow scalar 2")
        got = al.to_numpy(al.from_numpy(mixes["pos^real"][0]) ** 0.5)
        assert_bit_equal(got, orc.binary("pow", mixes["pos^real"][0], 0.5), "pow scalar 0.5")
        for name, x in {"pos": rng.uniform(1e-30, 1e30, n).astype(np.float32), "near1": near_one(n), "normal": rng.standard_normal(n).astype(np.float32),
                        "bits": any_bits, "small": rng.uniform(0.0, 2.0, n).astype(np.float32)}.items():
            assert_bit_equal(al.to_numpy(al.log(al.from_numpy(x))), orc.unary("log", x), f"log {name}")


@pytest.mark.parametrize("sa,sb", [((4001, 1, 7), (1, 9, 7)), ((301, 1, 3), (1, 257, 3)), ((33, 1, 5, 3), (1, 29, 1, 3)),
                                   ((129, 1, 5, 3), (1, 64, 1, 3)), ((77, 1, 6), (1, 40, 6)), ((50, 5, 1, 3), (50, 1, 21, 3)),
                                   ((9, 1, 2, 1, 3), (1, 11, 1, 13, 3)), ((1031, 1, 2), (1, 3, 2)), ((600, 1, 1), (1, 7, 5))])
@pytest.mark.parametrize("op", ["sum", "div", "min", "ge"])
def test_period_plan_is_used_and_exact(al, orc, rng, sa, sb, op):
    """Short inner axes under operands outside ew_run's cheap addressing modes go through ew_period_kernel (values or
    offsets of one period staged in shared memory, four shifted copies): bit-exact against the oracle and against
    ew_run (period_plan=0), on aligned operands and on 4-byte shifted views with a ragged element count."""
    from arraylib_b200._lib import lib
    a, b = special_values(rng, sa), special_values(rng, sb)
    A, B = al.from_numpy(a), al.from_numpy(b)
    f = getattr(al, "add" if op == "sum" else op)
    got = f(A, B)
    used = lib.al_last_kernel().decode()
    assert_bit_equal(al.to_numpy(got), orc.binary(op, a, b), f"{op} {sa} {sb}")
    lib.al_set_tuning(b"period_plan", 0)
    try:
        slow = al.to_numpy(f(A, B))
        assert lib.al_last_kernel().decode() != "ew_period"
    finally:
        lib.al_set_tuning(b"period_plan", 1)
    assert_bit_equal(al.to_numpy(got), slow, "period plan vs ew_run")
    if sa not in ((600, 1, 1), (301, 1, 3), (1031, 1, 2)):  # (a prime middle extent cannot be cut into periods: ew_run)
        assert used == "ew_period", used
    pa, pb = np.zeros(a.size + 1, np.float32), np.zeros(b.size + 1, np.float32)
    pa[1:], pb[1:] = a.reshape(-1), b.reshape(-1)

    def row_major(shape):
        return tuple(int(np.prod(shape[i + 1:])) for i in range(len(shape)))

    A2 = al.from_numpy(pa)[1:a.size + 1].as_strided(shape=sa, stride=row_major(sa))
    B2 = al.from_numpy(pb)[1:b.size + 1].as_strided(shape=sb, stride=row_major(sb))
    assert_bit_equal(al.to_numpy(f(A2, B2)), orc.binary(op, a, b), f"{op} {sa} {sb} shifted")


def test_period_plan_permuted_views_and_unary(al, orc, rng):
    """A permuted (non-contiguous) operand with a short axis, and a single hard operand under a unary / scalar op."""
    x, y = rng.standard_normal((5, 300, 3)).astype(np.float32), rng.standard_normal((300, 1, 3)).astype(np.float32)
    X = al.from_numpy(x).transpose((1, 0, 2))  # [300, 5, 3] with strides (3, 900, 1)
    got = X + al.from_numpy(y)
    assert_bit_equal(al.to_numpy(got), orc.binary("sum", np.ascontiguousarray(x.transpose(1, 0, 2)), y), "permuted operand")
    xt = np.ascontiguousarray(x.transpose(1, 0, 2))
    assert_bit_equal(al.to_numpy(X * 1.5), orc.binary("mul", xt, 1.5), "scalar op on a permuted view")
    assert_bit_equal(al.to_numpy(al.copy(X, deep=True)), xt, "deep copy of a permuted view")
    assert_bit_equal(al.to_numpy(al.sqrt(al.abs(X))), orc.unary("sqrt", orc.unary("abs", xt)), "unary on a permuted view")


LIBM_DOMAINS = {
    "exp": (-87.0, 87.0), "log": (1e-30, 1e30), "sin": (-200.0, 200.0), "cos": (-200.0, 200.0), "tan": (-200.0, 200.0),
    "asin": (-1.0, 1.0), "acos": (-1.0, 1.0), "atan": (-50.0, 50.0), "sinh": (-20.0, 20.0), "cosh": (-20.0, 20.0),
    "tanh": (-10.0, 10.0), "asinh": (-1e4, 1e4), "acosh": (1.0, 1e4), "atanh": (-1.0, 1.0),
}


@pytest.mark.parametrize("op", LIBM_UNARY)
def test_libm_ufunc_fast_paths_match_glibc_bit_for_bit(al, orc, rng, op):
    """csrc/fast_ufunc.h answers the elements whose float rounding is certain with a short FP64 kernel (validated on
    the CPU against glibc over all 2^32 operands); the rest run CUDA's full double-precision routine. On the GPU, against
    the oracle (glibc in double, rounded once: arraylib.c:31-49): 2^21 operands uniform over the function's domain, 2^21
    normal operands, 2^21 operands next to 0 and 1 must be BIT-EXACT; over 2^21 random bit patterns (half of them outside
    every fast path's range, i.e. CUDA's routine) at most 1 ulp on at most 1e-5 of the elements."""
    n = 1 << 21
    lo, hi = LIBM_DOMAINS[op]
    near = np.concatenate([rng.standard_normal(n // 2).astype(np.float32) * np.float32(1e-3),
                           (np.float32(1.0) + rng.integers(-4096, 4097, size=n // 2).astype(np.float32) * np.float32(2.0 ** -23)).astype(np.float32)])
    mixes = {"uniform": rng.uniform(lo, hi, n).astype(np.float32), "normal": rng.standard_normal(n).astype(np.float32) * np.float32(2.0),
             "near 0 and 1": near}
    with np.errstate(all="ignore"):
        for name, x in mixes.items():
            got = al.to_numpy(getattr(al, op)(al.from_numpy(x)))
            assert_bit_equal(got, orc.unary(op, x), f"{op} {name}")
        x = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32).view(np.float32)
        got = al.to_numpy(getattr(al, op)(al.from_numpy(x)))
        want = orc.unary(op, x)
    nan = np.isnan(want)
    assert np.array_equal(np.isnan(got), nan)
    gi, wi = got[~nan].view(np.int32).astype(np.int64), want[~nan].view(np.int32).astype(np.int64)
    ulp = np.abs(gi - wi)
    assert ulp.max() <= 1 and (ulp != 0).mean() <= 1e-5, f"{op} any bits: max {ulp.max()} ulp, {(ulp != 0).sum()} differ"
