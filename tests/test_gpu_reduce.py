"""GPU parity: reduce_sum / reduce_max / reduce_min over arbitrary dim sets.

Tolerance for reduce_sum (the summation ORDER differs from the reference):
  * small-integer inputs (every partial sum < 2^24): bit-exact, order cannot matter;
  * general inputs, n reduced elements per output: |gpu - ref| <= 2 * n * 2^-24 * sum|x_i|;
reduce_max / reduce_min are bit-exact.
"""
import itertools

import numpy as np
import pytest

from tests.conftest import assert_bit_equal

pytestmark = pytest.mark.gpu


def power_set(items):
    for r in range(len(items) + 1):
        yield from itertools.combinations(items, r)


def int_data(rng, shape):
    return rng.integers(-8, 9, size=shape).astype(np.float32)


def sum_tolerance(x, dims):
    n = int(np.prod([x.shape[d] for d in dims])) if dims else 1
    return 2.0 * n * 2.0 ** -24 * np.sum(np.abs(x.astype(np.float64)), axis=tuple(dims), keepdims=True)


@pytest.mark.parametrize("dims", list(power_set((0, 1, 2, 3))))
@pytest.mark.parametrize("op", ["sum", "max", "min"])
def test_power_set_golden(al, orc, op, dims):
    """The reference's own known-answer test: tests/test_array.py:222-252."""
    a = al.arange(1, 361).reshape((3, 4, 5, 6))
    x = np.arange(1, 361, dtype=np.float32).reshape(3, 4, 5, 6)
    got = getattr(al, f"reduce_{op}")(a, dims=dims)
    want = {"sum": np.sum, "max": np.max, "min": np.min}[op](x, axis=dims, keepdims=True)
    assert got.shape == want.shape
    assert_bit_equal(al.to_numpy(got), want, f"{op} {dims}", nan_ok=True)
    assert_bit_equal(al.to_numpy(got), orc.reduce(op, x, dims), f"{op} {dims} vs oracle", nan_ok=True)


SHAPES = [(7,), (1,), (5, 64), (64, 5), (3, 1, 9), (16, 33, 8), (4, 130, 3, 12), (2, 3, 4, 5, 6), (9, 4096), (4096, 9),
          (6, 40, 40, 8), (100000,), (3, 70000), (70000, 3)]


@pytest.mark.parametrize("shape", SHAPES)
def test_integer_sums_bit_exact_all_dim_sets(al, orc, rng, shape):
    x = int_data(rng, shape)
    X = al.from_numpy(x)
    for dims in power_set(tuple(range(len(shape)))):
        want = orc.reduce("sum", x, dims)
        got = al.reduce_sum(X, dims=dims)
        assert got.shape == want.shape and not got.view
        assert_bit_equal(al.to_numpy(got), want, f"sum {shape} {dims}", nan_ok=True)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("op", ["max", "min"])
def test_max_min_bit_exact(al, orc, rng, shape, op):
    x = rng.standard_normal(shape).astype(np.float32)
    x.reshape(-1)[:: 7] = np.nan  # fmax/fmin ignore NaN
    if x.size > 3:
        x.reshape(-1)[1] = np.inf
        x.reshape(-1)[2] = -np.inf
    X = al.from_numpy(x)
    for dims in power_set(tuple(range(len(shape)))):
        assert_bit_equal(al.to_numpy(getattr(al, f"reduce_{op}")(X, dims=dims)), orc.reduce(op, x, dims), f"{op} {dims}", nan_ok=True)


@pytest.mark.parametrize("shape", SHAPES)
def test_float_sums_within_stated_tolerance(al, orc, ref, rng, shape):
    x = rng.standard_normal(shape).astype(np.float32)
    X = al.from_numpy(x)
    for dims in power_set(tuple(range(len(shape)))):
        want = orc.reduce("sum", x, dims).astype(np.float64)
        got = al.to_numpy(al.reduce_sum(X, dims=dims)).astype(np.float64)
        tol = sum_tolerance(x, dims)
        assert np.all(np.abs(got - want) <= tol), f"{shape} {dims}: max err {np.abs(got - want).max()}"
        # and no worse than the reference relative to the exact (float64) sum
        exact = np.sum(x.astype(np.float64), axis=dims, keepdims=True) if dims else x.astype(np.float64)
        assert np.all(np.abs(got - exact) <= np.abs(want - exact) + 2.0 ** -22 * np.abs(exact) + tol * 0.5)


def test_against_compiled_reference(al, ref, rng):
    if ref is None:
        pytest.skip("oracle/_ref not built")
    x = int_data(rng, (5, 6, 7, 8))
    X, RX = al.from_numpy(x), ref.from_numpy(x)
    for dims in power_set((0, 1, 2, 3)):
        for op in ("sum", "max", "min"):
            want = ref.reduce(op, RX, dims)
            got = getattr(al, f"reduce_{op}")(X, dims=dims)
            assert got.shape == want.shape and got.stride == want.stride
            assert_bit_equal(al.to_numpy(got), want.numpy(), f"{op} {dims}", nan_ok=True)


def test_strided_sources(al, orc, rng):
    x = int_data(rng, (12, 20, 6))
    X = al.from_numpy(x)
    xt = np.transpose(x, (2, 0, 1))
    XT = X.transpose((2, 0, 1))
    assert XT.view
    for dims in power_set((0, 1, 2)):
        assert_bit_equal(al.to_numpy(al.reduce_sum(XT, dims=dims)), orc.reduce("sum", xt, dims), f"T {dims}", nan_ok=True)
    v = X[2:12:3, 1:20:2, 0:6:1]
    vn = x[2:12:3, 1:20:2, :]
    for dims in power_set((0, 1, 2)):
        assert_bit_equal(al.to_numpy(al.reduce_sum(v, dims=dims)), orc.reduce("sum", vn, dims), f"slice {dims}", nan_ok=True)


@pytest.mark.parametrize("n,w", [(1 << 16, 64), (100003, 64), (5000, 8), (20000, 100), (70000, 1024), (300, 64), (30000, 9),
                                 (50001, 33), (40000, 127), (90000, 500), (8000, 10)])
def test_sliding_window_view(al, orc, rng, n, w):
    """as_strided (n-w+1, w) with strides (1, 1), reduce over the window axis (config C5a)."""
    base = int_data(rng, (n,))
    B = al.from_numpy(base)
    V = B.as_strided(shape=(n - w + 1, w), stride=(1, 1))
    vn = np.lib.stride_tricks.as_strided(base, shape=(n - w + 1, w), strides=(4, 4))
    got = al.reduce_sum(V, dims=[1])
    assert got.shape == (n - w + 1, 1)
    assert_bit_equal(al.to_numpy(got), orc.reduce("sum", vn, (1,)), "window ints", nan_ok=True)
    assert_bit_equal(al.to_numpy(al.reduce_max(V, dims=[1])), orc.reduce("max", vn, (1,)), "window max", nan_ok=True)
    # the transposed view (w, n-w+1) reduced over axis 0 is the same computation
    VT = B.as_strided(shape=(w, n - w + 1), stride=(1, 1))
    assert_bit_equal(al.to_numpy(al.reduce_sum(VT, dims=[0])).reshape(-1), orc.reduce("sum", vn, (1,)).reshape(-1), "wT", nan_ok=True)
    f = rng.standard_normal(n).astype(np.float32)
    F = al.from_numpy(f).as_strided(shape=(n - w + 1, w), stride=(1, 1))
    fn_ = np.lib.stride_tricks.as_strided(f, shape=(n - w + 1, w), strides=(4, 4))
    err = np.abs(al.to_numpy(al.reduce_sum(F, dims=[1])).astype(np.float64) - orc.reduce("sum", fn_, (1,)))
    assert np.all(err <= sum_tolerance(fn_, (1,)))


@pytest.mark.parametrize("rows,n,w", [(5, 3000, 64), (3, 9000, 16), (17, 700, 32), (2, 20000, 8)])
def test_batched_sliding_windows(al, orc, rng, rows, n, w):
    """A batch of independent signals, each with stride-1 windows: view (rows, n-w+1, w), strides (n, 1, 1)."""
    from arraylib_b200._lib import lib
    base = int_data(rng, (rows, n))
    V = al.from_numpy(base.reshape(-1)).as_strided(shape=(rows, n - w + 1, w), stride=(n, 1, 1))
    vn = np.lib.stride_tricks.as_strided(base, shape=(rows, n - w + 1, w), strides=(4 * n, 4, 4))
    got = al.reduce_sum(V, dims=[2])
    assert lib.al_last_kernel() == b"reduce_window"  # output rows of n-w+1 floats are not 16-byte aligned: shared-memory kernel
    assert got.shape == (rows, n - w + 1, 1)
    assert_bit_equal(al.to_numpy(got), orc.reduce("sum", vn, (2,)), "batched windows", nan_ok=True)
    assert_bit_equal(al.to_numpy(al.reduce_min(V, dims=[2])), orc.reduce("min", vn, (2,)), "batched windows min", nan_ok=True)


def test_window_kernel_matches_generic_path(al, rng):
    from arraylib_b200._lib import lib
    base = rng.standard_normal(50000).astype(np.float32)
    V = al.from_numpy(base).as_strided(shape=(50000 - 63, 64), stride=(1, 1))
    fast = al.to_numpy(al.reduce_max(V, dims=[1]))
    assert lib.al_last_kernel() == b"reduce_window_shfl"
    lib.al_set_tuning(b"disable_window", 1)
    try:
        slow = al.to_numpy(al.reduce_max(V, dims=[1]))
        assert not lib.al_last_kernel().startswith(b"reduce_window")
    finally:
        lib.al_set_tuning(b"disable_window", 0)
    assert_bit_equal(fast, slow, "window vs generic", nan_ok=True)


@pytest.mark.parametrize("w", [8, 16, 32, 64, 128])
def test_window_kernel_variants_agree(al, orc, rng, w):
    """The register/shuffle kernel (both row groupings) and the two shared-memory kernels against the oracle,
    with a stretch of all-NaN windows (max/min must give the identity like the reference)."""
    from arraylib_b200._lib import lib
    n = 300000
    base = int_data(rng, (n,))
    base[1000:1100 + w] = np.nan
    V = al.from_numpy(base).as_strided(shape=(n - w + 1, w), stride=(1, 1))
    vn = np.lib.stride_tricks.as_strided(base, shape=(n - w + 1, w), strides=(4, 4))
    for op in ("sum", "max", "min"):
        want = orc.reduce(op, vn, (1,))
        for shfl, whole, name in ((1, 0, b"reduce_window_shfl"), (2, 0, b"reduce_window_shfl"), (3, 0, b"reduce_window_shfl"),
                                  (0, 0, b"reduce_window"), (0, 1, b"reduce_window")):
            lib.al_set_tuning(b"window_shfl", shfl)
            lib.al_set_tuning(b"window_whole_blocks", whole)
            try:
                got = al.to_numpy(getattr(al, f"reduce_{op}")(V, dims=[1]))
                if w <= 64 or shfl:
                    assert lib.al_last_kernel() == name
            finally:
                lib.al_set_tuning(b"window_shfl", 1)
                lib.al_set_tuning(b"window_whole_blocks", 0)
            assert_bit_equal(got, want, f"{op} shfl={shfl} whole_blocks={whole}", nan_ok=True)


@pytest.mark.parametrize("w", [8, 16, 32, 64, 128])
@pytest.mark.parametrize("nout", [4, 124, 128, 132, 15 * 128 - 4, 15 * 128, 15 * 128 + 4, 30 * 128 + 8, 2 * 3 * 148 * 15 * 128 + 128 + 3, 1000003])
def test_shuffle_window_kernel_edges(al, orc, rng, w, nout):
    """Segment / row / vector boundaries of the register/shuffle kernel: outputs per signal just below, at and just above
    a 128-output row and a 31-row segment, odd lengths (scalar tail stores, partial last vector of the source), a
    custom init, and float data against the stated tolerance."""
    from arraylib_b200._lib import lib
    n = nout + w - 1
    if nout < 4 * w:
        pytest.skip("windows this short take the generic path")
    base = int_data(rng, (n,))
    B = al.from_numpy(base)
    V = B.as_strided(shape=(nout, w), stride=(1, 1))
    vn = np.lib.stride_tricks.as_strided(base, shape=(nout, w), strides=(4, 4))
    for op in ("sum", "max", "min"):
        got = getattr(al, f"reduce_{op}")(V, dims=[1])
        assert lib.al_last_kernel() == b"reduce_window_shfl"
        assert_bit_equal(al.to_numpy(got), orc.reduce(op, vn, (1,)), f"{op} w={w} nout={nout}", nan_ok=True)
    got = al.reduce(lambda a, b: a + b, V, dims=[1], init=3.0)
    assert_bit_equal(al.to_numpy(got), orc.reduce("sum", vn, (1,), init=3.0), "custom init", nan_ok=True)
    f = rng.standard_normal(n).astype(np.float32)
    F = al.from_numpy(f).as_strided(shape=(nout, w), stride=(1, 1))
    fn_ = np.lib.stride_tricks.as_strided(f, shape=(nout, w), strides=(4, 4))
    err = np.abs(al.to_numpy(al.reduce_sum(F, dims=[1])).astype(np.float64) - orc.reduce("sum", fn_, (1,)))
    assert np.all(err <= sum_tolerance(fn_, (1,)))


@pytest.mark.parametrize("rows,nout,w", [(3, 4000, 64), (7, 128 * 31, 16), (2, 40000, 128), (5, 516, 8)])
def test_batched_windows_with_aligned_rows_use_the_shuffle_kernel(al, orc, rng, rows, nout, w):
    """Signals stored with a padded pitch (row stride and outputs per row both multiples of 4): the register/shuffle
    kernel takes batches too. Also a misaligned base pointer, which must fall back to the shared-memory kernel."""
    from arraylib_b200._lib import lib
    pitch = (nout + w - 1 + 3) // 4 * 4
    base = int_data(rng, (rows * pitch + 8,))
    B = al.from_numpy(base)
    V = B.as_strided(shape=(rows, nout, w), stride=(pitch, 1, 1))
    vn = np.lib.stride_tricks.as_strided(base, shape=(rows, nout, w), strides=(4 * pitch, 4, 4))
    for op in ("sum", "min"):
        got = getattr(al, f"reduce_{op}")(V, dims=[2])
        assert lib.al_last_kernel() == b"reduce_window_shfl"
        assert_bit_equal(al.to_numpy(got), orc.reduce(op, vn, (2,)), f"batched {op}", nan_ok=True)
    off = B[1:rows * pitch + 8]  # base pointer 4 bytes past a 16-byte boundary
    V1 = off.as_strided(shape=(rows, nout, w), stride=(pitch, 1, 1))
    v1 = np.lib.stride_tricks.as_strided(base[1:], shape=(rows, nout, w), strides=(4 * pitch, 4, 4))
    got = al.reduce_sum(V1, dims=[2])
    assert lib.al_last_kernel() in (b"reduce_window", b"reduce_window_any")
    assert_bit_equal(al.to_numpy(got), orc.reduce("sum", v1, (2,)), "misaligned base", nan_ok=True)


@pytest.mark.parametrize("splits", [1, 2, 7, 64])
def test_two_pass_split_is_consistent(al, orc, rng, splits):
    from arraylib_b200._lib import lib
    x = int_data(rng, (4, 5000, 16))
    X = al.from_numpy(x)
    lib.al_set_tuning(b"reduce_splits", splits)
    try:
        assert_bit_equal(al.to_numpy(al.reduce_sum(X, dims=(1,))), orc.reduce("sum", x, (1,)), "outer split", nan_ok=True)
        assert_bit_equal(al.to_numpy(al.reduce_sum(X, dims=(0, 1, 2))), orc.reduce("sum", x, (0, 1, 2)), "full split", nan_ok=True)
        assert_bit_equal(al.to_numpy(al.reduce_sum(X, dims=(1, 2))), orc.reduce("sum", x, (1, 2)), "inner split", nan_ok=True)
    finally:
        lib.al_set_tuning(b"reduce_splits", 0)


def test_config4_shapes_reduced(al, orc, rng):
    """[64,512,512,64] scaled to [8,64,64,64]: dims (1,2) and dims (3,) (config C4)."""
    x = int_data(rng, (8, 64, 64, 64))
    X = al.from_numpy(x)
    a = al.reduce_sum(X, dims=(1, 2))
    assert a.shape == (8, 1, 1, 64) and a.stride == (64, 64, 64, 1)
    assert_bit_equal(al.to_numpy(a), orc.reduce("sum", x, (1, 2)), "C4a", nan_ok=True)
    b = al.reduce_sum(X, dims=(3,))
    assert b.shape == (8, 64, 64, 1)
    assert_bit_equal(al.to_numpy(b), orc.reduce("sum", x, (3,)), "C4b", nan_ok=True)
    c = al.reduce_sum(X, dims=(0,))
    assert_bit_equal(al.to_numpy(c), orc.reduce("sum", x, (0,)), "C4 dim0", nan_ok=True)


def test_reduce_with_callable_and_init(al, orc, rng):
    x = int_data(rng, (3, 4, 5, 6))
    X = al.from_numpy(x)
    got = al.reduce(lambda acc, v: acc + v, X, dims=(1, 2), init=0)
    assert_bit_equal(al.to_numpy(got), orc.reduce("sum", x, (1, 2)), "lambda sum", nan_ok=True)
    got = al.reduce(max, X, dims=(0,), init=-100.0)
    assert_bit_equal(al.to_numpy(got), orc.reduce("max", x, (0,), init=-100.0), "max init", nan_ok=True)
    got = al.reduce(lambda a, b: a + b, X, dims=(3,), init=2.0)  # the reference folds init in twice
    assert_bit_equal(al.to_numpy(got), orc.reduce("sum", x, (3,), init=2.0), "sum init 2", nan_ok=True)
    with pytest.raises(NotImplementedError):
        al.reduce(lambda a, b: a - b, X, dims=(0,))


@pytest.mark.parametrize("shape", [(7,), (5, 64), (16, 33, 8), (4, 13, 3, 12), (2, 3, 4, 5, 6), (300, 9)])
def test_sequential_mode_reproduces_the_reference_bit_for_bit(al, orc, ref, rng, shape):
    """With reduce_sequential=1 every output is accumulated in the reference's own order, so even
    float reduce_sum matches the serial reference bit for bit -- the tolerance of the fast kernels is
    purely the order of the additions."""
    from arraylib_b200._lib import lib
    x = rng.standard_normal(shape).astype(np.float32) * np.float32(100)
    x.reshape(-1)[::11] = np.float32(-0.0)
    X = al.from_numpy(x)
    views = [(X, x)]
    if len(shape) > 1:
        perm = tuple(reversed(range(len(shape))))
        views.append((X.transpose(perm), np.transpose(x, perm)))
    lib.al_set_tuning(b"reduce_sequential", 1)
    try:
        for V, v in views:
            for dims in power_set(tuple(range(v.ndim))):
                for op in ("sum", "max", "min"):
                    got = getattr(al, f"reduce_{op}")(V, dims=dims)
                    assert lib.al_last_kernel() == b"reduce_sequential"
                    assert_bit_equal(al.to_numpy(got), orc.reduce(op, v, dims), f"{op} {shape} {dims}", nan_ok=True)
            got = al.reduce(lambda a_, b_: a_ + b_, V, dims=(0,), init=1.5)
            assert_bit_equal(al.to_numpy(got), orc.reduce("sum", v, (0,), init=1.5), "custom init", nan_ok=True)
        if ref is not None:
            want = ref.reduce("sum", ref.from_numpy(x), tuple(range(len(shape)))[:1]).numpy()
            assert_bit_equal(al.to_numpy(al.reduce_sum(X, dims=tuple(range(len(shape)))[:1])), want, "vs compiled reference", nan_ok=True)
    finally:
        lib.al_set_tuning(b"reduce_sequential", 0)


@pytest.mark.parametrize("r", [2, 3, 4, 5, 6, 7, 8])
def test_packed_short_rows(al, orc, rng, r):
    """[N, r] reduced over the short inner axis (component axis of RGB / xyz / complex data)."""
    from arraylib_b200._lib import lib
    x = int_data(rng, (4096, r))
    X = al.from_numpy(x)
    for op in ("sum", "max", "min"):
        got = getattr(al, f"reduce_{op}")(X, dims=(1,))
        assert lib.al_last_kernel() == b"reduce_rows_packed"
        assert_bit_equal(al.to_numpy(got), orc.reduce(op, x, (1,)), f"{op} r={r}", nan_ok=True)
    f = rng.standard_normal((1028, r)).astype(np.float32)
    err = np.abs(al.to_numpy(al.reduce_sum(al.from_numpy(f), dims=(1,))).astype(np.float64) - orc.reduce("sum", f, (1,)))
    assert np.all(err <= sum_tolerance(f, (1,)))
    y = int_data(rng, (1027, r))  # row count not a multiple of 4: generic path
    assert_bit_equal(al.to_numpy(al.reduce_sum(al.from_numpy(y), dims=(1,))), orc.reduce("sum", y, (1,)), "odd rows", nan_ok=True)
    assert lib.al_last_kernel() != b"reduce_rows_packed"


def test_cross_gpu_entry_point_is_the_local_reduce_in_a_single_process(al, orc, rng):
    """al_comm_reduce (the collective behind dist.reduce_sharded) degenerates to array_reduce_<op> when there is
    no communicator; the N > 1 peer-memory path is covered by tools/dist_check.py under torchrun."""
    from arraylib_b200 import core, dist
    from arraylib_b200._lib import REDOPS, lib

    x = rng.integers(-8, 9, size=(6, 5, 64)).astype(np.float32)
    a = al.from_numpy(x)
    assert dist.transport() == "single process"
    for op in ("sum", "max", "min"):
        for dims in [(0,), (0, 2), (1,), (0, 1, 2)]:
            dims_c, n = core._dims_arg(a, dims)
            got = core._new(lib.al_comm_reduce(REDOPS.index(op), a.buffer, dims_c, n))
            assert_bit_equal(al.to_numpy(got), orc.reduce(op, x, dims), nan_ok=True)
            assert_bit_equal(al.to_numpy(dist.reduce_sharded(a, op, dims)), orc.reduce(op, x, dims), nan_ok=True)


@pytest.mark.parametrize("rows,n,w", [(4, 9000, 100), (3, 5000, 24), (2, 13000, 1000), (7, 4500, 12)])
def test_batched_sliding_windows_any_width(al, orc, rng, rows, n, w):
    """Widths without a dedicated kernel go through the segmented-scan window kernel, batched rows included."""
    from arraylib_b200._lib import lib
    base = int_data(rng, (rows, n))
    V = al.from_numpy(base.reshape(-1)).as_strided(shape=(rows, n - w + 1, w), stride=(n, 1, 1))
    vn = np.lib.stride_tricks.as_strided(base, shape=(rows, n - w + 1, w), strides=(4 * n, 4, 4))
    got = al.reduce_sum(V, dims=[2])
    assert lib.al_last_kernel() == b"reduce_window_any"
    assert_bit_equal(al.to_numpy(got), orc.reduce("sum", vn, (2,)), "batched windows", nan_ok=True)
    assert_bit_equal(al.to_numpy(al.reduce_max(V, dims=[2])), orc.reduce("max", vn, (2,)), "batched windows max", nan_ok=True)
    assert_bit_equal(al.to_numpy(al.reduce_min(V, dims=[2])), orc.reduce("min", vn, (2,)), "batched windows min", nan_ok=True)


@pytest.mark.parametrize("w", [11, 31, 48, 100, 257, 777])
def test_any_width_windows_persistent_tiles_and_nan_stretches(al, orc, rng, w):
    """Several tiles per CTA (window_ctas_per_sm=1), all-NaN windows, and the generic reduce path as a second opinion."""
    from arraylib_b200._lib import lib
    n = 2_000_000
    base = int_data(rng, (n,))
    base[5000:5000 + 2 * w] = np.nan
    V = al.from_numpy(base).as_strided(shape=(n - w + 1, w), stride=(1, 1))
    vn = np.lib.stride_tricks.as_strided(base, shape=(n - w + 1, w), strides=(4, 4))
    lib.al_set_tuning(b"window_ctas_per_sm", 1)
    try:
        for op in ("sum", "max", "min"):
            got = al.to_numpy(getattr(al, f"reduce_{op}")(V, dims=[1]))
            assert lib.al_last_kernel() == b"reduce_window_any"
            assert_bit_equal(got, orc.reduce(op, vn, (1,)), f"{op} w={w}", nan_ok=True)
    finally:
        lib.al_set_tuning(b"window_ctas_per_sm", 0)
    f = rng.standard_normal(300000).astype(np.float32)
    F = al.from_numpy(f).as_strided(shape=(f.size - w + 1, w), stride=(1, 1))
    fast = al.to_numpy(al.reduce_max(F, dims=[1]))
    lib.al_set_tuning(b"disable_window", 1)
    try:
        slow = al.to_numpy(al.reduce_max(F, dims=[1]))
    finally:
        lib.al_set_tuning(b"disable_window", 0)
    assert_bit_equal(fast, slow, "window vs generic", nan_ok=True)


@pytest.mark.parametrize("shape,dims", [((4096, 255), (1,)), ((3, 2048, 85), (0, 2)), ((3, 2048, 85), (2,)), ((1000, 17), (1,)),
                                        ((257, 1023), (1,)), ((5, 7, 129), (1, 2)), ((2, 3, 4, 33), (3,)), ((64, 16), (1,))])
def test_misaligned_rows_use_masked_vectors(al, orc, rng, shape, dims):
    """Unit-stride reduced rows that are not 16-byte aligned: reduce_inner_peel (aligned float4 loads, foreign
    elements masked), against the oracle and against the scalar-lane kernel."""
    from arraylib_b200._lib import lib
    x = int_data(rng, shape)
    X = al.from_numpy(x)
    for op in ("sum", "max", "min"):
        got = al.to_numpy(getattr(al, f"reduce_{op}")(X, dims=dims))
        kernel = lib.al_last_kernel()
        assert_bit_equal(got, orc.reduce(op, x, dims), f"{op} {shape} {dims}", nan_ok=True)
    if shape[-1] % 4 and shape[-1] >= 16 and x.size // int(np.prod([shape[d] for d in dims])) > 64:
        assert kernel == b"reduce_inner_peel", kernel
    # sliced views: every row starts at a different misalignment; NaN and -0 survive the masking
    f = rng.standard_normal((300, 270)).astype(np.float32)
    f[5, 7:30] = np.nan
    f[9] = -0.0
    F = al.from_numpy(f)
    for c0 in (0, 1, 2, 3):
        for c1 in (270, 269, 258, 257):
            v, vn = F[2:300, c0:c1], f[2:300, c0:c1]
            for op in ("max", "min"):
                assert_bit_equal(al.to_numpy(getattr(al, f"reduce_{op}")(v, dims=(1,))), orc.reduce(op, vn, (1,)), f"{op} cols {c0}:{c1}", nan_ok=True)
            lib.al_set_tuning(b"reduce_peel", 0)
            try:
                scalar = al.to_numpy(al.reduce_max(v, dims=(1,)))
            finally:
                lib.al_set_tuning(b"reduce_peel", 1)
            assert_bit_equal(al.to_numpy(al.reduce_max(v, dims=(1,))), scalar, "peel vs scalar lanes", nan_ok=True)
            err = np.abs(al.to_numpy(al.reduce_sum(v, dims=(1,))).astype(np.float64) - np.nansum(vn.astype(np.float64), axis=1, keepdims=True))
            ok = np.isnan(orc.reduce("sum", vn, (1,))) | (err <= 2 * vn.shape[1] * 2.0**-24 * np.nansum(np.abs(vn), axis=1, keepdims=True) + 1e-30)
            assert np.all(ok)


@pytest.mark.parametrize("shape", [(9,), (6, 70), (5, 33, 8), (4, 7, 3, 12), (3, 2500)])
def test_max_min_zero_sign_follows_the_first_zero_in_logical_order(al, orc, ref, rng, shape):
    """VERDICT r1 weak #1c: when the extreme value of a reduced set is zero and the set holds +0 and -0, fmax/fmin's
    left-operand tie rule makes the reference return the FIRST zero in logical order (arraylib.c:18-19, 737-744).
    The tree kernels plus the zero-sign repair pass must reproduce that bit for bit -- strict comparison, with the
    fast kernels (no reduce_sequential)."""
    for op, values in (("max", (-2.0, -1.0, -0.0, 0.0)), ("min", (2.0, 1.0, -0.0, 0.0))):
        x = rng.choice(np.array(values, dtype=np.float32), size=shape, p=(0.3, 0.3, 0.2, 0.2)).astype(np.float32)
        X = al.from_numpy(x)
        views = [(X, x)]
        if len(shape) > 1:
            perm = tuple(reversed(range(len(shape))))
            views.append((X.transpose(perm), np.transpose(x, perm)))
        for V, v in views:
            for dims in power_set(tuple(range(v.ndim))):
                if not dims:
                    continue
                got = al.to_numpy(getattr(al, f"reduce_{op}")(V, dims=dims))
                want = orc.reduce(op, v, dims)
                assert_bit_equal(got, want, f"{op} {shape} dims={dims}")
                if ref is not None:
                    assert_bit_equal(got, ref.reduce(op, ref.from_numpy(np.ascontiguousarray(v)), dims).numpy(), f"{op} vs reference {dims}")
    # both zeros present in EVERY reduced set, so every output depends on the order
    z = np.where(rng.random((64, 48)) < 0.5, np.float32(0.0), np.float32(-0.0)).astype(np.float32)
    Z = al.from_numpy(z)
    for op in ("max", "min"):
        for dims in ((0,), (1,), (0, 1)):
            assert_bit_equal(al.to_numpy(getattr(al, f"reduce_{op}")(Z, dims=dims)), orc.reduce(op, z, dims), f"all-zero {op} {dims}")
    # sliding windows and a custom zero init (the init is the first element of the fold)
    base = rng.choice(np.array([-1.0, -0.0, 0.0], dtype=np.float32), size=5000).astype(np.float32)
    w = 16
    win = np.lib.stride_tricks.as_strided(base, shape=(base.size - w + 1, w), strides=(4, 4))
    got = al.to_numpy(al.from_numpy(base).as_strided(shape=win.shape, stride=(1, 1)).reduce_max(dims=[1]))
    assert_bit_equal(got, orc.reduce("max", win, (1,)), "window max zero signs")
    for init in (0.0, -0.0):
        got = al.to_numpy(al.reduce(max, Z, dims=(1,), init=init))
        assert_bit_equal(got, orc.reduce("max", z, (1,), init=init), f"max with init {init!r}")
