"""GPU parity for initialisers, views, copies and the golden vectors produced by the unmodified
reference (tests/golden/reference_vectors.npz). Everything here is bit-exact."""
import os

import numpy as np
import pytest

from tests.conftest import assert_bit_equal

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))
BINARY = ["sum", "sub", "mul", "div", "mod", "pow", "max", "min", "eq", "ne", "gt", "ge", "lt", "le"]
EXACT_UNARY = ["neg", "abs", "sqrt", "ceil", "floor"]
LIBM_UNARY = ["exp", "log", "sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "asinh", "acosh", "atanh"]


def test_config1_readme_pipeline(al):
    """BASELINE config 1, bit-exact against the reference's own output (README.md:15-28)."""
    import math
    a = al.arange(1, 1 + 3 * 4 * 5 * 6).reshape((3, 4, 5, 6)).reduce_sum(dims=(1, 2))
    assert a.shape == tuple(GOLD["c1_reduce_shape"]) and a.stride == tuple(GOLD["c1_reduce_stride"])
    assert_bit_equal(al.to_numpy(a), GOLD["c1_reduce"])
    a = a.apply(lambda x: math.sqrt(x))
    assert_bit_equal(al.to_numpy(a), GOLD["c1_sqrt"])
    a = a.reshape((3, 6)).transpose((1, 0))
    assert a.shape == tuple(GOLD["c1_a_shape"]) and a.stride == tuple(GOLD["c1_a_stride"])
    assert a.view == bool(GOLD["c1_a_view"])
    assert_bit_equal(al.to_numpy(a), GOLD["c1_a"])
    b = al.ones([3, 6]) + al.ones([6]) + al.ones([4, 3, 1]) + 1.0
    assert b.shape == (4, 3, 6) and b.stride == tuple(GOLD["c1_b_stride"])
    assert_bit_equal(al.to_numpy(b), GOLD["c1_b"])
    c = al.arange(1, 11).as_strided(shape=(8, 3), stride=(1, 1)).reduce_sum(dims=[0])
    assert c.shape == (1, 3)
    assert_bit_equal(al.to_numpy(c), GOLD["c1_c"])


@pytest.mark.parametrize("key", ["arange_sat", "arange_step3", "arange_step2"])
def test_arange_f32_recurrence(al, key):
    start, end, step = (int(v) for v in GOLD[key + "_args"])
    assert_bit_equal(al.to_numpy(al.arange(start, end, step)), GOLD[key], key)


def test_arange_large_prefix_and_saturated_tail(al, orc):
    n = (1 << 24) + 1000
    got = al.to_numpy(al.arange(0, n, 1))
    assert got[-1] == 16777216.0 and got[1 << 24] == 16777216.0 and got[(1 << 24) - 1] == 16777215.0
    assert_bit_equal(got, orc.arange(0, n, 1), "arange past 2^24")
    assert_bit_equal(al.to_numpy(al.arange(5, 3000000, 7)), orc.arange(5, 3000000, 7), "step 7")


def test_linspace_zeros_ones_array(al):
    assert_bit_equal(al.to_numpy(al.linspace(0, 10, 5)), GOLD["linspace_0_10_5"])
    assert_bit_equal(al.to_numpy(al.linspace(1, 100, 10)), GOLD["linspace_1_100_10"])
    assert_bit_equal(al.to_numpy(al.linspace(3, 7, 11)), GOLD["linspace_3_7_11"])
    assert_bit_equal(al.to_numpy(al.zeros((3, 5))), np.zeros((3, 5), np.float32))
    assert_bit_equal(al.to_numpy(al.ones((1, 1))), np.ones((1, 1), np.float32))
    assert_bit_equal(al.to_numpy(al.array([[1, 2, 3], [4, 5, 6]])), np.array([[1, 2, 3], [4, 5, 6]], np.float32))
    with pytest.raises(ValueError):
        al.zeros((0, 3))


@pytest.mark.parametrize("op", BINARY)
def test_binary_golden(al, op):
    X, Y = al.from_numpy(GOLD["bin_x"]), al.from_numpy(GOLD["bin_y"])
    name = {"sum": "add"}.get(op, op)
    assert_bit_equal(al.to_numpy(getattr(al, name)(X, Y)), GOLD[f"bin_{op}"], op)
    assert_bit_equal(al.to_numpy(getattr(al, name)(X, 0.75)), GOLD[f"sc_{op}"], op + " scalar")


@pytest.mark.parametrize("op", EXACT_UNARY)
def test_unary_golden_exact(al, op):
    assert_bit_equal(al.to_numpy(getattr(al, op)(al.from_numpy(GOLD["bin_x"]))), GOLD[f"un_{op}"], op)


@pytest.mark.parametrize("op", LIBM_UNARY)
def test_unary_golden_libm(al, op):
    got, want = al.to_numpy(getattr(al, op)(al.from_numpy(GOLD["bin_x"]))), GOLD[f"un_{op}"]
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = np.isfinite(want)
    assert np.array_equal(got[~ok & ~np.isnan(want)], want[~ok & ~np.isnan(want)])
    ulp = np.abs(got[ok].view(np.int32).astype(np.int64) - want[ok].view(np.int32).astype(np.int64))
    assert ulp.max(initial=0) <= 1


@pytest.mark.parametrize("mask", range(16))
def test_reduce_golden(al, mask):
    dims = tuple(d for d in range(4) if mask >> d & 1)
    Z = al.from_numpy(GOLD["red_z"])
    for op in ("sum", "max", "min"):
        assert_bit_equal(al.to_numpy(getattr(al, f"reduce_{op}")(Z, dims=dims)), GOLD[f"red_{op}_{mask}"], f"{op} {dims}")


def test_views_golden(al):
    t = al.arange(1, 361).reshape((3, 4, 5, 6)).transpose((1, 2, 0, 3))
    assert t.view and t.stride == tuple(GOLD["view_transpose_1203_stride"])
    assert_bit_equal(al.to_numpy(t), GOLD["view_transpose_1203"])
    w = al.arange(1, 41).as_strided(shape=(33, 8), stride=(1, 1))
    assert_bit_equal(al.to_numpy(w.reduce_sum(dims=[1])), GOLD["window_sum"])
    assert_bit_equal(al.to_numpy(w + w), GOLD["window_plus"])
    tt = al.arange(0, 24).reshape((4, 6)).transpose((1, 0))
    assert_bit_equal(al.to_numpy(tt + al.arange(0, 24).reshape((6, 4))), GOLD["t_plus"])
    r = tt.reshape((24,))
    assert not r.view  # a gathered copy, like the reference (array_deep_copy)
    assert_bit_equal(al.to_numpy(r), GOLD["t_reshape"])


def test_view_semantics(al, rng):
    x = rng.standard_normal((3, 4, 5, 6)).astype(np.float32)
    X = al.from_numpy(x)
    r = X.reshape((12, 30))
    assert r.view and r.stride == (30, 1) and r.size == 360
    assert_bit_equal(al.to_numpy(r), x.reshape(12, 30))
    m = al.move_dim(X, (1, 2), (2, 1))
    assert_bit_equal(al.to_numpy(m), np.moveaxis(x, (1, 2), (2, 1)))
    m = al.move_dim(X, (3, 0), (1, 2))
    assert_bit_equal(al.to_numpy(m), np.moveaxis(x, (3, 0), (1, 2)))
    assert_bit_equal(al.to_numpy(X.T), x.T)
    assert_bit_equal(al.to_numpy(X.ravel()), x.reshape(-1))
    # views share storage: writing through the base shows in the view
    v = X[1:3, 0:4, 2:5, 0:6:2]
    X[1, 0, 2, 0] = 99.0
    assert v[0, 0, 0, 0] == 99.0
    # transposing a NON-contiguous array gives the mathematically right answer (documented
    # deviation from the reference)
    xt = x.copy()
    xt[1, 0, 2, 0] = 99.0
    assert_bit_equal(al.to_numpy(X.T.T), xt)
    assert_bit_equal(al.to_numpy(X.T.transpose((1, 0, 3, 2))), np.transpose(xt.T, (1, 0, 3, 2)))
    with pytest.raises(ValueError):
        v.reshape((2 * 4 * 3 * 3,))  # element count differs from the base buffer


def test_setitem_cat_copy(al, rng):
    x = rng.standard_normal((4, 6)).astype(np.float32)
    y = rng.standard_normal((2, 3)).astype(np.float32)
    X, Y = al.from_numpy(x), al.from_numpy(y)
    X[1:3, 2:5] = Y
    x[1:3, 2:5] = y
    assert_bit_equal(al.to_numpy(X), x)
    X[0:4:2, 0:6:3] = 7.0
    x[0:4:2, 0:6:3] = 7.0
    assert_bit_equal(al.to_numpy(X), x)
    c = al.cat([X, X], [0])
    assert_bit_equal(al.to_numpy(c), np.concatenate([x, x], 0))
    c = al.cat([X, Y], [0, 1])
    want = np.zeros((6, 9), np.float32)
    want[:4, :6] = x
    want[4:, 6:] = y
    assert_bit_equal(al.to_numpy(c), want)
    s = al.copy(X)
    d = al.copy(X, deep=True)
    X[0, 0] = -5.0
    assert s[0, 0] == -5.0 and d[0, 0] == 7.0
    assert float(al.array([3.5])) == 3.5 and int(al.array([3.5])) == 3 and len(X) == 24
    assert str(al.array([[1, 2], [3, 4]])) == "[[1.0, 2.0], [3.0, 4.0]]"
    assert repr(X) == "NDArray(f32[4,6])"


def test_matmul_dot(al, orc, rng):
    a, b = rng.standard_normal((17, 40)).astype(np.float32), rng.standard_normal((40, 9)).astype(np.float32)
    assert_bit_equal(al.to_numpy(al.from_numpy(a) @ al.from_numpy(b)), orc.matmul(a, b), "matmul")
    assert_bit_equal(al.to_numpy(al.from_numpy(a) @ al.from_numpy(b.T.copy()).T), orc.matmul(a, b), "matmul strided")
    u, v = np.arange(1, 50, dtype=np.float32), np.arange(2, 51, dtype=np.float32)
    d = al.dot(al.from_numpy(u), al.from_numpy(v))
    assert d.shape == (1,) and float(d) == float(np.dot(u.astype(np.float64), v.astype(np.float64)))


def test_async_host_transfers(al, rng):
    """Pinned staging + the non-blocking upload / download streams give the same bits."""
    hin = al.host_buffer((3, 1000, 70))
    hin[...] = rng.standard_normal((3, 1000, 70)).astype(np.float32)
    hout = al.host_buffer((3, 1000, 70))
    outs = []
    for k in range(4):  # several transfers in flight at once, blocks recycled underneath
        X = al.from_numpy(hin, blocking=False)
        Y = X * float(k + 1) + 0.5
        o = al.host_buffer((3, 1000, 70))
        al.to_numpy(Y, out=o, blocking=False)
        del X, Y
        outs.append(o)
    T = al.from_numpy(hin, blocking=False).transpose((2, 0, 1))
    al.to_numpy(T, out=hout.reshape(70, 3, 1000), blocking=False)  # non-contiguous source: gathered first
    al.synchronize()
    for k, o in enumerate(outs):
        assert_bit_equal(o, hin * np.float32(k + 1) + np.float32(0.5), f"async {k}")
    assert_bit_equal(hout.reshape(70, 3, 1000), np.transpose(hin, (2, 0, 1)), "async transposed")


def test_uploads_are_awaited_per_array_at_first_use(al, rng):
    """An async upload does not stall the compute stream; the first op that touches the array waits for that
    array's own upload. Big upload first, small ones after it, ops on the small ones, then on the big one;
    uploads that are never used (freed at once) must not let the recycled block be overwritten early."""
    big = al.host_buffer((1 << 26,))  # 256 MiB: still on the bus while the small operands are used
    big[:] = rng.standard_normal(big.shape).astype(np.float32)
    sa, sb = al.host_buffer((100003,)), al.host_buffer((100003,))
    sa[:], sb[:] = rng.standard_normal(sa.shape).astype(np.float32), rng.standard_normal(sb.shape).astype(np.float32)
    for rep in range(3):
        X = al.from_numpy(big, blocking=False)
        A, B = al.from_numpy(sa, blocking=False), al.from_numpy(sb, blocking=False)
        small = A * B + 1.0
        first = al.to_numpy(small)                       # blocking download while X may still be in flight
        assert_bit_equal(first, sa * sb + np.float32(1.0), "small operands")
        view = X[1:1 + (1 << 20)]                        # a view shares the upload of its base
        assert_bit_equal(al.to_numpy(view + 2.0), big[1:1 + (1 << 20)] + np.float32(2.0), "view of a pending upload")
        assert float(X[12345]) == float(big[12345])
        unused = al.from_numpy(big, blocking=False)      # never used: freed while its upload is pending ...
        del unused
        Z = al.zeros((1 << 26,)) + 3.0                   # ... its block is recycled here and must end up all 3.0
        assert_bit_equal(al.to_numpy(al.reduce_max(Z)), np.float32([3.0]), "recycled block")
        assert_bit_equal(al.to_numpy(al.reduce_min(Z)), np.float32([3.0]), "recycled block")
        del X, A, B, Z
    al.synchronize()


def test_device_out_of_memory_raises_and_recovers(al):
    """A request beyond the 180 GB of HBM raises MemoryError (the reference would abort the process);
    the cache is handed back to the driver on the way and the library keeps working."""
    from arraylib_b200._lib import lib
    keep = al.ones((1 << 20,))
    with pytest.raises(MemoryError):
        al.zeros((1 << 36,))  # 256 GiB
    assert lib.al_last_error().startswith(b"MemoryError")
    assert float(al.to_numpy(al.reduce_sum(keep + 1.0)).reshape(-1)[0]) == 2.0 * (1 << 20)
    assert lib.al_trim_pool() == 0
    assert float(al.to_numpy(al.reduce_sum(keep)).reshape(-1)[0]) == float(1 << 20)


def test_arange_random_parameters_through_the_c_abi(al, orc):
    """array_arange takes f32 arguments; the Python API only passes ints. Integer and non-integer
    steps, starts near 2^24 (where the reference's f32 recurrence stops being exact), all bit-exact."""
    from arraylib_b200._lib import lib
    rng = np.random.default_rng(99)
    cases = [(0, 7, 1), (3, 50, 2.5), (1, 9, 1.75), (16777210, 16777300, 1), (16777000, 16779000, 3), (5, 40000, 7.25),
             (33554400, 33554600, 2), (33554400, 33554600, 5), (0.5, 20.5, 1), (2.75, 100, 3)]
    for _ in range(30):
        start = float(rng.integers(0, 1 << 25))
        step = float(rng.choice([1, 2, 3, 5, 1.5, 2.25, 7, 16, 33]))
        cases.append((start, start + float(rng.integers(1, 5000)), step))
    for start, end, step in cases:
        want = orc.arange(start, end, step)
        got = al.NDArray(lib.array_arange(start, end, step))
        assert got.shape == want.shape, (start, end, step)
        assert_bit_equal(al.to_numpy(got), want, f"arange({start}, {end}, {step})")
    for start, end, n in [(0, 10, 5), (1, 100, 10), (3, 7, 11), (0, 1, 1000), (5, 6, 3)]:
        assert_bit_equal(al.to_numpy(al.linspace(start, end, n)), orc.linspace(start, end, n), f"linspace {start} {end} {n}")


def test_concurrent_python_threads(al, rng):
    """ctypes drops the GIL around every call; the library serialises on its own lock and recycles
    device blocks in stream order, so independent threads can interleave ops and frees freely."""
    import threading
    data = [rng.standard_normal((257, 130)).astype(np.float32) for _ in range(6)]
    results, errors = {}, []

    def work(tid):
        try:
            x = data[tid]
            X = al.from_numpy(x)
            acc = X
            for k in range(25):
                acc = (acc + X) * 0.5 - float(k % 3)
                if k % 5 == 0:
                    tmp = al.copy(acc.transpose((1, 0)), deep=True)  # allocate / free churn
                    del tmp
            results[tid] = (al.to_numpy(acc), al.to_numpy(al.reduce_max(X, dims=(0,))))
        except Exception as exc:  # pragma: no cover
            errors.append(exc)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(data))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for tid, x in enumerate(data):
        acc = x.copy()
        for k in range(25):
            acc = (acc + x) * np.float32(0.5) - np.float32(k % 3)
        assert_bit_equal(results[tid][0], acc, f"thread {tid}")
        assert_bit_equal(results[tid][1], x.max(axis=0, keepdims=True), f"thread {tid} max")
