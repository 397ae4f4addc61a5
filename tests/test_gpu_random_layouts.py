"""Randomised differential tests: random base buffers viewed through random layouts (permutations,
slices with steps, broadcast axes, overlapping as_strided windows) on BOTH sides -- numpy views for
the CPU oracle, al views for the GPU -- then binary ops, copies, unary maps and reductions over random
dim sets. Seeds are fixed, so failures reproduce. Everything except float reduce_sum is bit-exact."""
import numpy as np
import pytest

from tests.conftest import assert_bit_equal

pytestmark = pytest.mark.gpu

BINARY = ["sum", "sub", "mul", "div", "max", "min", "lt", "ge", "ne", "mod"]
AL_NAME = {"sum": "add"}


def random_view(al, rng, max_ndim=4, max_extent=9, ints=False):
    """Returns (al_view, numpy_view) of the same random layout over the same random base."""
    kind = rng.choice(["contig", "perm", "slice", "strided", "perm_slice"])
    ndim = int(rng.integers(1, max_ndim + 1))
    shape = tuple(int(v) for v in rng.integers(1, max_extent + 1, ndim))
    if rng.random() < 0.3:  # sprinkle some larger axes so vector / tile paths are reached
        shape = tuple(int(s * rng.choice([1, 4, 8])) for s in shape)
    n = int(np.prod(shape))
    base = (rng.integers(-8, 9, n) if ints else rng.standard_normal(n)).astype(np.float32)
    A = al.from_numpy(base).reshape(shape)
    a = base.reshape(shape)
    if kind in ("perm", "perm_slice") and ndim > 1:
        perm = tuple(int(p) for p in rng.permutation(ndim))
        A, a = A.transpose(perm), np.transpose(a, perm)
    if kind in ("slice", "perm_slice"):
        sl = []
        for e in a.shape:
            lo = int(rng.integers(0, e))
            hi = int(rng.integers(lo + 1, e + 1))
            st = int(rng.integers(1, 4))
            sl.append(slice(lo, hi, st))
        A, a = A[tuple(sl)], a[tuple(sl)]
    if kind == "strided":
        # arbitrary (possibly overlapping) non-negative strides that stay inside the base buffer
        nd2 = int(rng.integers(1, 4))
        shp = tuple(int(v) for v in rng.integers(1, 7, nd2))
        strides = tuple(int(v) for v in rng.integers(1, 5, nd2))
        reach = sum((s - 1) * t for s, t in zip(shp, strides))
        if reach < n:
            flat = al.from_numpy(base)
            A = flat.as_strided(shape=shp, stride=strides)
            a = np.lib.stride_tricks.as_strided(base, shape=shp, strides=tuple(4 * t for t in strides))
    return A, a


def broadcastable_partner(al, rng, shape, ints=False):
    """A random array whose shape broadcasts against `shape` (axes dropped or set to 1)."""
    nd = len(shape)
    keep = int(rng.integers(0, nd + 1))
    sub = list(shape[nd - keep:]) if keep else [1]
    sub = [e if rng.random() < 0.6 else 1 for e in sub]
    if rng.random() < 0.2:
        sub = [int(rng.integers(1, 4))] * 0 + sub  # same rank
    n = int(np.prod(sub))
    b = (rng.integers(-8, 9, n) if ints else rng.standard_normal(n)).astype(np.float32).reshape(sub)
    return al.from_numpy(b), b


@pytest.mark.parametrize("seed", range(40))
def test_random_binary_ops_over_random_views(al, orc, seed):
    rng = np.random.default_rng(1000 + seed)
    for _ in range(8):
        A, a = random_view(al, rng)
        B, b = broadcastable_partner(al, rng, a.shape)
        op = str(rng.choice(BINARY))
        fn = getattr(al, AL_NAME.get(op, op))
        if rng.random() < 0.5:
            got, want = fn(A, B), orc.binary(op, a, b)
        else:
            got, want = fn(B, A), orc.binary(op, b, a)
        assert got.shape == want.shape
        assert_bit_equal(al.to_numpy(got), want, f"seed {seed} {op} {a.shape} {a.strides} x {b.shape}")
        s = float(rng.integers(-3, 4))
        assert_bit_equal(al.to_numpy(fn(A, s)), orc.binary(op, a, s), f"seed {seed} {op} scalar")


@pytest.mark.parametrize("seed", range(25))
def test_random_copies_and_unary_over_random_views(al, orc, seed):
    rng = np.random.default_rng(2000 + seed)
    for _ in range(8):
        A, a = random_view(al, rng, max_ndim=5, max_extent=8)
        assert_bit_equal(al.to_numpy(A), a, f"seed {seed} to_numpy {a.shape} {a.strides}")
        assert_bit_equal(al.to_numpy(al.copy(A, deep=True)), orc.gather(a), f"seed {seed} deep_copy")
        assert_bit_equal(al.to_numpy(A.ravel()), a.reshape(-1), f"seed {seed} ravel")
        assert_bit_equal(al.to_numpy(al.sqrt(al.abs(A))), orc.unary("sqrt", orc.unary("abs", a)), f"seed {seed} unary")
        C, c = random_view(al, rng, max_ndim=3)
        L = al.from_numpy(np.ascontiguousarray(c * 2))
        assert_bit_equal(al.to_numpy(al.where(al.gt(C, 0.0), C, L)), orc.where(orc.binary("gt", c, 0.0), c, np.ascontiguousarray(c * 2)), "where")


@pytest.mark.parametrize("seed", range(40))
def test_random_reductions_over_random_views(al, orc, seed):
    rng = np.random.default_rng(3000 + seed)
    for _ in range(6):
        A, a = random_view(al, rng, max_ndim=4, max_extent=12, ints=True)
        nd = a.ndim
        dims = tuple(int(d) for d in np.nonzero(rng.random(nd) < 0.5)[0])
        for op in ("sum", "max", "min"):
            got = getattr(al, f"reduce_{op}")(A, dims=dims)
            want = orc.reduce(op, a, dims)
            assert got.shape == want.shape
            assert_bit_equal(al.to_numpy(got), want, f"seed {seed} reduce_{op} {a.shape} {a.strides} dims {dims}")
        full = al.reduce_sum(A)
        assert_bit_equal(al.to_numpy(full), orc.reduce("sum", a, None), f"seed {seed} full reduce")


@pytest.mark.parametrize("seed", range(10))
def test_random_float_sums_within_tolerance(al, orc, seed):
    rng = np.random.default_rng(4000 + seed)
    for _ in range(6):
        A, a = random_view(al, rng, max_ndim=4, max_extent=16)
        dims = tuple(int(d) for d in np.nonzero(rng.random(a.ndim) < 0.6)[0])
        got = al.to_numpy(al.reduce_sum(A, dims=dims)).astype(np.float64)
        want = orc.reduce("sum", a, dims).astype(np.float64)
        n = int(np.prod([a.shape[d] for d in dims])) if dims else 1
        tol = 2.0 * n * 2.0 ** -24 * np.sum(np.abs(a.astype(np.float64)), axis=dims, keepdims=True) if dims else np.zeros_like(want)
        assert np.all(np.abs(got - want) <= tol + 0.0), f"seed {seed} {a.shape} {dims}"


@pytest.mark.parametrize("seed", [11, 12])
def test_short_fuzz_campaign_with_large_lopsided_shapes(seed):
    """tools/fuzz.py for a few seconds: the same differential idea with shapes up to 2M elements (tiny odd inner
    extents under long outer axes, operands broadcast in the middle, misaligned bases, window widths 8..1024)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz.py"), "8", str(seed)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "fuzz ok" in res.stdout
