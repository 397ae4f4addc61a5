"""Time-bounded differential fuzzing of the GPU path against the CPU oracle (development aid; the fixed-seed
subset that runs in CI is tests/test_gpu_random_layouts.py). Larger and more lopsided shapes than the tests use:
tiny odd inner extents under long outer axes, operands broadcast at both ends or in the middle, misaligned
bases, ragged rows under reductions, arbitrary window widths.
    python tools/fuzz.py [seconds] [seed]
Exit code 1 on the first mismatch, after printing the case."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import arraylib as al  # noqa: E402
from arraylib_b200._lib import lib  # noqa: E402
from oracle import oracle as orc  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(seed)
BINARY = ["sum", "sub", "mul", "div", "max", "min", "lt", "ge", "ne", "mod"]
counts = {}


def bits_equal(got, want):
    want = np.ascontiguousarray(want, dtype=np.float32)
    if got.shape != want.shape:
        return False
    return bool(np.array_equal(got.view(np.uint32), want.view(np.uint32)))  # strict: NaN sign and payload included


def fail(kind, detail):
    print(f"MISMATCH [{kind}] {detail} (seed {seed}, last kernel {lib.al_last_kernel().decode()})", flush=True)
    sys.exit(1)


def random_shape(max_total):
    nd = int(rng.integers(1, 6))
    pool = [1, 2, 3, 3, 4, 5, 7, 8, 9, 12, 16, 17, 31, 32, 33, 63, 64, 100, 255, 256, 1000, 4096]
    shape = [int(rng.choice(pool)) for _ in range(nd)]
    if rng.random() < 0.5:
        shape[int(rng.integers(0, nd))] = int(rng.integers(1, 200000))
    while int(np.prod(shape)) > max_total:
        shape[int(np.argmax(shape))] = max(1, shape[int(np.argmax(shape))] // 2)
    return tuple(shape)


def operand(out_shape, ints):
    """An operand that broadcasts to out_shape: random axes collapsed to 1, leading axes possibly dropped, and
    the base possibly shifted by 1-3 elements (misaligned) or permuted."""
    nd = len(out_shape)
    shp = [e if rng.random() < 0.65 else 1 for e in out_shape]
    drop = int(rng.integers(0, nd)) if rng.random() < 0.4 else 0
    shp = shp[drop:] or [1]
    n = int(np.prod(shp))
    data = (rng.integers(-8, 9, n) if ints else rng.standard_normal(n)).astype(np.float32)
    how = rng.choice(["plain", "shift", "perm"])
    if how == "shift":
        k = int(rng.integers(1, 4))
        base = np.zeros(n + k, np.float32)
        base[k:] = data
        strides = tuple(int(np.prod(shp[i + 1:])) for i in range(len(shp)))
        A = al.from_numpy(base)[k:n + k].as_strided(shape=tuple(shp), stride=strides)
        return A, base[k:].reshape(shp)
    if how == "perm" and len(shp) > 1:
        perm = tuple(int(p) for p in rng.permutation(len(shp)))
        inv = tuple(int(i) for i in np.argsort(perm))
        stored = np.ascontiguousarray(np.transpose(data.reshape(shp), perm))
        return al.from_numpy(stored).transpose(inv), np.transpose(stored, inv)
    return al.from_numpy(data.reshape(shp)), data.reshape(shp)


def case_binary():
    out_shape = random_shape(1 << 21)
    ints = rng.random() < 0.3
    A, a = operand(out_shape, ints)
    B, b = operand(out_shape, ints)
    op = str(rng.choice(BINARY))
    fn = getattr(al, "add" if op == "sum" else op)
    got = al.to_numpy(fn(A, B))
    if not bits_equal(got, orc.binary(op, a, b)):
        fail("binary", f"{op} {a.shape}{a.strides} x {b.shape}{b.strides}")
    if rng.random() < 0.3:
        s = float(rng.integers(-3, 4)) + 0.5
        if not bits_equal(al.to_numpy(fn(A, s)), orc.binary(op, a, s)):
            fail("scalar", f"{op} {a.shape}{a.strides} x {s}")
    if rng.random() < 0.3:
        if not bits_equal(al.to_numpy(al.copy(A, deep=True)), orc.gather(a)):
            fail("copy", f"{a.shape}{a.strides}")


def case_reduce():
    shape = random_shape(1 << 21)
    x = rng.integers(-8, 9, int(np.prod(shape))).astype(np.float32).reshape(shape)
    X = al.from_numpy(x)
    xv = x
    if rng.random() < 0.4 and shape[-1] > 4:  # ragged rows: slice the inner axis
        lo = int(rng.integers(0, 4))
        hi = shape[-1] - int(rng.integers(0, min(4, shape[-1] - lo - 1) + 1))
        sl = tuple([slice(0, e) for e in shape[:-1]] + [slice(lo, hi)])
        X, xv = X[sl], x[sl]
    if rng.random() < 0.3 and len(shape) > 1:
        perm = tuple(int(p) for p in rng.permutation(len(shape)))
        X, xv = X.transpose(perm), np.transpose(xv, perm)
    nd = xv.ndim
    dims = tuple(int(d) for d in np.nonzero(rng.random(nd) < 0.5)[0]) or (int(rng.integers(0, nd)),)
    op = str(rng.choice(["sum", "max", "min"]))
    got = al.to_numpy(getattr(al, f"reduce_{op}")(X, dims=dims))
    if not bits_equal(got, orc.reduce(op, xv, dims)):
        fail("reduce", f"{op} {xv.shape}{xv.strides} dims={dims}")


def case_window():
    w = int(rng.choice([8, 9, 16, 24, 31, 32, 33, 64, 65, 100, 127, 128, 500, 1000, 1024]))
    rows = int(rng.choice([1, 1, 1, 2, 5]))
    n = int(rng.integers(4 * w + 1, 400000))
    base = rng.integers(-8, 9, (rows, n)).astype(np.float32)
    if rows == 1:
        V = al.from_numpy(base.reshape(-1)).as_strided(shape=(n - w + 1, w), stride=(1, 1))
        vn = np.lib.stride_tricks.as_strided(base.reshape(-1), shape=(n - w + 1, w), strides=(4, 4))
        dims = (1,)
    else:
        V = al.from_numpy(base.reshape(-1)).as_strided(shape=(rows, n - w + 1, w), stride=(n, 1, 1))
        vn = np.lib.stride_tricks.as_strided(base, shape=(rows, n - w + 1, w), strides=(4 * n, 4, 4))
        dims = (2,)
    op = str(rng.choice(["sum", "max", "min"]))
    got = al.to_numpy(getattr(al, f"reduce_{op}")(V, dims=dims))
    if not bits_equal(got, orc.reduce(op, vn, dims)):
        fail("window", f"{op} rows={rows} n={n} w={w}")


cases = [(case_binary, 0.6), (case_reduce, 0.3), (case_window, 0.1)]
t0 = time.time()
done = 0
while time.time() - t0 < budget:
    r = rng.random()
    acc = 0.0
    for fn, p in cases:
        acc += p
        if r < acc:
            fn()
            k = lib.al_last_kernel().decode()
            counts[k] = counts.get(k, 0) + 1
            break
    done += 1
print(f"fuzz ok: {done} cases in {time.time() - t0:.0f} s (seed {seed}); last-kernel histogram: {dict(sorted(counts.items(), key=lambda kv: -kv[1]))}")
