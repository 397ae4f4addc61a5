"""`import arraylib as al` on N GPUs (arraylib_b200.spmd): the reference's API on GLOBAL arrays whose blocks live on
different GPUs, checked against NumPy on the whole arrays (integer-valued data: every result is exact, so the
comparison is bit for bit even where the summation order differs).

    ARRAYLIB_DEVICES=2 python tools/spmd_check.py          # the import re-launches this script once per GPU
    python -m arraylib_b200.launch --nproc 2 --spmd tools/spmd_check.py

Test infrastructure, not product code."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("ARRAYLIB_SHARD_MIN_BYTES", "4096")
import numpy as np  # noqa: E402

import arraylib as al  # noqa: E402

from arraylib_b200 import dist  # noqa: E402

assert getattr(al, "enabled", lambda: False)(), "not in SPMD mode: set ARRAYLIB_DEVICES=N or launch with --spmd"
rank, world = dist.rank(), dist.world()
rng = np.random.default_rng(5)
fails = []


def check(name, got, want, placement=None):
    host = al.to_numpy(got) if not isinstance(got, (float, np.ndarray)) else np.asarray(got, dtype=np.float32)
    want = np.ascontiguousarray(want, dtype=np.float32)
    ok = host.shape == want.shape and np.array_equal(host.view(np.uint32), want.view(np.uint32))
    if placement is not None and getattr(got, "placement", placement) != placement:
        ok = False
        name += f" (placement {got.placement}, expected {placement})"
    if not ok:
        fails.append(name)
    if rank == 0:
        print(f"{name:58s} {'ok' if ok else 'MISMATCH'}", flush=True)


L = 16 * world
a_np = rng.integers(-8, 9, size=(L, 1, 128)).astype(np.float32)
b_np = rng.integers(-8, 9, size=(1, 24, 128)).astype(np.float32)
c_np = rng.integers(-8, 9, size=(128,)).astype(np.float32)
A, B, Cv = al.from_numpy(a_np), al.from_numpy(b_np), al.from_numpy(c_np)
check("from_numpy round trip", A, a_np, "sharded")
u = A + B + Cv + 1.0
check("A + B + c + 1.0", u, a_np + b_np + c_np + 1.0, "sharded")
check("scalar on the left / comparisons", (2.0 - A) * (A > 0.0), (2.0 - a_np) * (a_np > 0.0), "sharded")
check("a % 3 (fmod)", al.mod(A, 3.0), np.fmod(a_np, 3.0), "sharded")
check("apply(sqrt) o abs", al.apply(np.sqrt, al.abs(A) if hasattr(al, "abs") else A.apply(abs)), np.sqrt(np.abs(a_np)), "sharded")

x_np = rng.integers(-8, 9, size=(L, 12, 10, 32)).astype(np.float32)
X = al.from_numpy(x_np)
for dims in [(1, 2), (3,), (0,), (0, 2), (0, 1, 2, 3), (1,), (0, 3)]:
    check(f"reduce_sum dims={dims}", X.reduce_sum(dims=dims), x_np.sum(axis=dims, keepdims=True, dtype=np.float64))
    check(f"reduce_max dims={dims}", al.reduce_max(X, dims=dims), x_np.max(axis=dims, keepdims=True))
check("reduce(lambda acc, v: acc + v, dims=(0, 1))", al.reduce(lambda acc, v: acc + v, X, dims=(0, 1), init=0), x_np.sum(axis=(0, 1), keepdims=True))
check("reduce_sum keeps the shard axis -> stays sharded", X.reduce_sum(dims=(1, 2)), x_np.sum(axis=(1, 2), keepdims=True), "sharded")

# creation
n = 1 << 16
check("arange(n) sharded", al.arange(n), np.arange(n), "sharded")
check("arange(3, n, 7)", al.arange(3, n, 7), np.arange(3, n, 7))
check("ones / zeros / full", al.ones((L, 300)) + al.zeros((L, 300)) * 2.0, np.ones((L, 300)), "sharded")
check("arange().reshape (README pipeline)", al.arange(1, 361).reshape((3, 4, 5, 6)).reduce_sum(dims=(1, 2)).apply(np.sqrt).transpose((3, 2, 1, 0)),
      np.sqrt(np.arange(1, 361, dtype=np.float32).reshape(3, 4, 5, 6).sum(axis=(1, 2), keepdims=True)).transpose(3, 2, 1, 0))

# views
base_np = rng.integers(-8, 9, size=(20000,)).astype(np.float32)
base = al.from_numpy(base_np)
w = 64
win = base.as_strided(shape=(base_np.size - w + 1, w), stride=(1, 1))
check("sliding-window reduce_sum (halo exchange)", win.reduce_sum(dims=[1]),
      np.lib.stride_tricks.sliding_window_view(base_np, w).sum(axis=1, keepdims=True, dtype=np.float64), "sharded")
m_np = rng.integers(-8, 9, size=(64 * world, 96)).astype(np.float32)
y_np = rng.integers(-8, 9, size=(96, 64 * world)).astype(np.float32)
M, Y = al.from_numpy(m_np), al.from_numpy(y_np)
check("M.transpose((1,0)) + Y (all-to-all re-shard)", M.transpose((1, 0)) + Y, m_np.T + y_np, "sharded")
check("M.T.reshape(-1)", al.reshape(M.T, (m_np.size,)), m_np.T.reshape(-1))
check("reshape keeps blocks", al.reshape(M, (64 * world, 8, 12)), m_np.reshape(64 * world, 8, 12), "sharded")
check("X[:, 2:9:3, :, 4:20] (shard axis whole)", X[0:L, 2:9:3, 0:10, 4:20], x_np[:, 2:9:3, :, 4:20], "sharded")
check("X[3:L-2:2, ...] (cuts the shard axis)", X[3:L - 2:2, 0:12, 0:10, 0:32], x_np[3:L - 2:2])
check("X[i, j, k, l] on every owner", np.array([X[i, 3, 4, 5] for i in range(0, L, 5)], dtype=np.float32), x_np[0:L:5, 3, 4, 5])

# in-place writes
z_np = np.zeros((L, 40), dtype=np.float32)
Z = al.zeros((L, 40))
Z[1:L - 1:3, 2:30:4] = 7.0
z_np[1:L - 1:3, 2:30:4] = 7.0
check("Z[1:L-1:3, 2:30:4] = 7.0", Z, z_np, "sharded")
v_np = rng.integers(-8, 9, size=z_np[2:L:5, 0:40:2].shape).astype(np.float32)
Z[2:L:5, 0:40:2] = al.from_numpy(v_np)
z_np[2:L:5, 0:40:2] = v_np
check("Z[2:L:5, ::2] = V", Z, z_np, "sharded")
for i in (0, L // 2, L - 1):
    Z[i, 7] = float(i)
    z_np[i, 7] = float(i)
check("Z[i, 7] = i", Z, z_np, "sharded")

# gathered fall-backs
p_np, q_np = rng.integers(-3, 4, size=(L, 20)).astype(np.float32), rng.integers(-3, 4, size=(20, 6)).astype(np.float32)
check("matmul", al.matmul(al.from_numpy(p_np), al.from_numpy(q_np)), p_np @ q_np)
check("where(cond, X, 0.0)", al.where(X > 1.0, X, 0.0), np.where(x_np > 1.0, x_np, 0.0), "sharded")
check("cat along the shard axis", al.cat([al.from_numpy(p_np), al.from_numpy(p_np + 1)], [0]), np.concatenate([p_np, p_np + 1], axis=0))
check("[N,1] + [N] falls back to a replicated result", al.from_numpy(p_np[:, :1].copy()) + al.from_numpy(p_np[:, 0].copy()), p_np[:, :1] + p_np[:, 0])
assert repr(X) == f"NDArray(f32[{L},12,10,32])" and X.shape == x_np.shape and X.ndim == 4 and len(X) == x_np.size

total = int(dist.allreduce_scalar(float(len(fails)), "sum"))
dist.destroy()
if rank == 0:
    print("SPMD_CHECK", "PASS" if total == 0 else f"FAIL {fails}", f"({world} ranks)")
sys.exit(0 if total == 0 else 1)
