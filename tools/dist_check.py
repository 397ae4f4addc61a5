"""Multi-GPU parity check, one rank per GPU, no PyTorch:
    python -m arraylib_b200.launch --nproc N tools/dist_check.py          (torchrun works as well)
Every rank builds the same global host arrays, uploads only its leading-axis shard (plus replicated
small operands / window halo), runs the sharded hot path on its GPU and the results are gathered
and compared with the CPU oracle on the whole arrays. Test infrastructure, not product code."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import arraylib as al  # noqa: E402
from arraylib_b200 import dist  # noqa: E402
from arraylib_b200._lib import lib  # noqa: E402
from oracle import oracle as orc  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
assert lib.al_init(local) == 0
dist.init(rank, world)  # NCCL id through the file rendezvous
rng = np.random.default_rng(11)
L = 8 * world
A = rng.standard_normal((L, 1, 256)).astype(np.float32)
B = rng.standard_normal((1, 24, 256)).astype(np.float32)
c = rng.standard_normal((256,)).astype(np.float32)
X = rng.integers(-8, 9, size=(L, 20, 12, 64)).astype(np.float32)
base = rng.integers(-8, 9, size=(40000,)).astype(np.float32)

fails = []


def check(name, got, want):
    ok = np.array_equal(got.view(np.uint32), np.ascontiguousarray(want, dtype=np.float32).view(np.uint32))
    if not ok:
        fails.append(name)
    if rank == 0:
        print(f"{name:28s} {'ok' if ok else 'MISMATCH'}", flush=True)


# broadcast chain: sharded A, replicated B and c; output stays sharded -> allgather to compare
u = dist.scatter_from_numpy(A) + al.from_numpy(B) + al.from_numpy(c) + 1.0
check("bcast chain (allgather)", al.to_numpy(dist.allgather(u)), orc.binary("sum", orc.binary("sum", orc.binary("sum", A, B), c), 1.0))
Xs = dist.scatter_from_numpy(X)
check("reduce dims=(1,2) local", al.to_numpy(dist.allgather(dist.reduce_sharded(Xs, "sum", (1, 2)))), orc.reduce("sum", X, (1, 2)))
check("reduce dims=(3,) local", al.to_numpy(dist.allgather(dist.reduce_sharded(Xs, "sum", (3,)))), orc.reduce("sum", X, (3,)))
check("reduce dims=(0,) allreduce", al.to_numpy(dist.reduce_sharded(Xs, "sum", (0,))), orc.reduce("sum", X, (0,)))
check("reduce all allreduce", al.to_numpy(dist.reduce_sharded(Xs, "sum", None)), orc.reduce("sum", X, None))
check("reduce_max dims=(0,2)", al.to_numpy(dist.reduce_sharded(Xs, "max", (0, 2))), orc.reduce("max", X, (0, 2)))
w = 64
shard, n_local = dist.window_shard_from_numpy(base, w)
win = shard.as_strided(shape=(n_local, w), stride=(1, 1)).reduce_sum(dims=[1])
n_out = base.size - w + 1
wn = np.lib.stride_tricks.as_strided(base, shape=(n_out, w), strides=(4, 4))
want = orc.reduce("sum", wn, (1,))
lo, hi = dist.shard_bounds(n_out, rank, world)
check(f"window halo shard (rank {rank})", al.to_numpy(win), want[lo:hi])
# the same through the ShardedArray front end
SA, SB, Sc, SX = dist.ShardedArray.scatter(A), dist.ShardedArray.replicate(B), dist.ShardedArray.replicate(c), dist.ShardedArray.scatter(X)
check("ShardedArray chain", (SA + SB + Sc + 1.0).to_numpy(), orc.binary("sum", orc.binary("sum", orc.binary("sum", A, B), c), 1.0))
check("ShardedArray sharded*replica", (SX * dist.ShardedArray.replicate(X)).to_numpy(), orc.binary("mul", X, X))
r = SX.reduce_sum(dims=(0, 2))  # 20 x 64 outputs: stays sharded along axis 1 when 20 divides by the world size
assert r.shape == (1, 20, 1, 64) and r.placement == ("sharded" if 20 % world == 0 else "replicated")
check("ShardedArray reduce (0,2)", r.to_numpy(), orc.reduce("sum", X, (0, 2)))
r = SX.reduce_sum(dims=(0, 2), replicated=True)
assert r.placement == "replicated" and r.shape == (1, 20, 1, 64)
check("ShardedArray reduce (0,2) replicated", r.to_numpy(), orc.reduce("sum", X, (0, 2)))
r = (SX - 1.0).reduce_max(dims=(1, 3))
assert r.placement == "sharded"
check("ShardedArray reduce_max (1,3)", r.to_numpy(), orc.reduce("max", orc.binary("sub", X, 1.0), (1, 3)))

# ---- re-sharding: transpose of a sharded array (local / one all-to-all / gather fallback), ragged gather ----
for shape, perm in [((8 * world, 16 * world), (1, 0)), ((4 * world, 6, 8 * world), (2, 0, 1)), ((4 * world, 6, 8 * world), (2, 1, 0)),
                    ((4 * world, 8 * world, 5), (1, 2, 0)), ((4 * world, 6, 10), (0, 2, 1)), ((4 * world + 1, 8 * world), (1, 0)),
                    ((4 * world, 8 * world + 3), (1, 0))]:
    G = rng.standard_normal(shape).astype(np.float32)
    T = dist.ShardedArray.scatter(G).transpose(perm)
    how = dist.reshard_plan(shape, perm, world)["how"]
    lo, hi = dist.shard_bounds(T.shape[0], rank, world)
    check(f"transpose {shape} {perm} [{how}] local", al.to_numpy(T.local), np.transpose(G, perm)[lo:hi])
    check(f"transpose {shape} {perm} [{how}] gathered", T.to_numpy(), np.transpose(G, perm))
# config C5b with BOTH operands globally sharded: X.T + Y needs the column blocks of X (one all-to-all)
m = 64 * world
Xg, Yg = rng.standard_normal((m, m)).astype(np.float32), rng.standard_normal((m, m)).astype(np.float32)
check("sharded X.T + Y", (dist.ShardedArray.scatter(Xg).transpose((1, 0)) + dist.ShardedArray.scatter(Yg)).to_numpy(), orc.binary("sum", Xg.T, Yg))

# ---- the cross-GPU reduce itself: peer-memory windows (default) against the oracle, against NCCL, across ranks ----
import ctypes as C  # noqa: E402
import hashlib  # noqa: E402

if rank == 0:
    print("transport before first use:", dist.transport(), flush=True)


def same_on_all_ranks(name, host):
    digest = hashlib.sha1(np.ascontiguousarray(host).tobytes()).digest()[:12]
    mine = np.frombuffer(digest, dtype=np.uint16).astype(np.float32)  # 6 exact small integers; padded to 8 for the wire
    mine = np.concatenate([mine, np.zeros(2, np.float32)])
    lo, hi = al.from_numpy(mine.copy()), al.from_numpy(mine.copy())
    dist.allreduce(lo, "min"), dist.allreduce(hi, "max")
    ok = bool(np.array_equal(al.to_numpy(lo), al.to_numpy(hi)))
    if not ok:
        fails.append(name + " (ranks differ)")
    if rank == 0 and not ok:
        print(f"{name:28s} RANKS DIFFER", flush=True)


for shape, dims, op in [((L, 7, 13), (0,), "sum"), ((L, 7, 13), (0,), "max"), ((L, 7, 13), (0,), "min"), ((L, 5, 64), (0, 1), "sum"),
                        ((L, 33), (0,), "sum"), ((L, 4096), (0,), "sum"), ((L, 3, 1000), (0, 2), "sum"), ((L, 64), (0, 1), "sum"),
                        ((L,), (0,), "sum"), ((L, 6, 40), (0, 1, 2), "max"), ((L, 1, 260), (0,), "min")]:
    G = rng.integers(-8, 9, size=shape).astype(np.float32)
    Gs = dist.scatter_from_numpy(G)
    for rep in range(3):  # repeated calls: the window epochs advance, the slots are reused
        got = al.to_numpy(dist.reduce_sharded(Gs, op, dims))
        check(f"peer {op} {shape} {dims} #{rep}", got, orc.reduce(op, G, dims))
        same_on_all_ranks(f"peer {op} {shape} {dims}", got)
# zero-sign rule across GPUs: the extreme value is a zero and the set holds +0 and -0 -> the FIRST zero in logical order
# (fmax / fmin keep their left operand on a tie); columns differ in which rank holds the first zero and with which sign
Z = np.where(rng.random((L, 3, 96)) < 0.5, np.float32(-0.0), np.float32(0.0)).astype(np.float32)
Z[rng.random(Z.shape) < 0.3] = np.float32(-3.0)
Z[:, :, :8] = np.float32(-3.0)       # columns whose maximum is not a zero at all
for op, data in (("max", Z), ("min", -Z)):
    Zs = dist.scatter_from_numpy(data)
    for dims in ((0,), (0, 1)):
        got = al.to_numpy(dist.reduce_sharded(Zs, op, dims))
        check(f"peer {op} zero sign {dims}", got, orc.reduce(op, data, dims))
        same_on_all_ranks(f"peer {op} zero sign {dims}", got)
# in-place allreduce of an already materialised array (same windows: push, combine, collect)
for n_el in (1, 5, 4096, 100003):
    parts = rng.integers(-8, 9, size=(world, n_el)).astype(np.float32)
    for op, fold in (("sum", np.sum), ("max", np.max), ("min", np.min)):
        mine = al.from_numpy(parts[rank].copy())
        dist.allreduce(mine, op)
        check(f"allreduce {op} n={n_el}", al.to_numpy(mine), fold(parts, axis=0).astype(np.float32))
# transposed shard: kept axis not unit-stride -> local reduce + push
G = rng.integers(-8, 9, size=(L, 48, 20)).astype(np.float32)
Gt = dist.scatter_from_numpy(G).transpose((0, 2, 1))
check("peer sum transposed shard", al.to_numpy(dist.reduce_sharded(Gt, "sum", (0,))), orc.reduce("sum", np.transpose(G, (0, 2, 1)), (0,)))
# random normals: rank-ordered combine is deterministic (same bits on every rank and on every run) and within the
# stated reduce_sum tolerance of the oracle
G = rng.standard_normal((L, 64, 1024)).astype(np.float32)
Gs = dist.scatter_from_numpy(G)
first = al.to_numpy(dist.reduce_sharded(Gs, "sum", (0,)))
same_on_all_ranks("peer normal sum", first)
for rep in range(3):
    check(f"peer normal sum rerun #{rep}", al.to_numpy(dist.reduce_sharded(Gs, "sum", (0,))), first)
want = orc.reduce("sum", G, (0,))
bound = 2 * L * 2.0**-24 * np.abs(G).sum(axis=0, keepdims=True)
ok = bool(np.all(np.abs(first.astype(np.float64) - want) <= bound))
if not ok:
    fails.append("peer normal sum tolerance")
if rank == 0:
    print(f"{'peer normal sum tolerance':28s} {'ok' if ok else 'MISMATCH'}", flush=True)
# more outputs than the first window holds (2 x 64 MiB): the windows are remapped collectively
big = rng.integers(-8, 9, size=(2 * world, (1 << 24) + 8)).astype(np.float32)
bs = dist.scatter_from_numpy(big)
check("peer sum after window remap", al.to_numpy(dist.reduce_sharded(bs, "sum", (0,))), big.sum(axis=0, keepdims=True, dtype=np.float32))
del bs, big
if rank == 0:
    print("transport:", dist.transport(), flush=True)

# timing at the C4 shape of the bench (one [64,512,512,64]-proportioned slab per rank is 4 GiB; use a quarter)
T = al.from_numpy(rng.standard_normal((16, 512, 512, 64)).astype(np.float32))
ms = C.c_float()
for mode in (1, 0, 1):
    lib.al_set_tuning(b"peer_collective", mode)
    for _ in range(3):
        dist.reduce_sharded(T, "sum", (0,))
    al.synchronize() if hasattr(al, "synchronize") else lib.al_sync()
    dist.barrier()
    lib.al_timer_start()
    for _ in range(10):
        dist.reduce_sharded(T, "sum", (0,))
    lib.al_timer_stop(C.byref(ms))
    t = dist.allreduce_scalar(ms.value / 10, "max")
    if rank == 0:
        print(f"reduce dims=(0,) of a (16,512,512,64) slab per rank, 2^24 outputs: {'peer-memory' if mode else 'nccl':12s} {t:.4f} ms", flush=True)
lib.al_set_tuning(b"peer_collective", 1)
a1 = al.to_numpy(dist.reduce_sharded(T, "sum", (0,)))
lib.al_set_tuning(b"peer_collective", 0)
a0 = al.to_numpy(dist.reduce_sharded(T, "sum", (0,)))
lib.al_set_tuning(b"peer_collective", 1)
ok = bool(np.allclose(a1, a0, rtol=1e-5, atol=1e-4))
if not ok:
    fails.append("peer vs nccl")
if rank == 0:
    print(f"{'peer vs nccl (allclose)':28s} {'ok' if ok else 'MISMATCH'}", flush=True)

total_fails = int(dist.allreduce_scalar(float(len(fails)), "sum"))
dist.destroy()
if rank == 0:
    print("DIST_CHECK", "PASS" if total_fails == 0 else f"FAIL {fails}")
sys.exit(0 if total_fails == 0 else 1)
