"""Multi-GPU parity check (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/dist_check.py
Every rank builds the same global host arrays, uploads only its leading-axis shard (plus replicated
small operands / window halo), runs the sharded hot path on its GPU and the results are gathered
and compared with the CPU oracle on the whole arrays. Test infrastructure, not product code."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as td  # noqa: E402

import arraylib as al  # noqa: E402
from arraylib_b200 import dist  # noqa: E402
from arraylib_b200._lib import lib  # noqa: E402
from oracle import oracle as orc  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
assert lib.al_init(local) == 0
torch.cuda.set_device(local)
td.init_process_group("nccl", device_id=torch.device("cuda", local))


def bcast(obj, src):
    box = [obj]
    td.broadcast_object_list(box, src=src)
    return box[0]


dist.init(rank, world, broadcast=bcast)
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
r = SX.reduce_sum(dims=(0, 2))
assert r.placement == "replicated" and r.shape == (1, 20, 1, 64)
check("ShardedArray reduce (0,2)", r.to_numpy(), orc.reduce("sum", X, (0, 2)))
r = (SX - 1.0).reduce_max(dims=(1, 3))
assert r.placement == "sharded"
check("ShardedArray reduce_max (1,3)", r.to_numpy(), orc.reduce("max", orc.binary("sub", X, 1.0), (1, 3)))
flag = torch.tensor([len(fails)], device="cuda")
td.all_reduce(flag)
dist.destroy()
td.destroy_process_group()
if rank == 0:
    print("DIST_CHECK", "PASS" if int(flag.item()) == 0 else f"FAIL {fails}")
sys.exit(0 if int(flag.item()) == 0 else 1)
