"""CPU tests of the multi-GPU host logic (arraylib_b200.dist planning functions), run with
world_size 2 and 3 over the gloo backend. Each rank plays one GPU: it computes its shard with the
CPU oracle, the only exchange is the allreduce the plan asks for, and the stitched result must
equal the oracle's answer on the whole array."""
import os
import socket

import numpy as np
import pytest

from arraylib_b200 import dist


def test_shard_bounds_partition():
    for extent in (1, 7, 8, 64, 1000, 1024):
        for world in (1, 2, 3, 4, 8):
            blocks = [dist.shard_bounds(extent, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == extent
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_operand_placement_and_reduce_plan():
    assert dist.operand_placement((1024, 1, 4096), (1024, 1024, 4096)) == "sharded"
    assert dist.operand_placement((1, 1024, 4096), (1024, 1024, 4096)) == "replicated"
    assert dist.operand_placement((4096,), (1024, 1024, 4096)) == "replicated"
    assert dist.reduce_plan(4, (1, 2)) == {"local_dims": (1, 2), "collective": None, "output": "sharded"}
    assert dist.reduce_plan(4, (0,))["collective"] == "allreduce"
    assert dist.reduce_plan(4, None) == {"local_dims": (0, 1, 2, 3), "collective": "allreduce", "output": "replicated"}


def test_window_halo():
    n, w = 1000, 64
    n_out = n - w + 1
    spans = [dist.window_halo_bounds(n_out, w, r, 4) for r in range(4)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for r in range(3):  # consecutive ranks overlap by exactly the halo
        assert spans[r][1] - spans[r + 1][0] == w - 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as td
    from oracle import oracle as orc

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    td.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)  # every rank builds the same global arrays
    A = rng.integers(-8, 9, size=(12, 1, 16)).astype(np.float32)
    B = rng.integers(-8, 9, size=(1, 5, 16)).astype(np.float32)
    c = rng.integers(-8, 9, size=(16,)).astype(np.float32)
    X = rng.integers(-8, 9, size=(12, 6, 5, 8)).astype(np.float32)
    base = rng.integers(-8, 9, size=(500,)).astype(np.float32)
    lo, hi = dist.shard_bounds(12, rank, world)
    res = {}
    # broadcast chain: A is sharded, B and c are replicated, no communication
    assert dist.operand_placement(A.shape, (12, 5, 16)) == "sharded"
    assert dist.operand_placement(B.shape, (12, 5, 16)) == "replicated"
    res["bcast"] = orc.binary("sum", orc.binary("sum", A[lo:hi], B), c)
    # reductions
    for name, dims in (("r12", (1, 2)), ("r3", (3,)), ("r0", (0,)), ("rall", None)):
        plan = dist.reduce_plan(4, dims)
        part = orc.reduce("sum", X[lo:hi], plan["local_dims"])
        if plan["collective"] == "allreduce":
            t = torch.from_numpy(part)
            td.all_reduce(t, op=td.ReduceOp.SUM)
            part = t.numpy()
        res[name] = part
    # sliding windows with halo
    w = 64
    n_out = base.size - w + 1
    blo, bhi = dist.window_halo_bounds(n_out, w, rank, world)
    local = base[blo:bhi]
    win = np.lib.stride_tricks.as_strided(local, shape=(local.size - w + 1, w), strides=(4, 4))
    res["win"] = orc.reduce("sum", win, (1,))
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **res)
    td.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_plans_match_the_oracle_over_gloo(world, tmp_path):
    import torch.multiprocessing as mp
    from oracle import oracle as orc

    orc.build()
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    rng = np.random.default_rng(7)
    A = rng.integers(-8, 9, size=(12, 1, 16)).astype(np.float32)
    B = rng.integers(-8, 9, size=(1, 5, 16)).astype(np.float32)
    c = rng.integers(-8, 9, size=(16,)).astype(np.float32)
    X = rng.integers(-8, 9, size=(12, 6, 5, 8)).astype(np.float32)
    base = rng.integers(-8, 9, size=(500,)).astype(np.float32)
    cat = lambda k: np.concatenate([p[k] for p in parts], axis=0)  # noqa: E731
    assert np.array_equal(cat("bcast"), orc.binary("sum", orc.binary("sum", A, B), c))
    assert np.array_equal(cat("r12"), orc.reduce("sum", X, (1, 2)))
    assert np.array_equal(cat("r3"), orc.reduce("sum", X, (3,)))
    for p in parts:  # reductions across the sharded axis are replicated after the allreduce
        assert np.array_equal(p["r0"], orc.reduce("sum", X, (0,)))
        assert np.array_equal(p["rall"], orc.reduce("sum", X, None))
    win = np.lib.stride_tricks.as_strided(base, shape=(base.size - 63, 64), strides=(4, 4))
    assert np.array_equal(cat("win"), orc.reduce("sum", win, (1,)))


@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("shape,perm", [((24, 48), (1, 0)), ((24, 6, 48), (2, 0, 1)), ((24, 6, 48), (2, 1, 0)), ((24, 48, 5), (1, 2, 0)),
                                        ((24, 5, 48, 3), (2, 3, 0, 1)), ((24, 48), (0, 1)), ((24, 6, 48), (0, 2, 1)), ((10, 48), (1, 0)),
                                        ((24, 7), (1, 0))])
def test_reshard_plan_reproduces_the_global_transpose(world, shape, perm):
    """ShardedArray.transpose's plan, with the all-to-all played out in NumPy: every rank must end up with its
    leading-axis block of the globally transposed array."""
    g = np.arange(int(np.prod(shape)), dtype=np.float32).reshape(shape)
    want = np.transpose(g, perm)
    plan = dist.reshard_plan(shape, perm, world)
    assert plan["shape"] == want.shape
    locals_ = [g[slice(*dist.shard_bounds(shape[0], r, world))] for r in range(world)]
    if plan["how"] == "local":
        assert perm[0] == 0
        got = [np.transpose(x, perm) for x in locals_]
    elif plan["how"] == "gather":
        assert shape[0] % world or want.shape[0] % world
        return
    else:
        pieces = [np.ascontiguousarray(np.transpose(x, perm)).reshape(world, -1) for x in locals_]  # block j -> rank j
        got = []
        for j in range(world):
            recv = np.stack([pieces[i][j] for i in range(world)])  # block i <- rank i
            got.append(np.transpose(recv.reshape(plan["recv_view"]), plan["interleave"]).reshape(plan["local_shape"]))
    for r in range(world):
        lo, hi = dist.shard_bounds(want.shape[0], r, world)
        np.testing.assert_array_equal(got[r], want[lo:hi])
