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


# ---- SPMD front end (arraylib_b200.spmd): pure planning ----------------------------------------------------


def test_binary_plan_covers_block_whole_rows_and_the_replicated_fallback():
    b8 = dist.all_bounds(1024, 8)
    A = ((1024, 1, 4096), "sharded", 0, b8)
    B = ((1, 1024, 4096), "replicated", 1, dist.all_bounds(1024, 8))
    c = ((4096,), "replicated", 0, dist.all_bounds(4096, 8))
    plan = dist.binary_plan([A, B], (1024, 1024, 4096))  # config C3: A's blocks, B whole
    assert plan == {"output": "sharded", "axis": 0, "bounds": b8, "operands": ["block", "whole"]}
    t = ((1024, 1024, 4096), "sharded", 0, b8)
    assert dist.binary_plan([t, c], (1024, 1024, 4096))["operands"] == ["block", "whole"]
    # a replica that spans the shard axis contributes this rank's rows
    full = ((1024, 1024, 4096), "replicated", 0, b8)
    assert dist.binary_plan([t, full], (1024, 1024, 4096))["operands"] == ["block", ("rows", 0)]
    # B cut along ITS leading non-unit axis (axis 1) while the output is cut along axis 0: gathered, used whole
    Bs = ((1, 1024, 4096), "sharded", 1, b8)
    assert dist.binary_plan([A, Bs], (1024, 1024, 4096))["operands"] == ["block", "whole"]
    # same axis, other rows (uneven cut): the second operand goes through a replica
    odd = tuple((lo + (1 if lo else 0), hi + (1 if hi < 1024 else 0)) for lo, hi in b8)
    t2 = ((1024, 1024, 4096), "sharded", 0, odd)
    assert dist.binary_plan([t, t2], (1024, 1024, 4096))["operands"] == ["block", ("rows", 0)]
    # [N,1] + [N] -> [N,N]: the output's shard axis is axis 0, the 1-D operand spans axis 1 only
    col, row = ((512, 1), "sharded", 0, dist.all_bounds(512, 2)), ((512,), "sharded", 0, dist.all_bounds(512, 2))
    assert dist.binary_plan([col, row], (512, 512))["operands"] == ["block", ("rows", 1)] or \
        dist.binary_plan([col, row], (512, 512))["operands"] == ["block", "whole"]
    # nothing is cut along the output's shard axis: replicated
    assert dist.binary_plan([B, c], (1, 1024, 4096))["output"] == "replicated"


def test_broadcast_shape_matches_numpy_and_rejects_like_the_reference():
    for lhs, rhs in [((3, 1, 5), (4, 5)), ((1,), (7, 2)), ((6, 1), (1, 6)), ((2, 3, 4), (2, 3, 4))]:
        assert dist.broadcast_shape(lhs, rhs) == tuple(np.broadcast_shapes(lhs, rhs))
    with pytest.raises(AssertionError):
        dist.broadcast_shape((3, 4), (5,))


def test_spmd_placement_policy_and_strided_rows():
    from arraylib_b200 import spmd
    assert spmd.placement_for((1024, 1, 4096), 8, 1 << 20) == "sharded"
    assert spmd.placement_for((3, 4), 8, 1 << 20) == "replicated"          # small
    assert spmd.placement_for((4, 1 << 20), 8, 1 << 20) == "replicated"    # fewer rows than ranks
    assert spmd.placement_for((1, 1, 64, 1 << 16), 8, 1 << 20) == "sharded"  # leading unit axes are skipped
    assert spmd.placement_for((1 << 24,), 1, 0) == "replicated"            # one rank
    # rows 1, 4, 7, ... < 20 split over blocks [0,5) [5,12) [12,20)
    seq = list(range(1, 20, 3))
    seen = []
    for lo, hi in ((0, 5), (5, 12), (12, 20)):
        got = spmd.strided_rows_in_block(1, 20, 3, lo, hi)
        k0, k1, l0, l1 = got
        rows = list(range(lo + l0, lo + l1, 3))
        assert rows == seq[k0:k1] and all(lo <= r < hi for r in rows)
        seen += rows
    assert seen == seq
    assert spmd.strided_rows_in_block(0, 4, 1, 8, 16) is None
    assert spmd.strided_rows_in_block(10, 40, 16, 0, 10) is None
    assert spmd.owner_of(5, ((0, 5), (5, 12))) == 1
    # exhaustive against numpy index arithmetic
    for start, end, step in [(0, 33, 1), (2, 31, 5), (7, 8, 3), (0, 64, 64), (5, 60, 7)]:
        want = list(range(start, end, step))
        got = []
        for lo, hi in dist.all_bounds(64, 5):
            r = spmd.strided_rows_in_block(start, end, step, lo, hi)
            if r:
                assert want[r[0]:r[1]] == list(range(lo + r[2], lo + r[3], step))
                got += want[r[0]:r[1]]
        assert got == want


def test_shim_selects_the_spmd_front_end_only_on_request(monkeypatch):
    import importlib
    import arraylib
    assert arraylib.NDArray.__module__ == "arraylib_b200.core"   # default: single device, even under torchrun
    monkeypatch.setenv("WORLD_SIZE", "2")
    monkeypatch.delenv("ARRAYLIB_SPMD", raising=False)
    assert importlib.reload(arraylib).NDArray.__module__ == "arraylib_b200.core"
    monkeypatch.delenv("WORLD_SIZE")
    importlib.reload(arraylib)


def test_scatter_refuses_fewer_rows_than_ranks_on_every_rank(monkeypatch):
    """with fewer leading rows than ranks some ranks would own an empty block and fail alone while the
    others walked into a collective. The check runs before anything is uploaded, with the same outcome on every rank."""
    monkeypatch.setitem(dist._state, "world", 4)
    for rank in range(4):
        monkeypatch.setitem(dist._state, "rank", rank)
        with pytest.raises(ValueError, match="smaller than the world size"):
            dist.scatter_from_numpy(np.zeros((3, 8), np.float32))
        with pytest.raises(ValueError, match="smaller than the world size"):
            dist.ShardedArray.scatter(np.zeros((1, 2, 8), np.float32))
        with pytest.raises(ValueError, match="smaller than the world size"):
            dist.window_shard_from_numpy(np.zeros((66,), np.float32), 64)
