"""Multi-GPU support: one process per GPU, arrays sharded along their LEADING axis.

Nothing in the reference corresponds to this module (it is single-process).
Design:
  * every rank owns rows [rank*L/W, (rank+1)*L/W) of a global array whose leading extent is L;
  * elementwise / broadcast ops are purely local; operands that are broadcast along the leading
    axis (extent 1 or missing) are replicated on every rank;
  * reduce over axes that exclude the leading axis is purely local (output stays sharded);
  * reduce over a set that INCLUDES the leading axis reduces the local slab to a full-shape
    partial and then runs ONE collective: ncclAllReduce over the partial (sum / max / min / prod);
  * overlapping as_strided windows need a (window-1)-element halo, replicated when sharding.
The planning functions at the top are pure host logic (tested on CPU with gloo in
tests/test_dist_cpu.py); the collectives call NCCL through the C ABI (al_comm_*).
"""
from __future__ import annotations

import ctypes as C

# ---- pure host-side planning (no GPU, no library calls) --------------------------------------------


def shard_bounds(extent: int, rank: int, world: int) -> tuple:
    """[lo, hi) of the leading axis owned by `rank`: contiguous blocks, remainder to the low ranks."""
    assert 0 <= rank < world and extent >= 0
    base, rem = divmod(extent, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def operand_placement(global_shape, out_shape) -> str:
    """'sharded' when the operand spans the output's leading axis, else 'replicated'
    (its leading extent is 1 or the axis is missing: stride 0 under broadcasting)."""
    if len(global_shape) < len(out_shape) or global_shape[0] == 1 and out_shape[0] != 1:
        return "replicated"
    return "sharded"


def reduce_plan(ndim: int, dims) -> dict:
    """What a reduce over `dims` of a leading-axis-sharded array needs."""
    dims = tuple(range(ndim)) if dims is None else tuple(dims)
    crosses = 0 in dims
    return {"local_dims": dims, "collective": "allreduce" if crosses else None,
            "output": "replicated" if crosses else "sharded"}


def window_halo_bounds(n_out: int, window: int, rank: int, world: int) -> tuple:
    """Base-buffer element range [lo, hi) a rank needs for its share of n_out sliding windows of
    length `window` and stride 1: its own outputs plus a (window-1)-element halo."""
    lo, hi = shard_bounds(n_out, rank, world)
    return lo, hi + window - 1


# ---- NCCL through the C ABI ---------------------------------------------------------------------------

_state = {"rank": 0, "world": 1}


def init(rank: int, world: int, unique_id: bytes | None = None, *, broadcast=None) -> None:
    """Create the NCCL communicator. Either pass the 128-byte id obtained on rank 0 with
    `unique_id()`, or a `broadcast(obj, src)` callable (e.g. built on torch.distributed) that
    ships it from rank 0."""
    from ._lib import check_rc, lib

    if world == 1:
        _state.update(rank=0, world=1)
        return
    if unique_id is None:
        assert broadcast is not None
        unique_id = broadcast(globals()["unique_id"]() if rank == 0 else None, 0)
    assert len(unique_id) == 128
    check_rc(lib.al_comm_init(unique_id, rank, world))
    _state.update(rank=rank, world=world)


def unique_id() -> bytes:
    from ._lib import check_rc, lib

    buf = C.create_string_buffer(128)
    check_rc(lib.al_comm_unique_id(buf))
    return buf.raw


def rank() -> int:
    return _state["rank"]


def world() -> int:
    return _state["world"]


def allreduce(array, op: str = "sum"):
    """In-place NCCL allreduce of a contiguous array (no-op for world == 1)."""
    from ._lib import REDOPS, check_rc, lib

    if _state["world"] > 1:
        check_rc(lib.al_comm_allreduce(array.buffer, REDOPS.index(op)))
    return array


def reduce_sharded(array, op: str = "sum", dims=None):
    """reduce_<op> of a leading-axis shard. Local unless axis 0 is reduced; then ONE collective call
    (al_comm_reduce): the local kernel stores its partial straight into the other GPUs' peer windows and
    the partials are combined in rank order (ncclAllReduce when the windows cannot be mapped)."""
    from . import core
    from ._lib import REDOPS, lib

    plan = reduce_plan(array.ndim, dims)
    if not plan["collective"] or _state["world"] == 1:
        return getattr(core, f"reduce_{op}")(array, dims=plan["local_dims"])
    dims_c, n = core._dims_arg(array, plan["local_dims"])
    return core._new(lib.al_comm_reduce(REDOPS.index(op), array.buffer, dims_c, n))


def transport() -> str:
    """How cross-GPU reductions travel: 'peer-memory', 'nccl (<why>)' or 'single process'."""
    from ._lib import lib

    return lib.al_comm_transport().decode()


def scatter_from_numpy(global_array):
    """Upload this rank's leading-axis block of a host array that every rank holds in full."""
    from . import core

    lo, hi = shard_bounds(global_array.shape[0], _state["rank"], _state["world"])
    return core.from_numpy(global_array[lo:hi])


def window_shard_from_numpy(base, window: int):
    """Upload the slice of a 1-D base buffer this rank needs for its share of the stride-1 sliding
    windows of length `window` (own outputs + halo), and return (array, n_local_outputs)."""
    from . import core

    n_out = base.shape[0] - window + 1
    lo, hi = window_halo_bounds(n_out, window, _state["rank"], _state["world"])
    return core.from_numpy(base[lo:hi]), hi - lo - window + 1


def allgather(shard):
    """Concatenate equal-sized leading-axis shards on every rank (NCCL allgather)."""
    from . import core
    from ._lib import check_rc, lib

    if _state["world"] == 1:
        return shard
    src = shard
    if src.stride != tuple(_row_major(src.shape)):
        src = core.copy(src, deep=True)
    full = core.zeros((src.shape[0] * _state["world"],) + tuple(src.shape[1:]))
    check_rc(lib.al_comm_allgather(src.buffer, full.buffer))
    return full


def allgather_ragged(shard, total_rows: int):
    """allgather for leading-axis blocks of unequal height (shard_bounds): pad to the tallest block, gather,
    drop the padding."""
    from . import core

    world = _state["world"]
    if world == 1:
        return shard
    if total_rows % world == 0:
        return allgather(shard)
    rest = tuple(shard.shape[1:])
    tallest = -(-total_rows // world)
    padded = core.zeros((tallest,) + rest)
    padded[tuple([slice(0, shard.shape[0])] + [slice(0, e) for e in rest])] = shard
    stacked = allgather(padded)
    blocks = []
    for r in range(world):
        lo, hi = shard_bounds(total_rows, r, world)
        if hi > lo:
            blocks.append(stacked[tuple([slice(r * tallest, r * tallest + hi - lo)] + [slice(0, e) for e in rest])])
    return blocks[0] if len(blocks) == 1 else core.cat(blocks, [0])


def alltoall(send):
    """Block r (of `world` equal blocks of the flattened array) goes to rank r; returns the received blocks."""
    from . import core
    from ._lib import check_rc, lib

    src = send if send.stride == tuple(_row_major(send.shape)) else core.copy(send, deep=True)
    recv = core.zeros(src.shape)
    check_rc(lib.al_comm_alltoall(src.buffer, recv.buffer))
    return recv


def reshard_plan(global_shape, perm, world: int) -> dict:
    """How transpose(perm) of a leading-axis-sharded array is carried out (pure planning, CPU-testable).
    `local`: axis 0 stays in front, every rank permutes its own block. `alltoall`: the new leading axis is cut
    into `world` blocks, block j of every rank's transposed piece travels to rank j, and the pieces are
    interleaved along `q`, the new position of the old leading axis. `gather`: extents do not divide evenly;
    replicate, transpose, keep the own block."""
    perm = tuple(perm)
    assert sorted(perm) == list(range(len(global_shape))), "perm must be a permutation of the axes"
    new_shape = tuple(global_shape[p] for p in perm)
    if perm[0] == 0 or world == 1:
        return {"how": "local", "shape": new_shape}
    if global_shape[0] % world or new_shape[0] % world:
        return {"how": "gather", "shape": new_shape}
    q = perm.index(0)
    rows, lead = global_shape[0] // world, new_shape[0] // world
    a = 1
    for e in new_shape[1:q]:
        a *= e
    b = 1
    for e in new_shape[q + 1:]:
        b *= e
    return {"how": "alltoall", "shape": new_shape, "q": q, "rows": rows, "lead": lead,
            "recv_view": (world, lead, a, rows, b), "interleave": (1, 2, 0, 3, 4),
            "local_shape": (lead,) + new_shape[1:q] + (rows * world,) + new_shape[q + 1:]}


def _row_major(shape):
    out, run = [], 1
    for e in reversed(shape):
        out.append(run)
        run *= e
    return tuple(reversed(out))


class ShardedArray:
    """A global array whose leading axis is split into contiguous blocks, one per rank ("sharded"), or
    that every rank holds in full ("replicated"). Arithmetic follows the plan in the module docstring:
    everything is local except a reduction over axis 0, which ends in one NCCL allreduce."""

    def __init__(self, local, global_shape, placement: str):
        assert placement in ("sharded", "replicated")
        self.local, self.shape, self.placement = local, tuple(global_shape), placement

    # -- construction ------------------------------------------------------------------------------
    @classmethod
    def scatter(cls, global_array):
        """Every rank passes the same host array; each uploads only its block of axis 0."""
        return cls(scatter_from_numpy(global_array), global_array.shape, "sharded")

    @classmethod
    def replicate(cls, global_array):
        from . import core
        return cls(core.from_numpy(global_array), global_array.shape, "replicated")

    # -- helpers -----------------------------------------------------------------------------------
    def _local_block(self, out_ndim: int):
        """The local operand to use inside an op whose output is sharded."""
        if self.placement == "sharded":
            return self.local
        if len(self.shape) < out_ndim or self.shape[0] == 1:
            return self.local  # broadcast along axis 0: the replica is used as it is
        lo, hi = shard_bounds(self.shape[0], _state["rank"], _state["world"])
        sl = tuple([slice(lo, hi)] + [slice(0, e) for e in self.shape[1:]])
        return self.local[sl]  # a view of this rank's rows of the replica

    def _binary(self, name: str, other, reverse: bool = False):
        from . import core
        fn = getattr(core, name)
        if isinstance(other, core.Number):
            res = fn(other, self.local) if reverse else fn(self.local, other)
            return ShardedArray(res, self.shape, self.placement)
        if isinstance(other, core.NDArray):
            other = ShardedArray(other, other.shape, "replicated")
        assert isinstance(other, ShardedArray)
        import numpy as np
        out_shape = tuple(np.broadcast_shapes(self.shape, other.shape))
        sharded = "sharded" in (self.placement, other.placement)
        if sharded:
            for op in (self, other):
                if op.placement == "sharded":
                    assert len(op.shape) == len(out_shape) and op.shape[0] == out_shape[0], \
                        "a sharded operand must span the output's leading axis"
            a, b = self._local_block(len(out_shape)), other._local_block(len(out_shape))
        else:
            a, b = self.local, other.local
        res = fn(b, a) if reverse else fn(a, b)
        return ShardedArray(res, out_shape, "sharded" if sharded else "replicated")

    def __add__(self, o): return self._binary("add", o)
    def __radd__(self, o): return self._binary("add", o, True)
    def __sub__(self, o): return self._binary("sub", o)
    def __rsub__(self, o): return self._binary("sub", o, True)
    def __mul__(self, o): return self._binary("mul", o)
    def __rmul__(self, o): return self._binary("mul", o, True)
    def __truediv__(self, o): return self._binary("div", o)
    def __rtruediv__(self, o): return self._binary("div", o, True)

    def _reduce(self, op: str, dims):
        from . import core
        dims = tuple(range(len(self.shape))) if dims is None else tuple(dims)
        out_shape = tuple(1 if d in dims else e for d, e in enumerate(self.shape))
        if self.placement == "replicated":
            return ShardedArray(getattr(core, f"reduce_{op}")(self.local, dims=dims), out_shape, "replicated")
        plan = reduce_plan(len(self.shape), dims)
        return ShardedArray(reduce_sharded(self.local, op, dims), out_shape, plan["output"])

    def reduce_sum(self, *, dims=None): return self._reduce("sum", dims)
    def reduce_max(self, *, dims=None): return self._reduce("max", dims)
    def reduce_min(self, *, dims=None): return self._reduce("min", dims)

    # -- re-sharding ---------------------------------------------------------------------------------
    def transpose(self, perm):
        """transpose(perm) of the GLOBAL array; the result is sharded along its new leading axis. Local when
        axis 0 stays in front, otherwise one all-to-all (see reshard_plan)."""
        from . import core
        world, me = _state["world"], _state["rank"]
        if self.placement == "replicated":
            return ShardedArray(core.transpose(self.local, perm), tuple(self.shape[p] for p in perm), "replicated")
        plan = reshard_plan(self.shape, perm, world)
        if plan["how"] == "local":
            return ShardedArray(core.transpose(self.local, perm), plan["shape"], "sharded")
        if plan["how"] == "gather":
            whole = core.transpose(allgather_ragged(self.local, self.shape[0]), perm)
            lo, hi = shard_bounds(plan["shape"][0], me, world)
            sl = tuple([slice(lo, hi)] + [slice(0, e) for e in plan["shape"][1:]])
            return ShardedArray(core.copy(whole[sl], deep=True), plan["shape"], "sharded")
        piece = core.copy(core.transpose(self.local, perm), deep=True)   # (new leading axis in full, ..., own rows at q, ...)
        got = alltoall(piece)                                              # block i = rank i's rows of MY leading block
        mixed = core.transpose(core.reshape(got, plan["recv_view"]), plan["interleave"])
        return ShardedArray(core.reshape(mixed, plan["local_shape"]), plan["shape"], "sharded")

    # -- materialisation ----------------------------------------------------------------------------
    def gather(self):
        """The full array on every rank (allgather; ragged blocks are padded for the exchange)."""
        if self.placement == "replicated" or _state["world"] == 1:
            return self.local
        return allgather_ragged(self.local, self.shape[0])

    def to_numpy(self):
        from . import core
        return core.to_numpy(self.gather())


def destroy() -> None:
    from ._lib import lib

    lib.al_comm_destroy()
    _state.update(rank=0, world=1)
