"""Multi-GPU support: one process per GPU, arrays sharded along their LEADING NON-UNIT axis.

Nothing in the reference corresponds to this module (it is single-process).
Design:
  * every rank owns a contiguous block of rows of a global array ("sharded"), or the whole array
    ("replicated"); the shard axis is the first axis whose global extent is > 1, so a block is one
    contiguous run of the flattened array;
  * elementwise / broadcast ops are purely local; operands that are broadcast along the shard axis
    (extent 1 or missing) are replicated on every rank;
  * reduce over axes that exclude the shard axis is purely local (output stays sharded);
  * reduce over a set that INCLUDES the shard axis reduces the local slab to a full-shape partial and
    then runs ONE exchange over NVLink: the result stays sharded along its leading non-reduced axis
    (reduce-scatter: one hand-over, no broadcast) or, on request / when it cannot be cut evenly, is
    replicated (allreduce);
  * overlapping as_strided windows need a (window-1)-element halo from the next rank.
The process group needs NO PyTorch: ranks find each other through the torchrun-style environment
(RANK / WORLD_SIZE / LOCAL_RANK / MASTER_PORT) and a file rendezvous for the 128-byte NCCL id;
barriers and scalar reductions run through the library's own collectives.
The planning functions at the top are pure host logic (tested on CPU with gloo in
tests/test_dist_cpu.py); the collectives go through the C ABI (al_comm_*).
"""
from __future__ import annotations

import ctypes as C
import os
import time

# ---- pure host-side planning (no GPU, no library calls) --------------------------------------------


def shard_bounds(extent: int, rank: int, world: int) -> tuple:
    """[lo, hi) of the shard axis owned by `rank`: contiguous blocks, remainder to the low ranks."""
    assert 0 <= rank < world and extent >= 0
    base, rem = divmod(extent, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_bounds(extent: int, world: int) -> tuple:
    return tuple(shard_bounds(extent, r, world) for r in range(world))


def shard_axis(shape) -> int:
    """The axis a global array is cut along: its first axis with extent > 1 (0 for an all-ones shape)."""
    for i, e in enumerate(shape):
        if e > 1:
            return i
    return 0


def operand_placement(global_shape, out_shape) -> str:
    """'sharded' when the operand spans the output's leading axis, else 'replicated'
    (its leading extent is 1 or the axis is missing: stride 0 under broadcasting)."""
    if len(global_shape) < len(out_shape) or global_shape[0] == 1 and out_shape[0] != 1:
        return "replicated"
    return "sharded"


def reduce_plan(ndim: int, dims, axis: int = 0, shape=None, world: int = 1) -> dict:
    """What a reduce over `dims` of an array sharded along `axis` needs. With `shape` and `world` the plan also says
    whether the result can stay sharded (reduce-scatter) along its leading non-reduced axis."""
    dims = tuple(range(ndim)) if dims is None else tuple(dims)
    crosses = axis in dims
    plan = {"local_dims": dims, "collective": "allreduce" if crosses else None,
            "output": "replicated" if crosses else "sharded"}
    if crosses and shape is not None and world > 1:
        kept = tuple(1 if d in dims else e for d, e in enumerate(shape))
        ax = shard_axis(kept)
        nout = 1
        for e in kept:
            nout *= e
        if kept[ax] > 1 and kept[ax] % world == 0 and (nout // world) % 4 == 0:
            plan.update(collective="reduce_scatter", output="sharded", out_axis=ax)
    return plan


def window_halo_bounds(n_out: int, window: int, rank: int, world: int) -> tuple:
    """Base-buffer element range [lo, hi) a rank needs for its share of n_out sliding windows of
    length `window` and stride 1: its own outputs plus a (window-1)-element halo."""
    lo, hi = shard_bounds(n_out, rank, world)
    return lo, hi + window - 1


def window_view_bounds(base_bounds, window: int) -> tuple:
    """Row ranges of the (n - window + 1, window) stride-(1, 1) view of a sharded 1-D base whose blocks are
    `base_bounds`: rank r keeps the windows that START in its block (it borrows window-1 elements from the ranks
    after it); the windows that would start in the last window-1 elements do not exist."""
    n = base_bounds[-1][1]
    n_out = n - window + 1
    return tuple((min(lo, n_out), min(hi, n_out)) for lo, hi in base_bounds)


def reshape_plan(global_shape, bounds, new_shape) -> dict:
    """How reshape(new_shape) of an array sharded along its leading non-unit axis is carried out. `local`: every
    block boundary falls on a row boundary of the new shape's leading non-unit axis, so each rank relabels its own
    block; `gather`: replicate first."""
    total = 1
    for e in global_shape:
        total *= e
    new_total = 1
    for e in new_shape:
        new_total *= e
    assert total == new_total, f"incorrect reshape {tuple(new_shape)} != {tuple(global_shape)}"
    ax = shard_axis(global_shape)
    inner = 1
    for e in global_shape[ax + 1:]:
        inner *= e
    nax = shard_axis(new_shape)
    ninner = 1
    for e in new_shape[nax + 1:]:
        ninner *= e
    cuts = [(lo * inner, hi * inner) for lo, hi in bounds]
    if all(lo % ninner == 0 and hi % ninner == 0 and hi > lo for lo, hi in cuts):
        return {"how": "local", "axis": nax, "bounds": tuple((lo // ninner, hi // ninner) for lo, hi in cuts)}
    return {"how": "gather"}


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


def broadcast_shape(lhs, rhs) -> tuple:
    """NumPy-style right-aligned broadcast of two shapes (arraylib.c:171-206); AssertionError like the reference's
    Python layer (core.py:371-379) when they do not match."""
    nd = max(len(lhs), len(rhs))
    l = (1,) * (nd - len(lhs)) + tuple(lhs)
    r = (1,) * (nd - len(rhs)) + tuple(rhs)
    assert all(a == b or a == 1 or b == 1 for a, b in zip(l, r)), f"cannot broadcast shapes {tuple(lhs)} and {tuple(rhs)}"
    return tuple(max(a, b) for a, b in zip(l, r))


def binary_plan(operands, out_shape) -> dict:
    """How `lhs o rhs` runs on sharded operands (pure planning). `operands`: (shape, placement, axis, bounds) each.
    The output is cut along ITS leading non-unit axis with the bounds of the first operand that is already cut
    there; every operand is then used as `block` (its local block as it is), `whole` (it is broadcast along that
    axis) or `("rows", k)` (this rank's rows of its axis k, taken from a replica -- a sharded operand that is cut
    along another axis or at other rows is gathered first). When no operand is cut along the output's shard axis
    everything is computed replicated."""
    nd = len(out_shape)
    out_axis = shard_axis(out_shape)
    bounds = None
    for shape, placement, axis, b in operands:
        k = out_axis - (nd - len(shape))
        if placement == "sharded" and k == axis and shape[k] == out_shape[out_axis]:
            bounds = tuple(b)
            break
    if bounds is None:
        return {"output": "replicated", "operands": ["whole"] * len(operands)}
    how = []
    for shape, placement, axis, b in operands:
        k = out_axis - (nd - len(shape))
        if k < 0 or shape[k] == 1:
            how.append("whole")
        elif placement == "sharded" and k == axis and tuple(b) == bounds:
            how.append("block")
        else:
            how.append(("rows", k))
    return {"output": "sharded", "axis": out_axis, "bounds": bounds, "operands": how}


def _count(shape) -> int:
    n = 1
    for e in shape:
        n *= e
    return n


def _row_major(shape):
    out, run = [], 1
    for e in reversed(shape):
        out.append(run)
        run *= e
    return tuple(reversed(out))


# ---- rendezvous without PyTorch ------------------------------------------------------------------------


def rendezvous_dir() -> str:
    """Directory through which the ranks of ONE launch on this node exchange the NCCL id. ARRAYLIB_RDZV_DIR wins;
    otherwise the name is built from what all workers of a torchrun (or arraylib_b200.launch) launch share and
    other launches do not: the master port and the launcher's pid (the workers' common parent)."""
    explicit = os.environ.get("ARRAYLIB_RDZV_DIR")
    if explicit:
        return explicit
    port = os.environ.get("MASTER_PORT", "0")
    run = os.environ.get("TORCHELASTIC_RUN_ID", "none")
    return os.path.join(os.environ.get("TMPDIR", "/tmp"), f"arraylib_b200_rdzv_{os.getuid()}_{port}_{run}_{os.getppid()}")


def exchange_bytes(key: str, rank: int, payload: bytes | None, *, directory: str | None = None, timeout: float = 300.0) -> bytes:
    """Rank 0 publishes `payload` under `key` (write + atomic rename); every other rank polls for it."""
    directory = directory or rendezvous_dir()
    path = os.path.join(directory, key)
    if rank == 0:
        assert payload is not None
        os.makedirs(directory, exist_ok=True)
        tmp = f"{path}.tmp{os.getpid()}"
        with open(tmp, "wb") as f:
            f.write(payload)
        os.replace(tmp, path)
        return payload
    deadline = time.monotonic() + timeout
    while True:
        try:
            with open(path, "rb") as f:
                data = f.read()
            if data:
                return data
        except FileNotFoundError:
            pass
        if time.monotonic() > deadline:
            raise TimeoutError(f"rank {rank}: nothing published under {path} within {timeout:.0f} s")
        time.sleep(0.005)


# ---- communicator ---------------------------------------------------------------------------------------

_state = {"rank": 0, "world": 1, "generation": 0, "rdzv": None}


def init(rank: int, world: int, unique_id: bytes | None = None, *, broadcast=None) -> None:
    """Create the NCCL communicator. Pass the 128-byte id obtained on rank 0 with `unique_id()`, or a
    `broadcast(obj, src)` callable that ships it from rank 0, or neither: the id then travels through the file
    rendezvous (`rendezvous_dir()`), which is what `init_from_env()` does."""
    from ._lib import check_rc, lib

    if world == 1:
        _state.update(rank=0, world=1)
        return
    if unique_id is None:
        mine = globals()["unique_id"]() if rank == 0 else None
        if broadcast is not None:
            unique_id = broadcast(mine, 0)
        else:
            _state["generation"] += 1
            _state["rdzv"] = rendezvous_dir()
            unique_id = exchange_bytes(f"nccl_id_{_state['generation']}", rank, mine, directory=_state["rdzv"])
    assert len(unique_id) == 128
    check_rc(lib.al_comm_init(unique_id, rank, world))
    _state.update(rank=rank, world=world)


def init_from_env() -> tuple:
    """Bind this process to GPU LOCAL_RANK and join the job described by RANK / WORLD_SIZE (the variables torchrun
    and `python -m arraylib_b200.launch` export). Returns (rank, world). No PyTorch involved."""
    from ._lib import check_rc, lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    check_rc(lib.al_init(local))
    init(rank, world)
    return rank, world


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
    """In-place allreduce of a contiguous array (no-op for world == 1)."""
    from ._lib import REDOPS, check_rc, lib

    if _state["world"] > 1:
        check_rc(lib.al_comm_allreduce(array.buffer, REDOPS.index(op)))
    return array


def allreduce_scalar(value: float, op: str = "sum") -> float:
    """A host scalar reduced over the ranks (float32 on the wire). Synchronises the calling rank."""
    from . import core

    if _state["world"] == 1:
        return float(value)
    box = core.full((4,), float(value))
    allreduce(box, op)
    return float(core.to_numpy(box)[0])


def broadcast_scalar(value: float, src: int) -> float:
    """The float32 `value` of rank `src` on every rank (bit pattern preserved: NaN payloads and -0.0 survive,
    which a sum over zero-padded contributions would not guarantee)."""
    from . import core

    if _state["world"] == 1:
        return float(value)
    got = core.to_numpy(allgather(core.full((4,), float(value))))
    return float(got[4 * src])


def barrier() -> None:
    """Every rank has reached this point and its device is idle."""
    from . import core

    if _state["world"] > 1:
        allreduce_scalar(0.0, "sum")
    core.synchronize()


def reduce_sharded(array, op: str = "sum", dims=None, *, output: str = "replicated", axis: int = 0):
    """reduce_<op> of a shard (cut along `axis`). Local unless the shard axis is reduced; then ONE collective call:
    the local kernel stores its partial straight into the other GPUs' peer windows and the partials are combined in
    rank order. `output="replicated"`: every rank gets the whole result (al_comm_reduce); `output="sharded"`: rank r
    gets block r of the result's leading non-reduced axis (al_comm_reduce_scatter; raises ValueError when the result
    cannot be cut evenly -- ask `reduce_plan` first)."""
    from . import core
    from ._lib import REDOPS, lib

    assert output in ("replicated", "sharded")
    plan = reduce_plan(array.ndim, dims, axis)
    if not plan["collective"] or _state["world"] == 1:
        return getattr(core, f"reduce_{op}")(array, dims=plan["local_dims"])
    dims_c, n = core._dims_arg(array, plan["local_dims"])
    return core._new(lib.al_comm_reduce_ex(REDOPS.index(op), array.buffer, dims_c, n, axis, 1 if output == "sharded" else 0))


def transport() -> str:
    """How cross-GPU reductions travel: 'peer-memory', 'nccl (<why>)' or 'single process'."""
    from ._lib import lib

    return lib.al_comm_transport().decode()


def _check_rows(extent: int, what: str) -> None:
    if extent < _state["world"]:
        raise ValueError(f"{what}: shard axis extent {extent} is smaller than the world size {_state['world']} "
                         "(every rank must own at least one row)")


def scatter_from_numpy(global_array):
    """Upload this rank's leading-axis block of a host array that every rank holds in full."""
    from . import core

    _check_rows(global_array.shape[0], "scatter_from_numpy")
    lo, hi = shard_bounds(global_array.shape[0], _state["rank"], _state["world"])
    return core.from_numpy(global_array[lo:hi])


def window_shard_from_numpy(base, window: int):
    """Upload the slice of a 1-D base buffer this rank needs for its share of the stride-1 sliding
    windows of length `window` (own outputs + halo), and return (array, n_local_outputs)."""
    from . import core

    n_out = base.shape[0] - window + 1
    _check_rows(n_out, "window_shard_from_numpy")
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


def allgather_ragged(shard, total_rows: int, bounds=None):
    """allgather for leading-axis blocks of unequal height (`bounds`, default shard_bounds): pad to the tallest
    block, gather, drop the padding."""
    from . import core

    world = _state["world"]
    if world == 1:
        return shard
    bounds = bounds or all_bounds(total_rows, world)
    heights = [hi - lo for lo, hi in bounds]
    if len(set(heights)) == 1:
        return allgather(shard)
    rest = tuple(shard.shape[1:])
    tallest = max(heights)
    padded = core.zeros((tallest,) + rest)
    padded[tuple([slice(0, shard.shape[0])] + [slice(0, e) for e in rest])] = shard
    stacked = allgather(padded)
    blocks = []
    for r in range(world):
        if heights[r] > 0:
            blocks.append(stacked[tuple([slice(r * tallest, r * tallest + heights[r])] + [slice(0, e) for e in rest])])
    return blocks[0] if len(blocks) == 1 else core.cat(blocks, [0])


def alltoall(send):
    """Block r (of `world` equal blocks of the flattened array) goes to rank r; returns the received blocks."""
    from . import core
    from ._lib import check_rc, lib

    src = send if send.stride == tuple(_row_major(send.shape)) else core.copy(send, deep=True)
    recv = core.zeros(src.shape)
    check_rc(lib.al_comm_alltoall(src.buffer, recv.buffer))
    return recv


_BINARY = {"add": "add", "sub": "sub", "mul": "mul", "truediv": "div", "pow": "pow", "lt": "lt", "le": "le", "gt": "gt",
           "ge": "ge", "eq": "eq", "ne": "ne"}
_REFLECTED = ("add", "sub", "mul", "truediv")
_UNARY = ("neg", "abs", "sqrt", "exp", "log", "sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh",
          "asinh", "acosh", "atanh", "ceil", "floor")


class ShardedArray:
    """A global array that is cut along its leading non-unit axis into contiguous blocks, one per rank ("sharded"),
    or that every rank holds in full ("replicated"). The `al.*` array operations are available as methods and
    operators with the semantics of the module docstring: everything is local except a reduction over the shard
    axis (one exchange), `as_strided` sliding windows (a halo from the next rank), and a transpose that moves the
    shard axis (one all-to-all)."""

    __hash__ = None

    def __init__(self, local, global_shape, placement: str, axis: int | None = None, bounds=None):
        assert placement in ("sharded", "replicated")
        self.local, self.shape, self.placement = local, tuple(global_shape), placement
        self.axis = shard_axis(self.shape) if axis is None else axis
        self.bounds = tuple(bounds) if bounds is not None else all_bounds(self.shape[self.axis], _state["world"])

    @property
    def ndim(self) -> int:
        return len(self.shape)

    # -- construction ------------------------------------------------------------------------------
    @classmethod
    def scatter(cls, global_array):
        """Every rank passes the same host array; each uploads only its block of the shard axis."""
        from . import core

        ax = shard_axis(global_array.shape)
        _check_rows(global_array.shape[ax], "ShardedArray.scatter")
        lo, hi = shard_bounds(global_array.shape[ax], _state["rank"], _state["world"])
        index = tuple([slice(None)] * ax + [slice(lo, hi)])
        return cls(core.from_numpy(global_array[index]), global_array.shape, "sharded", ax)

    @classmethod
    def replicate(cls, global_array):
        from . import core
        return cls(core.from_numpy(global_array), global_array.shape, "replicated")

    # -- helpers -----------------------------------------------------------------------------------
    def _my_bounds(self):
        return self.bounds[_state["rank"]]

    def _block_of_replica(self, out_axis_extent_is_one: bool):
        """The local operand to use inside an op whose output is sharded."""
        if self.placement == "sharded" or out_axis_extent_is_one:
            return self.local
        lo, hi = self._my_bounds()
        sl = tuple(slice(lo, hi) if i == self.axis else slice(0, e) for i, e in enumerate(self.shape))
        return self.local[sl]  # a view of this rank's rows of the replica

    def _like(self, local, shape=None, placement=None, axis=None, bounds=None):
        return type(self)(local, self.shape if shape is None else shape, placement or self.placement,
                            self.axis if axis is None else axis, self.bounds if bounds is None else bounds)

    def _binary(self, name: str, other, reverse: bool = False):
        from . import core
        fn = getattr(core, name)
        if isinstance(other, core.Number):
            res = fn(other, self.local) if reverse else fn(self.local, other)
            return self._like(res)
        if isinstance(other, core.NDArray):
            other = type(self)(other, other.shape, "replicated")
        assert isinstance(other, ShardedArray)
        out_shape = broadcast_shape(self.shape, other.shape)
        plan = binary_plan([(op.shape, op.placement, op.axis, op.bounds) for op in (self, other)], out_shape)
        locs = []
        for op, how in zip((self, other), plan["operands"]):
            if how == "block":          # cut the same way as the output: the local block is the operand
                locs.append(op.local)
            elif how == "whole":        # broadcast along the output's shard axis (or nothing is sharded)
                locs.append(op.gather())
            else:                       # ("rows", k): this rank's rows of a replica (gathered first if need be)
                lo, hi = plan["bounds"][_state["rank"]]
                whole = op.gather()
                locs.append(whole[tuple(slice(lo, hi) if i == how[1] else slice(0, e) for i, e in enumerate(op.shape))])
        a, b = locs
        res = fn(b, a) if reverse else fn(a, b)
        if plan["output"] == "replicated":
            return type(self)(res, out_shape, "replicated")
        return type(self)(res, out_shape, "sharded", plan["axis"], plan["bounds"])

    def _reduce(self, op: str, dims, replicated: bool = False):
        from . import core
        dims = tuple(range(self.ndim)) if dims is None else tuple(dims)
        out_shape = tuple(1 if d in dims else e for d, e in enumerate(self.shape))
        if self.placement == "replicated":
            return type(self)(getattr(core, f"reduce_{op}")(self.local, dims=dims), out_shape, "replicated")
        even = len({hi - lo for lo, hi in self.bounds}) == 1
        plan = reduce_plan(self.ndim, dims, self.axis, self.shape, _state["world"])
        if plan["collective"] == "reduce_scatter" and not replicated and even:
            return type(self)(reduce_sharded(self.local, op, dims, output="sharded", axis=self.axis), out_shape, "sharded",
                                plan["out_axis"])
        if plan["collective"]:
            return type(self)(reduce_sharded(self.local, op, dims, axis=self.axis), out_shape, "replicated")
        return type(self)(getattr(core, f"reduce_{op}")(self.local, dims=dims), out_shape, "sharded", self.axis, self.bounds)

    def reduce_sum(self, *, dims=None, replicated=False): return self._reduce("sum", dims, replicated)
    def reduce_max(self, *, dims=None, replicated=False): return self._reduce("max", dims, replicated)
    def reduce_min(self, *, dims=None, replicated=False): return self._reduce("min", dims, replicated)

    def apply(self, fn):
        """Elementwise map (a named device ufunc or a traceable arithmetic callable, see core.apply): local."""
        from . import core
        return self._like(core.apply(fn, self.local))

    def __neg__(self):
        from . import core
        return self._like(core.neg(self.local))

    # -- views ---------------------------------------------------------------------------------------
    def reshape(self, shape):
        """reshape of the GLOBAL array. Metadata-only on every rank when the block boundaries stay row boundaries of
        the new leading non-unit axis; otherwise the array is replicated first."""
        from . import core
        shape = tuple(shape)
        if self.placement == "replicated":
            return type(self)(core.reshape(self.local, shape), shape, "replicated")
        plan = reshape_plan(self.shape, self.bounds, shape)
        if plan["how"] == "gather":
            return type(self)(core.reshape(self.gather(), shape), shape, "replicated")
        lo, hi = plan["bounds"][_state["rank"]]
        local_shape = tuple(hi - lo if i == plan["axis"] else e for i, e in enumerate(shape))
        return type(self)(core.reshape(self.local, local_shape), shape, "sharded", plan["axis"], plan["bounds"])

    def as_strided(self, *, shape, stride):
        """Sliding windows over a sharded 1-D array: shape (n - w + 1, w), stride (1, 1) (config C5a). Rank r keeps
        the windows that start in its block and borrows the w - 1 elements that follow it from the next ranks (one
        small allgather of block heads); the result is sharded along axis 0. Replicated arrays take any layout."""
        from . import core
        shape, stride = tuple(shape), tuple(stride)
        if self.placement == "replicated":
            return type(self)(core.as_strided(self.local, shape, stride), shape, "replicated")
        n = self.shape[0]
        if not (self.ndim == 1 and len(shape) == 2 and stride == (1, 1) and shape[0] + shape[1] - 1 == n):
            raise NotImplementedError("as_strided of a sharded array supports stride-(1, 1) sliding windows over a 1-D "
                                      "array; gather() it for other layouts")
        w = shape[1]
        world, me = _state["world"], _state["rank"]
        view_bounds = window_view_bounds(self.bounds, w)
        lo, hi = self.bounds[me]
        if world == 1 or w == 1:
            return type(self)(core.as_strided(self.local, (hi - lo - w + 1, w), stride), shape, "sharded", 0, view_bounds)
        heights = [h - l for l, h in self.bounds]
        if min(heights) < w - 1:
            raise NotImplementedError("sliding windows wider than a rank's block: gather() the array first")
        heads = allgather(core.copy(self.local[(slice(0, w - 1),)], deep=True))  # (world * (w-1),): the first w-1 elements of every block
        rows = view_bounds[me][1] - view_bounds[me][0]
        if me == world - 1:
            base = self.local
        else:
            base = core.cat([self.local, heads[(slice((me + 1) * (w - 1), (me + 2) * (w - 1)),)]], [0])
        return type(self)(core.as_strided(base, (rows, w), stride), shape, "sharded", 0, view_bounds)

    # -- re-sharding ---------------------------------------------------------------------------------
    def transpose(self, perm):
        """transpose(perm) of the GLOBAL array; the result is sharded along its new leading axis. Local when
        axis 0 stays in front, otherwise one all-to-all (see reshard_plan)."""
        from . import core
        world, me = _state["world"], _state["rank"]
        new_shape = tuple(self.shape[p] for p in perm)
        if self.placement == "replicated":
            return type(self)(core.transpose(self.local, perm), new_shape, "replicated")
        even = self.bounds == all_bounds(self.shape[self.axis], world)
        if self.axis != 0 or not even:  # unusual cuts: replicate, transpose, keep the own block of the new shard axis
            whole = core.transpose(self.gather(), perm)
            ax = shard_axis(new_shape)
            lo, hi = shard_bounds(new_shape[ax], me, world)
            sl = tuple(slice(lo, hi) if i == ax else slice(0, e) for i, e in enumerate(new_shape))
            return type(self)(core.copy(whole[sl], deep=True), new_shape, "sharded", ax)
        plan = reshard_plan(self.shape, perm, world)
        if plan["how"] == "local":
            return type(self)(core.transpose(self.local, perm), plan["shape"], "sharded", 0)
        if plan["how"] == "gather":
            whole = core.transpose(allgather_ragged(self.local, self.shape[0]), perm)
            ax = shard_axis(plan["shape"])
            lo, hi = shard_bounds(plan["shape"][ax], me, world)
            sl = tuple(slice(lo, hi) if i == ax else slice(0, e) for i, e in enumerate(plan["shape"]))
            return type(self)(core.copy(whole[sl], deep=True), plan["shape"], "sharded", ax)
        piece = core.copy(core.transpose(self.local, perm), deep=True)   # (new leading axis in full, ..., own rows at q, ...)
        got = alltoall(piece)                                              # block i = rank i's rows of MY leading block
        mixed = core.transpose(core.reshape(got, plan["recv_view"]), plan["interleave"])
        return type(self)(core.reshape(mixed, plan["local_shape"]), plan["shape"], "sharded", 0)

    # -- materialisation ----------------------------------------------------------------------------
    def gather(self):
        """The full array on every rank (allgather; ragged blocks are padded for the exchange)."""
        from . import core
        if self.placement == "replicated" or _state["world"] == 1:
            return self.local
        ax = self.axis
        if ax == 0:
            return allgather_ragged(self.local, self.shape[0], self.bounds)
        # leading unit axes: the block is still one contiguous run; gather it as (rows, rest...) and relabel
        own = self.local
        if own.stride != _row_major(own.shape) or own.size != _count(own.shape):
            own = core.copy(own, deep=True)
        flat = core.reshape(own, (own.shape[ax],) + tuple(self.shape[ax + 1:]))
        whole = allgather_ragged(flat, self.shape[ax], self.bounds)
        if whole.size != _count(whole.shape):
            whole = core.copy(whole, deep=True)
        return core.reshape(whole, self.shape)

    def to_numpy(self):
        from . import core
        return core.to_numpy(self.gather())


def _attach_operators():
    for stem, fn in _BINARY.items():
        setattr(ShardedArray, f"__{stem}__", (lambda f: lambda self, o: self._binary(f, o))(fn))
    for stem in _REFLECTED:
        setattr(ShardedArray, f"__r{stem}__", (lambda f: lambda self, o: self._binary(f, o, True))(_BINARY[stem]))
    for name in _UNARY:
        if name != "neg":
            setattr(ShardedArray, name, (lambda n: lambda self: self.apply(n))(name))


_attach_operators()


def destroy() -> None:
    from ._lib import lib

    lib.al_comm_destroy()
    if _state["rdzv"] and _state["rank"] == 0:
        import shutil
        shutil.rmtree(_state["rdzv"], ignore_errors=True)
    _state.update(rank=0, world=1, rdzv=None)
