"""`import arraylib as al` on N GPUs: the reference's Python API over leading-axis shards, one process per GPU.

The reference is single-process and single-device; its API is `core.py:46-802`. With

    ARRAYLIB_DEVICES=8 python script.py            # the shim re-launches the script once per GPU
    python -m arraylib_b200.launch --nproc 8 script.py      (or torchrun)

every rank runs the same script (SPMD) and `al.*` names resolve to this module: arrays are GLOBAL objects
(`NDArray` below, a `dist.ShardedArray`) whose `shape`, indexing, operators, `reduce_*`, `reshape`, `transpose`,
`as_strided`, `apply`, `to_numpy` ... mean what they mean in the reference, while the data of an array that is
large enough (`ARRAYLIB_SHARD_MIN_BYTES`, default 1 MiB) lives as contiguous blocks of its leading non-unit
axis, one block per GPU. What runs where:

  * creation (`ones/zeros/full/arange/from_numpy/array`): every rank materialises only its block;
  * arithmetic, comparisons, scalars, `apply`, `where`: local kernels on the blocks; an operand that is
    broadcast along the shard axis is replicated (small by construction);
  * `reduce_*` keeping the shard axis: local. Reducing it: local partial + ONE exchange over NVLink peer memory
    (reduce-scatter when the result can stay sharded, else allreduce) -- `arraylib.c:732-750`'s merge of
    per-thread accumulators is the nearest thing in the reference;
  * `as_strided` sliding windows: a (window-1)-element halo from the next rank; `transpose` that moves the shard
    axis: one all-to-all; `reshape`: metadata when block boundaries stay row boundaries;
  * everything else (`matmul`, `dot`, `cat`, `move_dim`, `ravel`, partial-range views, layouts the planner
    does not cover): the operands are gathered and every rank computes the same replicated result -- always
    correct, never sharded. Small arrays are replicated from the start, so scripts written for the reference
    (its own tests/test_array.py included) run unchanged.

Deviations: a slice view that cuts the shard axis is a gathered COPY (writes through it do not reach the parent);
`size` / `len()` of a sharded array report the global element count.
"""
from __future__ import annotations

import builtins
import os
import typing as tp

from . import core, dist
from .core import Layout, Number, cdiv, normalize_index, slice_to_range  # noqa: F401
from .dist import ShardedArray, _count, all_bounds, shard_axis, shard_bounds

SHARD_MIN_BYTES = int(os.environ.get("ARRAYLIB_SHARD_MIN_BYTES", str(1 << 20)))


# ---- pure planning (CPU-testable) ------------------------------------------------------------------------


def placement_for(shape, world: int, min_bytes: int | None = None) -> str:
    """'sharded' when a new float32 array of `shape` is worth cutting over `world` ranks, else 'replicated'."""
    min_bytes = SHARD_MIN_BYTES if min_bytes is None else min_bytes
    if world <= 1:
        return "replicated"
    ax = shard_axis(shape)
    return "sharded" if 4 * _count(shape) >= min_bytes and shape[ax] >= world else "replicated"


def block_index(shape, axis: int, lo: int, hi: int) -> tuple:
    """Full-rank tuple of slices selecting rows [lo, hi) of `axis`."""
    return tuple(slice(lo, hi) if i == axis else slice(0, e) for i, e in enumerate(shape))


def strided_rows_in_block(start: int, end: int, step: int, lo: int, hi: int):
    """Rows start, start+step, ... < end that fall into [lo, hi): (k0, k1, local_start, local_end) with k the
    position in the strided sequence and the local_* bounds relative to `lo`; None when the block holds none."""
    count = cdiv(end - start, step)
    k0 = builtins.max(0, cdiv(lo - start, step))  # (this module defines al.max / al.min)
    k1 = builtins.min(count, cdiv(hi - start, step))
    if k1 <= k0:
        return None
    first, last = start + k0 * step, start + (k1 - 1) * step
    return k0, k1, first - lo, last - lo + 1


def owner_of(row: int, bounds) -> int:
    for r, (lo, hi) in enumerate(bounds):
        if lo <= row < hi:
            return r
    raise IndexError(f"row {row} is outside {bounds}")


# ---- the global array ------------------------------------------------------------------------------------


class NDArray(ShardedArray):
    """Global float32 array with the reference's NDArray interface (core.py:46-199)."""

    __hash__ = None

    # -- metadata ------------------------------------------------------------------------------------
    @property
    def stride(self) -> tuple:
        return self.local.stride  # blocks are cut along the leading non-unit axis: same element strides

    @property
    def size(self) -> int:
        if self.placement == "replicated" or self.local.size != _count(self.local.shape):
            return self.local.size
        return _count(self.shape)

    @property
    def view(self) -> bool:
        return self.local.view

    @property
    def layout(self) -> Layout:
        return Layout(self.shape, self.stride, self.ndim)

    @property
    def T(self):
        return self.transpose(list(range(self.ndim - 1, -1, -1)))

    def __len__(self) -> int:
        return self.size

    def __float__(self) -> float:
        assert self.size == 1
        return float(self.gather())

    def __int__(self) -> int:
        return int(float(self))

    def __array__(self, dtype=None, copy=None):
        out = self.to_numpy()
        return out if dtype is None else out.astype(dtype)

    def __repr__(self) -> str:
        return f"NDArray(f32[{','.join(map(str, self.shape))}])"

    def __str__(self) -> str:
        return core.ppstr(self.gather())

    def __copy__(self):
        return copy(self)

    def __deepcopy__(self, memo):
        return copy(self, deep=True)

    # -- host materialisation ----------------------------------------------------------------------------
    def to_numpy(self, out=None, blocking=True):
        return core.to_numpy(self.gather(), out=out, blocking=blocking)

    # -- indexing (core.py:514-641) ------------------------------------------------------------------------
    def __getitem__(self, index):
        index = normalize_index(index)
        assert len(index) == self.ndim, f"{len(index)} != {self.ndim}"
        if self.placement == "replicated" or dist.world() == 1:
            got = core.getitem(self.local, index)
            return got if isinstance(got, float) else type(self)(got, got.shape, "replicated")
        if all(isinstance(i, int) for i in index):
            assert all(i < e for i, e in zip(index, self.shape))
            owner = owner_of(index[self.axis], self.bounds)
            value = 0.0
            if owner == dist.rank():
                lo = self.bounds[owner][0]
                value = core.getitem(self.local, tuple(i - lo if d == self.axis else i for d, i in enumerate(index)))
            return dist.broadcast_scalar(value, owner)
        if all(isinstance(i, slice) for i in index):
            start, end, step = slice_to_range(self.shape, self.ndim, index)
            ax = self.axis
            if (start[ax], end[ax], step[ax]) == (0, self.shape[ax], 1):  # the shard axis is taken whole: still sharded
                lo, hi = self._my_bounds()
                local = self.local[tuple(slice(0, hi - lo) if d == ax else s for d, s in enumerate(index))]
                shape = tuple(cdiv(e - s, st) for s, e, st in zip(start, end, step))
                if shard_axis(shape) == ax:
                    return type(self)(local, shape, "sharded", ax, self.bounds)
            whole = core.getitem(self.gather(), index)  # a gathered COPY (module docstring)
            return type(self)(whole, whole.shape, "replicated")
        raise NotImplementedError(f"Index type not supported: {index}")

    def __setitem__(self, index, value):
        index = normalize_index(index)
        assert len(index) == self.ndim, "full index required to match ndim"
        if isinstance(value, ShardedArray):
            assert self.ndim == value.ndim
        if self.placement == "replicated" or dist.world() == 1:
            core.setitem(self.local, index, value.gather() if isinstance(value, ShardedArray) else value)
            return self
        ax = self.axis
        lo, hi = self._my_bounds()
        if all(isinstance(i, int) for i in index):
            assert isinstance(value, Number), f"{type(value)=}"
            assert all(i < e for i, e in zip(index, self.shape))
            if lo <= index[ax] < hi:
                core.setitem(self.local, tuple(i - lo if d == ax else i for d, i in enumerate(index)), value)
            return self
        if all(isinstance(i, slice) for i in index):
            start, end, step = slice_to_range(self.shape, self.ndim, index)
            if isinstance(value, ShardedArray):
                assert all(cdiv(end[i] - start[i], step[i]) == value.shape[i] for i in range(self.ndim))
                value = value.gather()  # every rank then cuts the rows that land in its own block
            elif not isinstance(value, Number):
                raise NotImplementedError(f"Value type not supported: {value}")
            rows = strided_rows_in_block(start[ax], end[ax], step[ax], lo, hi)
            if rows is None:
                return self
            k0, k1, l0, l1 = rows
            mine = tuple(slice(l0, l1, step[ax]) if d == ax else s for d, s in enumerate(index))
            if not isinstance(value, Number):
                value = value[block_index(value.shape, ax, k0, k1)]
            core.setitem(self.local, mine, value)
            return self
        raise NotImplementedError(f"Index type not supported: {index}")

    # -- methods of the reference's NDArray that have no sharded form: gather, compute everywhere ------------
    def move_dim(self, src, dst):
        return move_dim(self, src, dst)

    def ravel(self):
        return ravel(self)

    def where(self, lhs, rhs):
        return where(self, lhs, rhs)

    def cat(self, arrays, *, dims):
        return cat([self] + list(arrays), dims)

    def __matmul__(self, other):
        return matmul(self, other)


# ---- helpers -------------------------------------------------------------------------------------------------


def _replicated(local) -> NDArray:
    return NDArray(local, local.shape, "replicated")


def _local(x):
    """The whole array on this rank (gathers a sharded one); numbers pass through."""
    return x.gather() if isinstance(x, ShardedArray) else x


def _maybe_shard(whole) -> NDArray:
    """A replicated local array, cut down to this rank's block when the placement policy says so."""
    shape = whole.shape
    world = dist.world()
    if placement_for(shape, world) == "replicated":
        return _replicated(whole)
    ax = shard_axis(shape)
    lo, hi = shard_bounds(shape[ax], dist.rank(), world)
    return NDArray(core.copy(whole[block_index(shape, ax, lo, hi)], deep=True), shape, "sharded", ax)


def _everywhere(fn):
    """Wrap a single-device function: gather the array arguments, run on every rank, return a replicated array."""

    def wrapper(*args, **kwargs):
        args = [[_local(a) for a in arg] if isinstance(arg, (list, tuple)) and any(isinstance(a, ShardedArray) for a in arg)
                else _local(arg) for arg in args]
        out = fn(*args, **{k: _local(v) for k, v in kwargs.items()})
        return _replicated(out) if isinstance(out, core.NDArray) else out

    wrapper.__name__, wrapper.__doc__ = fn.__name__, fn.__doc__
    return wrapper


# ---- creation (core.py:207-363) ----------------------------------------------------------------------------


def full(shape, value) -> NDArray:
    shape = tuple(core._int_seq(shape, "shape"))
    world = dist.world()
    if placement_for(shape, world) == "replicated":
        return _replicated(core.full(shape, value))
    ax = shard_axis(shape)
    lo, hi = shard_bounds(shape[ax], dist.rank(), world)
    return NDArray(core.full(tuple(hi - lo if d == ax else e for d, e in enumerate(shape)), value), shape, "sharded", ax)


def ones(shape) -> NDArray:
    return full(shape, 1.0)


def zeros(shape) -> NDArray:
    return full(shape, 0.0)


def arange(start: int, end: tp.Optional[int] = None, step: int = 1) -> NDArray:
    if end is None:
        start, end = 0, start
    assert isinstance(start, int) and isinstance(end, int) and isinstance(step, int)
    assert (0 <= start < end) and (step > 0)
    n = cdiv(end - start, step)
    world = dist.world()
    if placement_for((n,), world) == "sharded" and start + n * step <= 1 << 24:
        # every value is exact in float32, so a block is an arange of its own (arraylib.c:298-308 is a recurrence
        # only in form below 2^24)
        lo, hi = shard_bounds(n, dist.rank(), world)
        return NDArray(core.arange(start + lo * step, start + hi * step, step), (n,), "sharded", 0)
    return _maybe_shard(core.arange(start, end, step))  # past 2^24 the f32 recurrence is replayed in full, then cut


def linspace(start: int, end: int, n: int) -> NDArray:
    return _maybe_shard(core.linspace(start, end, n))


def array(elems) -> NDArray:
    return _maybe_shard(core.array(elems))


def from_numpy(array, blocking=True) -> NDArray:
    """Every rank passes the same host array and uploads only its block (or all of it when the array is small)."""
    import numpy as np

    assert isinstance(array, np.ndarray), f"expected numpy array, got {type(array)=}"
    shape = array.shape if array.ndim else (1,)
    world = dist.world()
    if array.ndim == 0 or placement_for(shape, world) == "replicated":
        return _replicated(core.from_numpy(array, blocking=blocking))
    ax = shard_axis(shape)
    lo, hi = shard_bounds(shape[ax], dist.rank(), world)
    return NDArray(core.from_numpy(array[block_index(shape, ax, lo, hi)], blocking=blocking), shape, "sharded", ax)


def to_numpy(array, out=None, blocking=True):
    assert isinstance(array, ShardedArray)
    return core.to_numpy(array.gather(), out=out, blocking=blocking)


host_buffer = core.host_buffer
synchronize = core.synchronize


def free(array) -> None:
    if isinstance(array, ShardedArray):
        core.free(array.local)


# ---- operators and functions -----------------------------------------------------------------------------------


def _binary_function(name: str):
    fn = getattr(core, name)

    def wrapper(lhs, rhs):
        if isinstance(lhs, ShardedArray):
            return lhs._binary(name, rhs)
        if isinstance(rhs, ShardedArray):
            return rhs._binary(name, lhs, reverse=True)
        return fn(lhs, rhs)

    wrapper.__name__, wrapper.__doc__ = fn.__name__, fn.__doc__
    return wrapper


for _name in ("add", "sub", "mul", "div", "mod", "pow", "max", "min", "eq", "ne", "gt", "ge", "lt", "le"):
    globals()[_name] = _binary_function(_name)
for _stem, _name in (("mod", "mod"),):
    setattr(NDArray, f"__{_stem}__", (lambda f: lambda self, o: self._binary(f, o))(_name))
for _name in dist._UNARY:
    globals()[_name] = (lambda n: lambda array: array.apply(n))(_name)
del _name, _stem


def apply(fn, array) -> NDArray:
    assert isinstance(array, ShardedArray)
    return array.apply(fn)


def reshape(array, shape) -> NDArray:
    shape = tuple(core._int_seq(shape, "shape"))
    assert _count(array.shape) == _count(shape), f"incorrect reshape {shape=}!={array.shape=}"
    return array.reshape(shape)


def transpose(array, dst) -> NDArray:
    dst = core._int_seq(dst, "dst")
    assert len(dst) == array.ndim
    return array.transpose(tuple(dst))


def as_strided(array, shape, stride) -> NDArray:
    try:
        return array.as_strided(shape=shape, stride=stride)
    except NotImplementedError:  # layouts other than 1-D sliding windows: view of the gathered array
        return _replicated(core.as_strided(array.gather(), shape, stride))


def copy(array, deep=False) -> NDArray:
    return type(array)(core.copy(array.local, deep=deep), array.shape, array.placement, array.axis, array.bounds)


def _reduce_function(op: str):
    def wrapper(array, dims=None):
        assert isinstance(array, ShardedArray)
        return array._reduce(op, dims)

    wrapper.__name__ = f"reduce_{op}"
    return wrapper


reduce_sum, reduce_max, reduce_min = (_reduce_function(op) for op in ("sum", "max", "min"))


def reduce(fn, array, dims=None, init=0.0) -> NDArray:
    """Keepdims reduction with an accumulator that must resolve to sum / max / min / prod (core.py:661-673). The
    sharded exchange folds `init` in once, like the single-device kernel; prod and a custom init on a sharded axis
    take the gathered route."""
    from . import ufuncs
    from ._lib import REDOPS

    op = REDOPS[ufuncs.resolve_reduce(fn)]
    identity = {"sum": 0.0, "max": float("-inf"), "min": float("inf"), "prod": 1.0}[op]
    dims_t = tuple(range(array.ndim)) if dims is None else tuple(dims)
    crosses = array.placement == "sharded" and array.axis in dims_t and dist.world() > 1
    if crosses and (op == "prod" or float(init) != identity):
        return _replicated(core.reduce(fn, array.gather(), dims=dims, init=init))
    if not crosses:
        out = core.reduce(fn, array.local, dims=dims, init=init)
        shape = tuple(1 if d in dims_t else e for d, e in enumerate(array.shape))
        if array.placement == "replicated":
            return _replicated(out)
        return NDArray(out, shape, "sharded", array.axis, array.bounds)
    return array._reduce(op, dims)


def where(cond, lhs, rhs) -> NDArray:
    """Elementwise select. Local on the blocks when every array operand is cut the same way."""
    arrays = [x for x in (cond, lhs, rhs) if isinstance(x, ShardedArray)]
    same = all(a.shape == cond.shape and a.placement == cond.placement and a.bounds == cond.bounds and a.axis == cond.axis
               for a in arrays)
    if same:
        loc = [x.local if isinstance(x, ShardedArray) else x for x in (cond, lhs, rhs)]
        return type(cond)(core.where(*loc), cond.shape, cond.placement, cond.axis, cond.bounds)
    return _replicated(core.where(_local(cond), _local(lhs), _local(rhs)))


matmul = _everywhere(core.matmul)
dot = _everywhere(core.dot)
cat = _everywhere(core.cat)
move_dim = _everywhere(core.move_dim)
ravel = _everywhere(core.ravel)
get_view_from_range = _everywhere(core.get_view_from_range)
is_broadcastable = core.is_broadcastable


def getitem(array, index):
    return array[index]


def setitem(array, index, value):
    array[index] = value
    return array


def pprepr(array) -> str:
    return repr(array)


def ppstr(array) -> str:
    return str(array)


def lazy(array):
    """Opt-in fusion works on the local blocks: `al.lazy(A.local)`. Global arrays evaluate eagerly."""
    raise NotImplementedError("al.lazy() takes a single-device array; use it on `array.local` inside an SPMD script")


def enabled() -> bool:
    return dist.world() > 1
