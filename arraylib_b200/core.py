"""Python face of the B200 array core: the same `al.*` names and semantics as the reference's
arraylib/core.py, bound with ctypes to libarraylib_b200.so (device-resident float32 buffers).

Reference citations (/root/reference/arraylib/core.py) are given per function. Differences from
the reference that a caller can observe are listed in DESIGN.md ("Deviations"); in short:
`to_numpy` is an owning device->host copy of the logical view, `apply`/`reduce` callables must
resolve to a named device kernel (else NotImplementedError), scalar-on-the-left arithmetic works,
and C-level failures raise Python exceptions instead of aborting the interpreter.
"""
from __future__ import annotations

import ctypes as C
import functools as ft
import typing as tp
from io import StringIO

from . import ufuncs as _uf
from ._lib import (BINOPS, REDOPS, UFUNCS, ArrayP, CLayout, check_ptr, check_rc, lib, raise_last_error,
                   size_array)

Number = (int, float)


class Layout(tp.NamedTuple):  # core.py:40-43
    shape: tuple
    stride: tuple
    ndim: int


# == NDArray (core.py:46-199) ==


class NDArray:
    """Float32 N-d array whose buffer lives in GPU HBM. `buffer` is the C `NDArray*` handle."""

    __slots__ = ("buffer", "__weakref__")

    def __init__(self, buffer):
        self.buffer = check_ptr(buffer)

    # -- metadata, read straight from the host-side C structs ------------------------------------
    @property
    def ndim(self) -> int:
        return int(self.buffer.contents.lay.contents.ndim)

    @property
    def shape(self) -> tuple:
        lay = self.buffer.contents.lay.contents
        return tuple(int(lay.shape[i]) for i in range(lay.ndim))

    @property
    def stride(self) -> tuple:
        lay = self.buffer.contents.lay.contents
        return tuple(int(lay.stride[i]) for i in range(lay.ndim))

    @property
    def size(self) -> int:  # element count of the BASE buffer, like core.py:61-62
        return int(self.buffer.contents.data.contents.size)

    @property
    def view(self) -> bool:
        return bool(self.buffer.contents.view)

    @property
    def layout(self) -> Layout:
        return Layout(self.shape, self.stride, self.ndim)

    @property
    def T(self):
        return self.transpose(list(range(self.ndim - 1, -1, -1)))

    # Methods and operators that simply forward to the module-level functions are attached below the
    # function definitions from two tables (_METHOD_TABLE, _OPERATOR_TABLE).

    __hash__ = None

    def __getitem__(self, index):
        return getitem(self, index)

    def __setitem__(self, index, value):
        return setitem(self, index, value)

    def __copy__(self):
        return copy(self)

    def __deepcopy__(self, memo):
        return copy(self, deep=True)

    def __float__(self) -> float:  # core.py:187-189: element 0 of the base buffer
        assert self.size == 1
        out = C.c_float()
        check_rc(lib.array_read_base(self.buffer, 0, 1, C.byref(out)))
        return float(out.value)

    def __int__(self) -> int:
        return int(float(self))

    def __len__(self) -> int:
        return self.size

    def __array__(self, dtype=None, copy=None):
        out = to_numpy(self)
        return out if dtype is None else out.astype(dtype)

    def __str__(self) -> str:
        return ppstr(self)

    def __repr__(self) -> str:
        return pprepr(self)

    def __del__(self):
        try:
            free(self)
        except Exception:  # interpreter shutdown: module globals may already be gone
            pass


def _new(ptr) -> NDArray:
    return NDArray(ptr)


def free(array) -> None:  # core.py:330-332 (idempotent here)
    buf = getattr(array, "buffer", None)
    if buf:
        array.buffer = None
        try:
            lib.array_free(buf)
        except Exception:  # interpreter shutdown
            pass


def synchronize() -> None:
    """Block until every enqueued kernel has finished (the library is asynchronous)."""
    check_rc(lib.al_sync())


# == helpers (core.py:207-244) ==


def cdiv(a, b):
    return (a + b - 1) // b


def normalize_index(index):
    return tuple(index) if isinstance(index, tp.Sequence) else (index,)


def slice_to_range(shape, ndim, slices):
    assert isinstance(slices, tp.Sequence)
    assert all(isinstance(s, slice) for s in slices)
    assert len(slices) == ndim
    start = tuple(s.start or 0 for s in slices)
    end = tuple(s.stop or shape[i] for i, s in enumerate(slices))
    step = tuple(s.step or 1 for s in slices)
    assert all(0 <= lo < hi <= shape[i] for i, (lo, hi) in enumerate(zip(start, end)))
    return start, end, step


@ft.cache
def is_broadcastable(lhs: Layout, rhs: Layout) -> bool:
    for le_, re_ in zip(reversed(lhs.shape), reversed(rhs.shape)):
        if le_ != re_ and le_ != 1 and re_ != 1:
            return False
    return True


def _int_seq(seq, what):
    assert isinstance(seq, tp.Sequence), f"{what} must be a sequence"
    assert all(isinstance(i, int) for i in seq), f"{what} must hold ints"
    return tuple(seq)


# == creation (core.py:252-327) ==


def array(elems) -> NDArray:
    assert isinstance(elems, tp.Sequence)
    extents: dict = {}
    flat: list = []

    def walk(node, depth):
        if not isinstance(node, tp.Sequence):
            assert isinstance(node, Number), f"elems={node!r}"
            flat.append(float(node))
            return
        if depth in extents:
            assert extents[depth] == len(node), f"dim={depth} mismatch: {extents[depth]}!={len(node)}"
        else:
            extents[depth] = len(node)
        for child in node:
            walk(child, depth + 1)

    walk(elems, 0)
    shape = tuple(extents.values())
    host = (C.c_float * len(flat))(*flat)
    return _new(lib.array_fill(host, size_array(shape), len(shape)))


def ones(shape) -> NDArray:
    shape = _int_seq(shape, "shape")
    return _new(lib.array_ones(size_array(shape), len(shape)))


def zeros(shape) -> NDArray:
    shape = _int_seq(shape, "shape")
    return _new(lib.array_zeros(size_array(shape), len(shape)))


def full(shape, value) -> NDArray:
    shape = _int_seq(shape, "shape")
    return _new(lib.array_full(size_array(shape), len(shape), float(value)))


def arange(start: int, end: tp.Optional[int] = None, step: int = 1) -> NDArray:
    if end is None:
        start, end = 0, start
    assert isinstance(start, int) and isinstance(end, int) and isinstance(step, int)
    assert (0 <= start < end) and (step > 0)
    return _new(lib.array_arange(start, end, step))


def linspace(start: int, end: int, n: int) -> NDArray:
    assert isinstance(start, int) and isinstance(end, int) and isinstance(n, int)
    assert (0 <= start < end) and (n > 0)
    return _new(lib.array_linspace(start, end, n))


# == host <-> device (core.py:340-363) ==


def from_numpy(array, blocking=True) -> NDArray:
    """Upload a float32 numpy array (any strides; 0-d becomes shape (1,)).

    `blocking=False` (extension): `array` must be C-contiguous PINNED memory (`host_buffer`) that is
    left untouched until `synchronize()`; the copy then runs on the library's H2D stream."""
    import numpy as np

    assert isinstance(array, np.ndarray), f"expected numpy array, got {type(array)=}"
    assert array.dtype == np.float32, f"expected float32 dtype, got {array.dtype=}"
    host = np.ascontiguousarray(array)
    shape = host.shape if host.ndim else (1,)
    if not blocking:
        assert host is array or host.ctypes.data == array.ctypes.data, "blocking=False needs a C-contiguous array"
        return _new(lib.array_fill_async(C.c_void_p(host.ctypes.data), size_array(shape), len(shape)))
    return _new(lib.array_fill(C.c_void_p(host.ctypes.data), size_array(shape), len(shape)))


def to_numpy(array: NDArray, out=None, blocking=True):
    """Device -> host COPY of the logical view (honours strides and the view offset).

    `out` (optional, extension): a C-contiguous float32 array of the same shape to copy into,
    e.g. one backed by pinned memory from `host_buffer`. With `blocking=False` (needs a pinned
    `out`) the copy is only enqueued on the library's D2H stream; `out` is valid after
    `synchronize()`."""
    import numpy as np

    assert isinstance(array, NDArray)
    if out is None:
        out = np.empty(array.shape, dtype=np.float32)
    else:
        assert isinstance(out, np.ndarray) and out.dtype == np.float32 and out.flags.c_contiguous
        assert out.shape == array.shape
    if not blocking:
        check_rc(lib.array_to_host_async(array.buffer, C.c_void_p(out.ctypes.data)))
        return out
    check_rc(lib.array_to_host(array.buffer, C.c_void_p(out.ctypes.data)))
    return out


def host_buffer(shape):
    """float32 numpy array backed by PINNED host memory (extension; staging for fast or non-blocking
    from_numpy / to_numpy). The memory is released when the array and all its views are collected."""
    import weakref

    import numpy as np

    shape = tuple(int(e) for e in shape)
    n = int(np.prod(shape)) if shape else 1
    n_alloc = n if n > 0 else 1  # (this module rebinds max/min/abs/pow to array ops)
    ptr = lib.al_host_alloc(n_alloc * 4)
    if not ptr:
        raise_last_error()
    raw = (C.c_float * n_alloc).from_address(ptr)
    weakref.finalize(raw, lib.al_host_free, ptr)  # every numpy view keeps `raw` alive through .base
    return np.ctypeslib.as_array(raw)[:n].reshape(shape)


# == binary ops (core.py:371-414) ==

_COMMUTATIVE = {"sum", "mul", "max", "min", "eq", "ne"}
_MIRROR = {"gt": "lt", "lt": "gt", "ge": "le", "le": "ge"}


def generate_binary_op(op: str):
    code = BINOPS.index(op)

    def wrapper(lhs, rhs):
        if isinstance(lhs, NDArray) and isinstance(rhs, NDArray):
            assert is_broadcastable(lhs.layout, rhs.layout), "cannot broadcast shapes"
            return _new(lib.array_array_op_named(code, lhs.buffer, rhs.buffer))
        if isinstance(lhs, NDArray) and isinstance(rhs, Number):
            return _new(lib.array_scalar_op_named(code, lhs.buffer, float(rhs)))
        if isinstance(lhs, Number) and isinstance(rhs, NDArray):
            # The reference raises NotImplementedError here (core.py:381); supported as
            # a superset: swap when the op allows it, else broadcast a one-element array.
            if op in _COMMUTATIVE:
                return wrapper(rhs, lhs)
            if op in _MIRROR:
                return globals()[_MIRROR[op]](rhs, lhs)
            return wrapper(full((1,), lhs), rhs)
        raise NotImplementedError(f"Not supported for {type(lhs)=} and {type(rhs)=}")

    wrapper.__doc__ = f"Binary {op} operation"
    wrapper.__name__ = f"array_array_{op}"
    return wrapper


# module-level names follow the reference (core.py:388-400): `add` is the op called "sum" in the C table
for _op in BINOPS:
    globals()["add" if _op == "sum" else _op] = generate_binary_op(_op)
del _op


def matmul(lhs, rhs) -> NDArray:
    assert lhs.shape[-1] == rhs.shape[0]
    assert lhs.ndim == 2 and rhs.ndim == 2
    return _new(lib.array_array_matmul(lhs.buffer, rhs.buffer))


def dot(lhs, rhs) -> NDArray:
    assert isinstance(lhs, NDArray) and isinstance(rhs, NDArray)
    assert lhs.ndim == 1 and rhs.ndim == 1, "dot product requires 1D arrays"
    return _new(lib.array_array_dot(lhs.buffer, rhs.buffer))


# == unary ufunc table and apply (core.py:422-460) ==


def generate_unary_op(op: str):
    code = UFUNCS.index(op)

    def wrapper(array):
        assert isinstance(array, NDArray)
        return _new(lib.array_op_named(code, array.buffer))

    wrapper.__doc__ = f"Elementwise {op} operation"
    wrapper.__name__ = f"array_{op}"
    return wrapper


for _op in UFUNCS:  # neg abs sqrt exp log sin ... floor (core.py:434-452)
    globals()[_op] = generate_unary_op(_op)
del _op


def apply(fn, array) -> NDArray:
    """Elementwise map on the device. `fn` is resolved, in this order, to
      1. a named device ufunc by identity / name (math.sqrt, np.log, abs, ...);
      2. a traced FP64 tape: arithmetic lambdas such as `lambda x: 2 * x * x + 1`, `lambda x: (x + 1) / (x - 1)`,
         `lambda x: np.exp(-x * x)` or `lambda x: math.sqrt(x) + 1` are recorded on a scalar proxy and run as one
         kernel in double precision with one final rounding (the arithmetic of the reference's per-element
         Python callback); a trace of exactly one ufunc (`lambda x: math.sqrt(x)`) runs the named kernel.
    Anything else (data-dependent branches, builtin max/min on the value, float(x)) raises NotImplementedError --
    there is no CPU fallback and no guessing from sample values."""
    assert callable(fn) or isinstance(fn, str)
    assert isinstance(array, NDArray)
    code = _uf.resolve_by_name(fn)
    if code is None and not isinstance(fn, str):
        from .lazy import single_instruction, trace_unary  # (the package also exports a FUNCTION named `lazy`)
        one = single_instruction(fn, 1)
        if one is not None and one[0] == 1:
            code = UFUNCS.index(one[1])
        else:
            traced = trace_unary(fn, array)
            if traced is not None:
                return traced
    if code is None:
        code = _uf.resolve_unary(fn)  # raises with the explanation
    return _new(lib.array_op_named(code, array.buffer))


# == views (core.py:468-506, 788-802) ==


def reshape(array, shape) -> NDArray:
    shape = _int_seq(shape, "shape")
    dst_total = ft.reduce(lambda x, y: x * y, shape, 1)
    src_total = ft.reduce(lambda x, y: x * y, array.shape, 1)
    assert src_total == dst_total, f"incorrect reshape {shape=}!={array.shape=}"
    return _new(lib.array_reshape(array.buffer, size_array(shape), len(shape)))


def transpose(array, dst) -> NDArray:
    dst = _int_seq(dst, "dst")
    assert len(dst) == array.ndim
    return _new(lib.array_transpose(array.buffer, size_array(dst)))


def move_dim(array, src, dst) -> NDArray:
    assert isinstance(array, NDArray)
    src, dst = _int_seq(src, "src"), _int_seq(dst, "dst")
    assert len(src) == len(dst)
    assert 0 < len(src) <= array.ndim
    return _new(lib.array_move_dim(array.buffer, size_array(src), size_array(dst), len(src)))


def ravel(array) -> NDArray:
    assert isinstance(array, NDArray)
    return _new(lib.array_ravel(array.buffer))


def as_strided(array, shape, stride) -> NDArray:
    assert isinstance(array, NDArray)
    shape, stride = _int_seq(shape, "shape"), _int_seq(stride, "stride")
    assert len(shape) == len(stride)
    assert all(s > 0 for s in stride)
    sh, st = size_array(shape), size_array(stride)
    lay = CLayout(len(shape), C.cast(sh, C.POINTER(C.c_size_t)), C.cast(st, C.POINTER(C.c_size_t)))
    return _new(lib.array_as_strided(array.buffer, C.byref(lay)))


# == indexing (core.py:514-641) ==


def _range_args(array, rng):
    start, end, step = rng
    for part in (start, end, step):
        assert isinstance(part, tp.Sequence)
        assert all(isinstance(i, int) for i in part)
    assert len(start) == len(end) == len(step) == array.ndim
    return size_array(start), size_array(end), size_array(step)


def get_view_from_range(array, range) -> NDArray:
    assert isinstance(array, NDArray)
    return _new(lib.array_get_view(array.buffer, *_range_args(array, range)))


def get_scalar_from_index(array, index) -> float:
    assert isinstance(array, NDArray)
    index = _int_seq(index, "index")
    assert len(index) == array.ndim
    assert all(idx < array.shape[i] for i, idx in enumerate(index))
    value = lib.array_get_elem(array.buffer, size_array(index))
    if value != value and lib.al_last_error():
        raise_last_error()
    return value


def getitem(array, index):
    index = normalize_index(index)
    assert len(index) == array.ndim, f"{len(index)} != {array.ndim}"
    if all(isinstance(i, int) for i in index):
        return get_scalar_from_index(array, index)
    if all(isinstance(i, slice) for i in index):
        return get_view_from_range(array, slice_to_range(array.shape, array.ndim, index))
    raise NotImplementedError(f"Index type not supported: {index}")


def set_elem_from_scalar(array, value, index):
    index = _int_seq(index, "index")
    assert len(index) == array.ndim
    assert all(idx < array.shape[i] for i, idx in enumerate(index))
    check_ptr(lib.array_set_elem_from_scalar(array.buffer, float(value), size_array(index)))
    return array


def set_view_from_scalar(array, value, index):
    start, end, step = _range_args(array, index)
    shape = array.shape
    assert all(0 <= index[0][i] < index[1][i] <= shape[i] for i in range(array.ndim))
    check_ptr(lib.array_set_view_from_scalar(array.buffer, float(value), start, end, step))
    return array


def set_view_from_array(array, value: NDArray, index):
    assert array.ndim == value.ndim
    start, end, step = index
    assert len(start) == len(end) == len(step) == value.ndim
    assert all(cdiv(end[i] - start[i], step[i]) == value.shape[i] for i in range(array.ndim))
    check_ptr(lib.array_set_view_from_array(array.buffer, value.buffer, *_range_args(array, index)))
    return array


def setitem(array, index, value):
    assert isinstance(array, NDArray)
    index = normalize_index(index)
    assert len(index) == array.ndim, "full index required to match ndim"
    if all(isinstance(i, int) for i in index):
        assert isinstance(value, Number), f"{type(value)=}"
        return set_elem_from_scalar(array, value, index)
    if all(isinstance(i, slice) for i in index):
        rng = slice_to_range(array.shape, array.ndim, index)
        if isinstance(value, Number):
            return set_view_from_scalar(array, value, rng)
        if isinstance(value, NDArray):
            return set_view_from_array(array, value, rng)
        raise NotImplementedError(f"Value type not supported: {value}")
    raise NotImplementedError(f"Index type not supported: {index}")


# == copy (core.py:649-653) ==


def copy(array, deep=False) -> NDArray:
    fn = lib.array_deep_copy if deep else lib.array_shallow_copy
    return _new(fn(array.buffer))


# == reductions (core.py:661-694) ==


def _dims_arg(array, dims):
    dims = tuple(range(array.ndim)) if dims is None else tuple(dims)
    assert all(isinstance(i, int) for i in dims)
    assert all(0 <= i < array.ndim for i in dims)
    return size_array(dims), len(dims)


def reduce(fn, array, dims=None, init=0.0) -> NDArray:
    """Keepdims reduction with an accumulator that must resolve to sum / max / min / prod."""
    assert isinstance(array, NDArray)
    assert isinstance(dims, tp.Sequence) or dims is None
    dims_c, n = _dims_arg(array, dims)
    return _new(lib.array_reduce_named(_uf.resolve_reduce(fn), array.buffer, dims_c, n, float(init)))


def generate_reduce_op(op: str):
    libfn = getattr(lib, f"array_reduce_{op}")

    def wrapper(array, dims=None):
        assert isinstance(array, NDArray)
        dims_c, n = _dims_arg(array, dims)
        return _new(libfn(array.buffer, dims_c, n))

    wrapper.__doc__ = f"Reduce an array along an axis with {op}"
    wrapper.__name__ = f"array_reduce_{op}"
    return wrapper


reduce_sum = generate_reduce_op("sum")
reduce_max = generate_reduce_op("max")
reduce_min = generate_reduce_op("min")


# == where / cat (core.py:702-748) ==


def where(cond, lhs, rhs) -> NDArray:
    assert isinstance(cond, NDArray)
    l_arr, r_arr = isinstance(lhs, NDArray), isinstance(rhs, NDArray)
    if l_arr and r_arr:
        assert cond.shape == lhs.shape == rhs.shape
        return _new(lib.array_array_array_where(cond.buffer, lhs.buffer, rhs.buffer))
    if l_arr and isinstance(rhs, Number):
        assert cond.shape == lhs.shape
        return _new(lib.array_array_scalar_where(cond.buffer, lhs.buffer, float(rhs)))
    if isinstance(lhs, Number) and r_arr:
        assert cond.shape == rhs.shape
        return _new(lib.array_scalar_array_where(cond.buffer, float(lhs), rhs.buffer))
    if isinstance(lhs, Number) and isinstance(rhs, Number):
        return _new(lib.array_scalar_scalar_where(cond.buffer, float(lhs), float(rhs)))
    raise NotImplementedError(f"not implemented for {type(cond)=} and {type(lhs)=}")


def cat(arrays, dims=None) -> NDArray:
    """Concatenate along `dims`; several dims give a block-diagonal layout padded with zeros."""
    assert isinstance(arrays, tp.Sequence)
    assert all(isinstance(a, NDArray) for a in arrays)
    ndim = arrays[0].ndim
    dims = list(dims) if dims else list(range(ndim))
    assert all(a.ndim == ndim for a in arrays)
    for d in range(ndim):
        if d not in dims:
            assert all(a.shape[d] == arrays[0].shape[d] for a in arrays)
    out_shape = [sum(a.shape[d] for a in arrays) if d in dims else arrays[0].shape[d] for d in range(ndim)]
    out = zeros(out_shape)
    start = [0] * ndim
    for a in arrays:
        end = [start[d] + a.shape[d] if d in dims else out_shape[d] for d in range(ndim)]
        set_view_from_array(out, a, (tuple(start), tuple(end), (1,) * ndim))
        for d in dims:
            start[d] += a.shape[d]
    return out


# == printing (core.py:756-780): formats a host copy ==


def pprepr(array) -> str:
    assert isinstance(array, NDArray)
    return f"NDArray(f32[{','.join(map(str, array.shape))}])"


def ppstr(array) -> str:
    assert isinstance(array, NDArray)
    host = to_numpy(array)
    out = StringIO()

    def emit(node):
        if node.ndim == 0:
            out.write(f"{float(node)}")
            return
        out.write("[")
        for i in range(node.shape[0]):
            if i:
                out.write(", ")
            emit(node[i])
        out.write("]")

    emit(host)
    return out.getvalue()


# == NDArray methods / operators that forward to the functions above (core.py:76-185 in the reference) ==

_METHOD_TABLE = {
    # name: (function, how the method's arguments map onto the function's)
    "reshape": lambda self, shape: reshape(self, shape),
    "transpose": lambda self, dst: transpose(self, dst),
    "move_dim": lambda self, src, dst: move_dim(self, src, dst),
    "ravel": lambda self: ravel(self),
    "as_strided": lambda self, *, shape, stride: as_strided(self, shape, stride),
    "reduce_sum": lambda self, *, dims=None: reduce_sum(self, dims=dims),
    "reduce_max": lambda self, *, dims=None: reduce_max(self, dims=dims),
    "reduce_min": lambda self, *, dims=None: reduce_min(self, dims=dims),
    "where": lambda self, lhs, rhs: where(self, lhs, rhs),
    "cat": lambda self, arrays, *, dims: cat([self] + list(arrays), dims),
    "apply": lambda self, fn: apply(fn, self),
}
_OPERATOR_TABLE = {  # dunder stem -> (function, reflected variant exists)
    "add": (add, True), "sub": (sub, True), "mul": (mul, True), "truediv": (div, True), "pow": (pow, False),
    "matmul": (matmul, False), "lt": (lt, False), "le": (le, False), "gt": (gt, False), "ge": (ge, False),
    "eq": (eq, False), "ne": (ne, False),
}


def _attach_forwarders():
    for name, impl in _METHOD_TABLE.items():
        impl.__name__ = impl.__qualname__ = name
        setattr(NDArray, name, impl)
    for stem, (fn, reflected) in _OPERATOR_TABLE.items():
        setattr(NDArray, f"__{stem}__", (lambda f: lambda self, other: f(self, other))(fn))
        if reflected:
            setattr(NDArray, f"__r{stem}__", (lambda f: lambda self, other: f(other, self))(fn))
    NDArray.__neg__ = lambda self: neg(self)
    NDArray.__hash__ = None  # __eq__ is elementwise, like the reference


_attach_forwarders()
