"""ctypes binding of the UNMODIFIED reference library compiled into oracle/_ref/ -- TEST
INFRASTRUCTURE ONLY (the strongest form of the oracle: the reference's own machine code).

oracle/_ref/arraylib_ref.so      serial build   (deterministic summation order)
oracle/_ref/arraylib_ref_omp.so  -fopenmp build (the CPU baseline that bench.py times)

Both come from `make -C oracle ref`, which compiles /root/reference/arraylib/src/arraylib.c in
place; no reference source is copied into this repo. The struct mirrors below describe the
reference ABI (arraylib/src/arraylib.h:57-77) so that ctypes can walk the returned pointers.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")


class _Data(C.Structure):
    _fields_ = [("mem", C.POINTER(C.c_float)), ("size", C.c_size_t), ("refs", C.c_size_t)]


class _Layout(C.Structure):
    _fields_ = [("ndim", C.c_size_t), ("shape", C.POINTER(C.c_size_t)), ("stride", C.POINTER(C.c_size_t))]


class _NDArray(C.Structure):
    _fields_ = [("data", C.POINTER(_Data)), ("ptr", C.POINTER(C.c_float)),
                ("lay", C.POINTER(_Layout)), ("view", C.c_bool)]


_P = C.POINTER(_NDArray)
_SZP = C.POINTER(C.c_size_t)


def available(omp: bool = False) -> bool:
    return os.path.exists(os.path.join(REF_DIR, "arraylib_ref_omp.so" if omp else "arraylib_ref.so"))


def _sz(seq):
    seq = [int(x) for x in seq]
    return (C.c_size_t * max(len(seq), 1))(*seq)


class RefArray:
    """Owns one reference NDArray*."""

    def __init__(self, lib: "RefLib", p):
        assert p, "reference returned NULL"
        self.lib, self.p = lib, p

    def __del__(self):
        try:
            self.lib.c.array_free(self.p)
        except Exception:
            pass

    @property
    def ndim(self):
        return int(self.p.contents.lay.contents.ndim)

    @property
    def shape(self):
        lay = self.p.contents.lay.contents
        return tuple(int(lay.shape[i]) for i in range(lay.ndim))

    @property
    def stride(self):
        lay = self.p.contents.lay.contents
        return tuple(int(lay.stride[i]) for i in range(lay.ndim))

    @property
    def view(self):
        return bool(self.p.contents.view)

    @property
    def base_size(self):
        return int(self.p.contents.data.contents.size)

    def numpy(self) -> np.ndarray:
        """Copy of the LOGICAL view (through ptr + strides), i.e. what the C loops read."""
        addr = C.cast(self.p.contents.ptr, C.c_void_p).value
        base = C.cast(self.p.contents.data.contents.mem, C.c_void_p).value
        off = (addr - base) // 4
        flat = np.ctypeslib.as_array(self.p.contents.data.contents.mem, shape=(self.base_size,))
        v = np.lib.stride_tricks.as_strided(flat[off:], shape=self.shape,
                                            strides=[4 * s for s in self.stride])
        return np.array(v, dtype=np.float32, copy=True)


class RefLib:
    def __init__(self, omp: bool = False):
        path = os.path.join(REF_DIR, "arraylib_ref_omp.so" if omp else "arraylib_ref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
        self.c = c = C.CDLL(path)
        self.omp = omp
        c.array_free.argtypes = [_P]
        c.array_free.restype = None
        for name in ("array_fill",):
            getattr(c, name).argtypes = [C.c_void_p, _SZP, C.c_size_t]
            getattr(c, name).restype = _P
        for name in ("array_ones", "array_zeros"):
            getattr(c, name).argtypes = [_SZP, C.c_size_t]
            getattr(c, name).restype = _P
        for name in ("array_arange", "array_linspace"):
            getattr(c, name).argtypes = [C.c_float] * 3
            getattr(c, name).restype = _P
        from .oracle import BINARY, UNARY  # op-name tables only
        for op in BINARY:
            f = getattr(c, f"array_array_{op}")
            f.argtypes, f.restype = [_P, _P], _P
            f = getattr(c, f"array_scalar_{op}")
            f.argtypes, f.restype = [_P, C.c_float], _P
        for op in UNARY:
            f = getattr(c, f"array_{op}")
            f.argtypes, f.restype = [_P], _P
        for op in ("sum", "max", "min"):
            f = getattr(c, f"array_reduce_{op}")
            f.argtypes, f.restype = [_P, _SZP, C.c_size_t], _P
        for name in ("array_deep_copy", "array_shallow_copy", "array_ravel"):
            getattr(c, name).argtypes, getattr(c, name).restype = [_P], _P
        c.array_reshape.argtypes, c.array_reshape.restype = [_P, _SZP, C.c_size_t], _P
        c.array_transpose.argtypes, c.array_transpose.restype = [_P, _SZP], _P
        c.array_move_dim.argtypes, c.array_move_dim.restype = [_P, _SZP, _SZP, C.c_size_t], _P
        c.array_as_strided.argtypes, c.array_as_strided.restype = [_P, C.POINTER(_Layout)], _P
        c.array_get_view.argtypes, c.array_get_view.restype = [_P, _SZP, _SZP, _SZP], _P
        c.array_array_matmul.argtypes, c.array_array_matmul.restype = [_P, _P], _P
        c.array_array_dot.argtypes, c.array_array_dot.restype = [_P, _P], _P
        c.array_array_array_where.argtypes, c.array_array_array_where.restype = [_P, _P, _P], _P
        c.array_array_scalar_where.argtypes, c.array_array_scalar_where.restype = [_P, _P, C.c_float], _P
        c.array_scalar_array_where.argtypes, c.array_scalar_array_where.restype = [_P, C.c_float, _P], _P
        c.array_set_view_from_array.argtypes = [_P, _P, _SZP, _SZP, _SZP]
        c.array_set_view_from_array.restype = _P

    # ---- creation ---------------------------------------------------------------------------
    def from_numpy(self, a: np.ndarray) -> RefArray:
        a = np.ascontiguousarray(a, dtype=np.float32)
        shape = a.shape if a.ndim else (1,)
        return RefArray(self, self.c.array_fill(C.c_void_p(a.ctypes.data), _sz(shape), len(shape)))

    def ones(self, shape):
        return RefArray(self, self.c.array_ones(_sz(shape), len(shape)))

    def zeros(self, shape):
        return RefArray(self, self.c.array_zeros(_sz(shape), len(shape)))

    def arange(self, start, end, step=1):
        return RefArray(self, self.c.array_arange(start, end, step))

    def linspace(self, start, end, n):
        return RefArray(self, self.c.array_linspace(start, end, n))

    # ---- hot path ---------------------------------------------------------------------------
    def binary(self, op: str, a: RefArray, b) -> RefArray:
        if isinstance(b, RefArray):
            return RefArray(self, getattr(self.c, f"array_array_{op}")(a.p, b.p))
        return RefArray(self, getattr(self.c, f"array_scalar_{op}")(a.p, float(b)))

    def unary(self, op: str, a: RefArray) -> RefArray:
        return RefArray(self, getattr(self.c, f"array_{op}")(a.p))

    def reduce(self, op: str, a: RefArray, dims=None) -> RefArray:
        dims = tuple(range(a.ndim)) if dims is None else tuple(dims)
        return RefArray(self, getattr(self.c, f"array_reduce_{op}")(a.p, _sz(dims), len(dims)))

    def deep_copy(self, a):
        return RefArray(self, self.c.array_deep_copy(a.p))

    def reshape(self, a, shape):
        return RefArray(self, self.c.array_reshape(a.p, _sz(shape), len(shape)))

    def transpose(self, a, dims):
        return RefArray(self, self.c.array_transpose(a.p, _sz(dims)))

    def move_dim(self, a, src, dst):
        return RefArray(self, self.c.array_move_dim(a.p, _sz(src), _sz(dst), len(src)))

    def ravel(self, a):
        return RefArray(self, self.c.array_ravel(a.p))

    def as_strided(self, a, shape, stride):
        sh, st = _sz(shape), _sz(stride)
        lay = _Layout(len(shape), C.cast(sh, _SZP), C.cast(st, _SZP))
        return RefArray(self, self.c.array_as_strided(a.p, C.byref(lay)))

    def get_view(self, a, start, end, step):
        return RefArray(self, self.c.array_get_view(a.p, _sz(start), _sz(end), _sz(step)))

    def set_view_from_array(self, dst, src, start, end, step):
        self.c.array_set_view_from_array(dst.p, src.p, _sz(start), _sz(end), _sz(step))
        return dst

    def matmul(self, a, b):
        return RefArray(self, self.c.array_array_matmul(a.p, b.p))

    def dot(self, a, b):
        return RefArray(self, self.c.array_array_dot(a.p, b.p))

    def where(self, cond, lhs, rhs):
        if isinstance(lhs, RefArray) and isinstance(rhs, RefArray):
            return RefArray(self, self.c.array_array_array_where(cond.p, lhs.p, rhs.p))
        if isinstance(lhs, RefArray):
            return RefArray(self, self.c.array_array_scalar_where(cond.p, lhs.p, float(rhs)))
        return RefArray(self, self.c.array_scalar_array_where(cond.p, float(lhs), rhs.p))
