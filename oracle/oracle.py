"""numpy-facing wrapper around liboracle.so (oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. Inputs are float32 numpy arrays of ANY (non-negative) strides, so numpy's
own `as_strided` / transposes exercise the same index arithmetic as the reference's views.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

BINARY = ["sum", "sub", "mul", "div", "mod", "pow", "max", "min", "eq", "ne", "gt", "ge", "lt", "le"]
UNARY = ["neg", "abs", "sqrt", "exp", "log", "sin", "cos", "tan", "asin", "acos", "atan", "sinh",
         "cosh", "tanh", "asinh", "acosh", "atanh", "ceil", "floor"]
REDUCE = ["sum", "max", "min", "mul"]
REDUCE_INIT = {"sum": 0.0, "max": -np.inf, "min": np.inf, "mul": 1.0}


def build(force: bool = False) -> str:
    """Compile oracle.c (and, when /root/reference exists, oracle/_ref/) via the Makefile."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_arange.restype = C.c_size_t
        _lib.orc_arange.argtypes = [C.c_float, C.c_float, C.c_float, C.c_void_p]
        _lib.orc_linspace.restype = C.c_size_t
        _lib.orc_linspace.argtypes = [C.c_float, C.c_float, C.c_float, C.c_void_p]
        _lib.orc_broadcast.restype = C.c_size_t
    return _lib


def _sz(seq):
    seq = [int(x) for x in seq]
    return (C.c_size_t * max(len(seq), 1))(*seq)


def _view(a: np.ndarray):
    """(base pointer, element strides) of a float32 array whose strides are multiples of 4."""
    assert a.dtype == np.float32
    assert all(s >= 0 and s % 4 == 0 for s in a.strides), a.strides
    return C.c_void_p(a.ctypes.data), _sz([s // 4 for s in a.strides])


def _ptr(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


def broadcast_layouts(a: np.ndarray, b: np.ndarray):
    na, nb = a.ndim, b.ndim
    n = max(na, nb)
    shape, sa, sb = _sz([0] * n), _sz([0] * n), _sz([0] * n)
    assert lib().orc_broadcastable(_sz(a.shape), C.c_size_t(na), _sz(b.shape), C.c_size_t(nb))
    lib().orc_broadcast(_sz(a.shape), _view(a)[1], C.c_size_t(na), _sz(b.shape), _view(b)[1],
                        C.c_size_t(nb), shape, sa, sb)
    return tuple(shape[:n]), sa, sb


def binary(op: str, a: np.ndarray, b) -> np.ndarray:
    o = BINARY.index(op)
    if not isinstance(b, np.ndarray):
        out = np.empty(a.shape, np.float32)
        pa, sa = _view(a)
        lib().orc_binary_scalar(o, pa, sa, _sz(a.shape), C.c_size_t(a.ndim), C.c_float(float(b)), _ptr(out))
        return out
    shape, sa, sb = broadcast_layouts(a, b)
    out = np.empty(shape, np.float32)
    lib().orc_binary(o, _ptr(a), sa, _ptr(b), sb, _sz(shape), C.c_size_t(len(shape)), _ptr(out))
    return out


def unary(op: str, a: np.ndarray) -> np.ndarray:
    out = np.empty(a.shape, np.float32)
    pa, sa = _view(a)
    lib().orc_unary(UNARY.index(op), pa, sa, _sz(a.shape), C.c_size_t(a.ndim), _ptr(out))
    return out


def gather(a: np.ndarray) -> np.ndarray:
    out = np.empty(a.shape, np.float32)
    pa, sa = _view(a)
    lib().orc_gather(pa, sa, _sz(a.shape), C.c_size_t(a.ndim), _ptr(out))
    return out


def reduce(op: str, a: np.ndarray, dims=None, init=None) -> np.ndarray:
    dims = tuple(range(a.ndim)) if dims is None else tuple(dims)
    mask = (C.c_uint8 * a.ndim)(*[1 if d in dims else 0 for d in range(a.ndim)])
    oshape = tuple(1 if d in dims else a.shape[d] for d in range(a.ndim))
    out = np.empty(oshape, np.float32)
    pa, sa = _view(a)
    init = REDUCE_INIT[op] if init is None else init
    lib().orc_reduce(REDUCE.index(op), pa, sa, _sz(a.shape), C.c_size_t(a.ndim), mask,
                     C.c_float(init), _ptr(out))
    return out


def arange(start, end, step=1) -> np.ndarray:
    n = lib().orc_arange(start, end, step, None)
    out = np.empty(n, np.float32)
    lib().orc_arange(start, end, step, _ptr(out))
    return out


def linspace(start, end, n) -> np.ndarray:
    cnt = lib().orc_arange(0.0, n, 1.0, None)
    out = np.empty(cnt, np.float32)
    lib().orc_linspace(start, end, n, _ptr(out))
    return out


def where(cond: np.ndarray, lhs, rhs) -> np.ndarray:
    out = np.empty(cond.shape, np.float32)
    pc, sc = _view(cond)
    pl, sl = _view(lhs) if isinstance(lhs, np.ndarray) else (None, None)
    pr, sr = _view(rhs) if isinstance(rhs, np.ndarray) else (None, None)
    lib().orc_where(pc, sc, pl, sl, C.c_float(0.0 if pl else float(lhs)), pr, sr,
                    C.c_float(0.0 if pr else float(rhs)), _sz(cond.shape), C.c_size_t(cond.ndim), _ptr(out))
    return out


def matmul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    M, K = a.shape
    K2, N = b.shape
    assert K == K2
    out = np.empty((M, N), np.float32)
    (pa, sa), (pb, sb) = _view(a), _view(b)
    lib().orc_matmul(pa, C.c_size_t(sa[0]), C.c_size_t(sa[1]), pb, C.c_size_t(sb[0]), C.c_size_t(sb[1]),
                     C.c_size_t(M), C.c_size_t(K), C.c_size_t(N), _ptr(out))
    return out
