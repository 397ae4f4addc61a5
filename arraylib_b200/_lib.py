"""ctypes loader for libarraylib_b200.so (declared in include/arraylib_b200.h).

This replaces the reference's `ffi.cdef(...)` + `ffi.dlopen(".../arraylib.so")`
(/root/reference/arraylib/core.py:17-33). There is NO fallback: if the CUDA library has not
been built the import fails, and if no GPU is visible the first device call raises.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libarraylib_b200.so")

BINOPS = ("sum", "sub", "mul", "div", "mod", "pow", "max", "min", "eq", "ne", "gt", "ge", "lt", "le")
UFUNCS = ("neg", "abs", "sqrt", "exp", "log", "sin", "cos", "tan", "asin", "acos", "atan", "sinh",
          "cosh", "tanh", "asinh", "acosh", "atanh", "ceil", "floor")
REDOPS = ("sum", "max", "min", "prod")


class Data(C.Structure):  # arraylib.h:57-61; `mem` is a DEVICE pointer here
    _fields_ = [("mem", C.c_void_p), ("size", C.c_size_t), ("refs", C.c_size_t)]


class CLayout(C.Structure):  # arraylib.h:63-67
    _fields_ = [("ndim", C.c_size_t), ("shape", C.POINTER(C.c_size_t)), ("stride", C.POINTER(C.c_size_t))]


class CArray(C.Structure):  # arraylib.h:72-77; `ptr` is a DEVICE pointer here
    _fields_ = [("data", C.POINTER(Data)), ("ptr", C.c_void_p), ("lay", C.POINTER(CLayout)), ("view", C.c_bool)]


ArrayP = C.POINTER(CArray)
SizeP = C.POINTER(C.c_size_t)
F32 = C.c_float

_EXC = {
    "ValueError": ValueError, "TypeError": TypeError, "IndexError": IndexError, "KeyError": KeyError,
    "MemoryError": MemoryError, "NotImplementedError": NotImplementedError, "RuntimeError": RuntimeError,
    "AssertionError": AssertionError,
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or `make -C arraylib_b200/csrc`). "
            "arraylib_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)

    def sig(name, res, *args):
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, list(args)

    sig("al_init", C.c_int, C.c_int)
    sig("al_device", C.c_int)
    sig("al_device_count", C.c_int)
    sig("al_sync", C.c_int)
    sig("al_stream", C.c_void_p)
    sig("al_last_error", C.c_char_p)
    sig("al_version", C.c_char_p)
    sig("al_launch_count", C.c_uint64)
    sig("al_last_kernel", C.c_char_p)
    sig("al_timer_start", C.c_int)
    sig("al_timer_stop", C.c_int, C.POINTER(C.c_float))
    sig("al_event_record", C.c_int, C.c_int)
    sig("al_event_elapsed", C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float))
    sig("al_bytes_in_use", C.c_uint64)
    sig("al_trim_pool", C.c_int)
    sig("al_bytes_cached", C.c_uint64)
    sig("al_host_alloc", C.c_void_p, C.c_size_t)
    sig("al_host_free", C.c_int, C.c_void_p)
    sig("al_set_tuning", C.c_int, C.c_char_p, C.c_long)

    sig("array_empty", ArrayP, SizeP, C.c_size_t)
    sig("array_free", None, ArrayP)
    sig("array_shallow_copy", ArrayP, ArrayP)
    sig("array_deep_copy", ArrayP, ArrayP)
    sig("array_zeros", ArrayP, SizeP, C.c_size_t)
    sig("array_ones", ArrayP, SizeP, C.c_size_t)
    sig("array_fill", ArrayP, C.c_void_p, SizeP, C.c_size_t)
    sig("array_full", ArrayP, SizeP, C.c_size_t, F32)
    sig("array_arange", ArrayP, F32, F32, F32)
    sig("array_linspace", ArrayP, F32, F32, F32)
    sig("array_to_host", C.c_int, ArrayP, C.c_void_p)
    sig("array_fill_async", ArrayP, C.c_void_p, SizeP, C.c_size_t)
    sig("array_to_host_async", C.c_int, ArrayP, C.c_void_p)
    sig("array_read_base", C.c_int, ArrayP, C.c_size_t, C.c_size_t, C.c_void_p)
    sig("array_write_base", C.c_int, ArrayP, C.c_size_t, C.c_size_t, C.c_void_p)
    sig("array_get_elem", F32, ArrayP, SizeP)
    sig("array_get_view", ArrayP, ArrayP, SizeP, SizeP, SizeP)
    sig("array_set_elem_from_scalar", ArrayP, ArrayP, F32, SizeP)
    sig("array_set_view_from_scalar", ArrayP, ArrayP, F32, SizeP, SizeP, SizeP)
    sig("array_set_view_from_array", ArrayP, ArrayP, ArrayP, SizeP, SizeP, SizeP)
    sig("array_as_strided", ArrayP, ArrayP, C.POINTER(CLayout))
    sig("array_reshape", ArrayP, ArrayP, SizeP, C.c_size_t)
    sig("array_transpose", ArrayP, ArrayP, SizeP)
    sig("array_move_dim", ArrayP, ArrayP, SizeP, SizeP, C.c_size_t)
    sig("array_ravel", ArrayP, ArrayP)
    sig("array_scalar_op_named", ArrayP, C.c_int, ArrayP, F32)
    sig("array_array_op_named", ArrayP, C.c_int, ArrayP, ArrayP)
    for op in BINOPS:
        sig(f"array_array_{op}", ArrayP, ArrayP, ArrayP)
        sig(f"array_scalar_{op}", ArrayP, ArrayP, F32)
    sig("array_array_matmul", ArrayP, ArrayP, ArrayP)
    sig("array_array_dot", ArrayP, ArrayP, ArrayP)
    sig("array_op", ArrayP, C.c_void_p, ArrayP)
    sig("array_op_named", ArrayP, C.c_int, ArrayP)
    for op in UFUNCS:
        sig(f"array_{op}", ArrayP, ArrayP)
    sig("array_reduce", ArrayP, C.c_void_p, ArrayP, SizeP, C.c_size_t, F32)
    sig("array_reduce_named", ArrayP, C.c_int, ArrayP, SizeP, C.c_size_t, F32)
    for op in ("sum", "max", "min"):
        sig(f"array_reduce_{op}", ArrayP, ArrayP, SizeP, C.c_size_t)
    sig("array_array_array_where", ArrayP, ArrayP, ArrayP, ArrayP)
    sig("array_array_scalar_where", ArrayP, ArrayP, ArrayP, F32)
    sig("array_scalar_array_where", ArrayP, ArrayP, F32, ArrayP)
    sig("array_scalar_scalar_where", ArrayP, ArrayP, F32, F32)
    sig("al_comm_unique_id", C.c_int, C.c_char_p)
    sig("al_comm_init", C.c_int, C.c_char_p, C.c_int, C.c_int)
    sig("al_comm_allreduce", C.c_int, ArrayP, C.c_int)
    sig("al_comm_allgather", C.c_int, ArrayP, ArrayP)
    sig("al_comm_alltoall", C.c_int, ArrayP, ArrayP)
    sig("al_comm_reduce", ArrayP, C.c_int, ArrayP, SizeP, C.c_size_t)
    sig("al_comm_reduce_scatter", ArrayP, C.c_int, ArrayP, SizeP, C.c_size_t)
    sig("al_comm_reduce_ex", ArrayP, C.c_int, ArrayP, SizeP, C.c_size_t, C.c_size_t, C.c_int)
    sig("al_comm_can_scatter", C.c_int, SizeP, C.c_size_t, SizeP, C.c_size_t)
    sig("al_comm_transport", C.c_char_p)
    sig("al_comm_destroy", C.c_int)
    return lib


lib = _load()


def raise_last_error(default="RuntimeError: libarraylib_b200 call failed"):
    """Turn the library's thread-local "<ExcName>: message" string into a Python exception."""
    msg = (lib.al_last_error() or b"").decode() or default
    name, _, rest = msg.partition(": ")
    raise _EXC.get(name, RuntimeError)(rest if name in _EXC else msg)


def check_ptr(p):
    if not p:
        raise_last_error()
    return p


def check_rc(rc):
    if rc != 0:
        raise_last_error()


def size_array(seq):
    seq = tuple(int(x) for x in seq)
    return (C.c_size_t * max(len(seq), 1))(*seq)
