"""Opt-in fusion of elementwise chains.

The reference (and the default `al.*` operators here) evaluate `a + b + c + 1.0` eagerly: three
passes over memory and two temporaries. Wrapping the first operand in `lazy()` records the chain
instead and `eval()` runs it as ONE kernel (`array_chain_eval`), with bit-identical results:

    u = (al.lazy(A) + B + c + 1.0).eval()          # or: al.fused(lambda a, b, c: a + b + c + 1.0, A, B, c)

A chain is left-deep: each step combines the running value with one array or scalar, or applies a
named ufunc. When the other operand is itself a lazy chain it is evaluated first; chains longer
than the kernel's limits (4 arrays, 8 steps) are flushed to a temporary and continued.
"""
from __future__ import annotations

import ctypes as C

from . import core
from ._lib import BINOPS, UFUNCS, ArrayP, check_ptr, lib

MAX_ARRAYS, MAX_STEPS = 4, 8
ACC_IN, IN_ACC, ACC_SCALAR, SCALAR_ACC, UNARY = range(5)


class ChainStep(C.Structure):
    _fields_ = [("kind", C.c_int), ("op", C.c_int), ("scalar", C.c_float)]


lib.array_chain_eval.restype = ArrayP
lib.array_chain_eval.argtypes = [C.POINTER(ArrayP), C.c_size_t, C.POINTER(ChainStep), C.c_size_t]


class Expr:
    """A recorded elementwise chain over device arrays."""

    __slots__ = ("arrays", "steps")

    def __init__(self, arrays, steps=()):
        self.arrays, self.steps = list(arrays), list(steps)

    # -- recording ---------------------------------------------------------------------------------
    def _flushed(self, needs_array: bool):
        """Self if it still has room for one more step (and one more array when the step consumes
        one), else a fresh chain over its evaluated value."""
        if len(self.steps) >= MAX_STEPS or (needs_array and len(self.arrays) >= MAX_ARRAYS):
            return Expr([self.eval()])
        return self

    def _binary(self, op: str, other, reverse: bool):
        code = BINOPS.index(op)
        if isinstance(other, Expr):
            other = other.eval()
        if isinstance(other, core.NDArray):
            base = self._flushed(True)
            return Expr(base.arrays + [other], base.steps + [(IN_ACC if reverse else ACC_IN, code, 0.0)])
        if isinstance(other, core.Number):
            base = self._flushed(False)
            return Expr(base.arrays, base.steps + [(SCALAR_ACC if reverse else ACC_SCALAR, code, float(other))])
        return NotImplemented

    def apply(self, fn):
        from . import ufuncs
        base = self._flushed(False)
        return Expr(base.arrays, base.steps + [(UNARY, ufuncs.resolve_unary(fn), 0.0)])

    def __add__(self, o): return self._binary("sum", o, False)
    def __radd__(self, o): return self._binary("sum", o, True)
    def __sub__(self, o): return self._binary("sub", o, False)
    def __rsub__(self, o): return self._binary("sub", o, True)
    def __mul__(self, o): return self._binary("mul", o, False)
    def __rmul__(self, o): return self._binary("mul", o, True)
    def __truediv__(self, o): return self._binary("div", o, False)
    def __rtruediv__(self, o): return self._binary("div", o, True)
    def __pow__(self, o): return self._binary("pow", o, False)
    def __lt__(self, o): return self._binary("lt", o, False)
    def __le__(self, o): return self._binary("le", o, False)
    def __gt__(self, o): return self._binary("gt", o, False)
    def __ge__(self, o): return self._binary("ge", o, False)
    def __neg__(self): return self.apply("neg")
    def maximum(self, o): return self._binary("max", o, False)
    def minimum(self, o): return self._binary("min", o, False)
    def sqrt(self): return self.apply("sqrt")
    def abs(self): return self.apply("abs")

    # -- evaluation --------------------------------------------------------------------------------
    def eval(self) -> "core.NDArray":
        if not self.steps:
            return core.copy(self.arrays[0], deep=True)
        n = len(self.arrays)
        ptrs = (ArrayP * n)(*[a.buffer for a in self.arrays])
        steps = (ChainStep * len(self.steps))(*[ChainStep(k, o, s) for k, o, s in self.steps])
        return core.NDArray(check_ptr(lib.array_chain_eval(ptrs, n, steps, len(self.steps))))

    def __repr__(self):
        names = [f"{('acc∘in', 'in∘acc', 'acc∘s', 's∘acc', 'ufunc')[k]}:{(UFUNCS if k == UNARY else BINOPS)[o]}" for k, o, _ in self.steps]
        return f"Expr({len(self.arrays)} arrays; {' -> '.join(names) or 'identity'})"


def lazy(array) -> Expr:
    assert isinstance(array, core.NDArray)
    return Expr([array])


def fused(fn, *arrays):
    """Run `fn` over lazy wrappers of `arrays` and evaluate the resulting chain in one kernel."""
    out = fn(*[lazy(a) for a in arrays])
    return out.eval() if isinstance(out, Expr) else out
