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

import contextlib
import ctypes as C
import threading

from . import core
from ._lib import BINOPS, UFUNCS, ArrayP, check_ptr, lib

MAX_ARRAYS, MAX_STEPS = 4, 8
ACC_IN, IN_ACC, ACC_SCALAR, SCALAR_ACC, UNARY = range(5)


class ChainStep(C.Structure):
    _fields_ = [("kind", C.c_int), ("op", C.c_int), ("scalar", C.c_float)]


class TapeInsn(C.Structure):
    _fields_ = [("kind", C.c_int), ("op", C.c_int), ("a_src", C.c_int), ("b_src", C.c_int),
                ("a_const", C.c_double), ("b_const", C.c_double)]


lib.array_chain_eval.restype = ArrayP
lib.array_chain_eval.argtypes = [C.POINTER(ArrayP), C.c_size_t, C.POINTER(ChainStep), C.c_size_t]
lib.array_tape_eval_f64.restype = ArrayP
lib.array_tape_eval_f64.argtypes = [C.POINTER(ArrayP), C.c_size_t, C.POINTER(TapeInsn), C.c_size_t, C.c_size_t]
TAPE_CONST, TAPE_MAX = -100, 32

_NP_BINARY = {"add": "sum", "subtract": "sub", "multiply": "mul", "divide": "div", "true_divide": "div", "power": "pow",
              "maximum": "max", "minimum": "min", "fmax": "max", "fmin": "min", "less": "lt", "less_equal": "le",
              "greater": "gt", "greater_equal": "ge", "fmod": "mod"}
_NP_UNARY = {"negative": "neg", "absolute": "abs", "fabs": "abs", "sqrt": "sqrt", "exp": "exp", "log": "log", "sin": "sin",
             "cos": "cos", "tan": "tan", "arcsin": "asin", "arccos": "acos", "arctan": "atan", "sinh": "sinh", "cosh": "cosh",
             "tanh": "tanh", "arcsinh": "asinh", "arccosh": "acosh", "arctanh": "atanh", "ceil": "ceil", "floor": "floor"}


class Expr:
    """A recorded elementwise chain over device arrays."""

    __slots__ = ("arrays", "steps")
    __array_priority__ = 1000

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
            if not other.steps:          # a bare input (e.g. the `x` in `x * x`): just another operand
                other = other.arrays[0]
            else:
                other = other.eval()
        if isinstance(other, core.NDArray):
            base = self._flushed(True)
            return Expr(base.arrays + [other], base.steps + [(IN_ACC if reverse else ACC_IN, code, 0.0)])
        if isinstance(other, core.Number) or type(other).__module__ == "numpy" and getattr(other, "ndim", 1) == 0:
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
    def __eq__(self, o): return self._binary("eq", o, False)
    def __ne__(self, o): return self._binary("ne", o, False)
    def __mod__(self, o): return self._binary("mod", o, False)      # C fmod, like the reference's array % array
    def __rmod__(self, o): return self._binary("mod", o, True)
    def __rpow__(self, o): return self._binary("pow", o, True)
    __hash__ = None  # == builds an expression, so instances cannot be dictionary keys
    def __neg__(self): return self.apply("neg")
    def __pos__(self): return self
    def __abs__(self): return self.apply("abs")
    def __ceil__(self): return self.apply("ceil")
    def __floor__(self): return self.apply("floor")

    def __bool__(self):  # `if x > 0:` or builtin max/min inside a traced lambda: data dependent, not traceable
        raise TypeError("a lazy arraylib expression has no truth value")

    def __float__(self):  # math.sqrt(expr) etc.: math.* needs a real float, so tracing stops here
        raise TypeError("a lazy arraylib expression cannot be converted to a Python float")

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        """numpy ufuncs applied to a lazy expression are recorded as chain steps."""
        if method != "__call__" or kwargs:
            return NotImplemented
        name = ufunc.__name__
        if len(inputs) == 1 and name in _NP_UNARY:
            return self.apply(_NP_UNARY[name])
        if len(inputs) == 2 and name in _NP_BINARY:
            lhs, rhs = inputs
            if lhs is self:
                return self._binary(_NP_BINARY[name], rhs, False)
            return self._binary(_NP_BINARY[name], lhs, True)
        return NotImplemented

    def maximum(self, o): return self._binary("max", o, False)
    def minimum(self, o): return self._binary("min", o, False)
    def sqrt(self): return self.apply("sqrt")
    def abs(self): return self.apply("abs")

    # -- evaluation --------------------------------------------------------------------------------
    def eval(self) -> "core.NDArray":
        if not self.steps:
            return core.copy(self.arrays[0], deep=True)
        if len(self.steps) > MAX_STEPS or len(self.arrays) > MAX_ARRAYS:
            raise NotImplementedError("chain too long")
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


class _Trace:
    """Shared state of one traced lambda: the input arrays and the SSA instruction list."""

    def __init__(self, arrays):
        self.arrays, self.insns = list(arrays), []

    def emit(self, kind, op, a, b=None):
        if len(self.insns) >= TAPE_MAX:
            raise NotImplementedError("traced expression is longer than the device tape")
        self.insns.append((kind, op, a, b))
        return len(self.insns) - 1


class TExpr:
    """Scalar proxy handed to an `apply()` callable. Each arithmetic operation appends one
    instruction to the shared FP64 tape; common sub-expressions are shared, so any expression DAG
    (not just left-deep chains) traces. `src` is an operand reference of the tape ABI."""

    __slots__ = ("trace", "src")
    __array_priority__ = 1000

    def __init__(self, trace, src):
        self.trace, self.src = trace, src

    def _operand(self, other):
        if isinstance(other, TExpr):
            if other.trace is not self.trace:
                raise NotImplementedError("operands from different traces")
            return other.src, 0.0
        if isinstance(other, bool) or isinstance(other, core.Number) or (type(other).__module__ == "numpy" and getattr(other, "ndim", 1) == 0):
            return TAPE_CONST, float(other)
        return None

    def _binary(self, op, other, reverse):
        operand = self._operand(other)
        if operand is None:
            return NotImplemented
        mine = (self.src, 0.0)
        a, b = (operand, mine) if reverse else (mine, operand)
        return TExpr(self.trace, self.trace.emit(0, BINOPS.index(op), a, b))

    def _unary(self, op):
        return TExpr(self.trace, self.trace.emit(1, UFUNCS.index(op), (self.src, 0.0)))

    def __add__(self, o): return self._binary("sum", o, False)
    def __radd__(self, o): return self._binary("sum", o, True)
    def __sub__(self, o): return self._binary("sub", o, False)
    def __rsub__(self, o): return self._binary("sub", o, True)
    def __mul__(self, o): return self._binary("mul", o, False)
    def __rmul__(self, o): return self._binary("mul", o, True)
    def __truediv__(self, o): return self._binary("div", o, False)
    def __rtruediv__(self, o): return self._binary("div", o, True)
    def __pow__(self, o): return self._binary("pow", o, False)
    def __rpow__(self, o): return self._binary("pow", o, True)
    def __lt__(self, o): return self._binary("lt", o, False)
    def __le__(self, o): return self._binary("le", o, False)
    def __gt__(self, o): return self._binary("gt", o, False)
    def __ge__(self, o): return self._binary("ge", o, False)
    def __eq__(self, o): return self._binary("eq", o, False)
    def __ne__(self, o): return self._binary("ne", o, False)
    def __mod__(self, o): return self._binary("mod", o, False)      # C fmod (sign of the dividend), NOT Python's %
    def __rmod__(self, o): return self._binary("mod", o, True)
    __hash__ = None
    def __neg__(self): return self._unary("neg")
    def __pos__(self): return self
    def __abs__(self): return self._unary("abs")
    def __ceil__(self): return self._unary("ceil")
    def __floor__(self): return self._unary("floor")

    def __bool__(self):  # `if x > 0:` or builtin max/min: data dependent, cannot be traced
        raise TypeError("a traced arraylib scalar has no truth value")

    def __float__(self):  # math.sqrt(x) etc. need a real float: tracing stops here
        raise TypeError("a traced arraylib scalar cannot be converted to a Python float")

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__" or kwargs:
            return NotImplemented
        name = ufunc.__name__
        if len(inputs) == 1 and name in _NP_UNARY:
            return self._unary(_NP_UNARY[name])
        if len(inputs) == 2 and name in _NP_BINARY:
            lhs, rhs = inputs
            if isinstance(lhs, TExpr):
                return lhs._binary(_NP_BINARY[name], rhs, False)
            return self._binary(_NP_BINARY[name], lhs, True)
        return NotImplemented


_MATH_UNARY = {"sqrt": "sqrt", "exp": "exp", "log": "log", "sin": "sin", "cos": "cos", "tan": "tan", "asin": "asin",
               "acos": "acos", "atan": "atan", "sinh": "sinh", "cosh": "cosh", "tanh": "tanh", "asinh": "asinh",
               "acosh": "acosh", "atanh": "atanh", "fabs": "abs", "ceil": "ceil", "floor": "floor"}
_MATH_BINARY = {"pow": "pow", "fmod": "mod"}
_math_patch_lock = threading.RLock()


@contextlib.contextmanager
def _math_traceable():
    """While a callable is being traced, `math.sqrt(x)` & co. must record an instruction instead of asking the
    proxy for a float. The `math` module's functions are swapped for wrappers that do that for proxies and call
    the original for everything else (so other threads keep working), and restored afterwards."""
    import math
    with _math_patch_lock:
        saved = {}

        def unary(name, op):
            orig = saved[name]

            def wrapper(x, *rest):
                if isinstance(x, TExpr) and not rest:
                    return x._unary(op)
                return orig(x, *rest)
            return wrapper

        def binary(name, op):
            orig = saved[name]

            def wrapper(x, y):
                if isinstance(x, TExpr):
                    return x._binary(op, y, False)
                if isinstance(y, TExpr):
                    return y._binary(op, x, True)
                return orig(x, y)
            return wrapper

        for name in list(_MATH_UNARY) + list(_MATH_BINARY):
            saved[name] = getattr(math, name)
        try:
            for name, op in _MATH_UNARY.items():
                setattr(math, name, unary(name, op))
            for name, op in _MATH_BINARY.items():
                setattr(math, name, binary(name, op))
            yield
        finally:
            for name, orig in saved.items():
                setattr(math, name, orig)


def _run_traced(fn, n_inputs):
    """(trace, result proxy) of `fn` called on `n_inputs` scalar proxies, or None when it does not trace
    (data-dependent branch, builtin max/min, float(), unsupported operation, wrong return type)."""
    trace = _Trace([None] * n_inputs)
    try:
        with _math_traceable():
            out = fn(*[TExpr(trace, -1 - i) for i in range(n_inputs)])
    except (TypeError, NotImplementedError, AttributeError, ValueError):
        return None
    if not isinstance(out, TExpr) or out.trace is not trace:
        return None
    return trace, out


def single_instruction(fn, n_inputs):
    """If `fn` traces to exactly ONE instruction over its distinct inputs (`lambda x: math.sqrt(x)`,
    `lambda a, b: a + b`), return (kind, op name): kind 1 = unary ufunc, 0 = binary op. Else None."""
    traced = _run_traced(fn, n_inputs)
    if traced is None:
        return None
    trace, out = traced
    if len(trace.insns) != 1 or out.src != 0:
        return None
    kind, op, a, b = trace.insns[0]
    if kind == 1 and n_inputs == 1 and a[0] == -1:
        return 1, UFUNCS[op]
    if kind == 0 and n_inputs == 2 and {a[0], b[0]} == {-1, -2}:
        return 0, BINOPS[op]
    return None


def trace_unary(fn, array):
    """Try to record `fn` (a Python callable of one scalar) as an FP64 tape over `array`: arithmetic
    operators, comparisons, abs/neg/ceil/floor, numpy ufuncs and `math.*` functions trace; branches on the
    value, builtin max/min and anything else do not. Returns the evaluated NDArray, or None when `fn` did not trace."""
    traced = _run_traced(fn, 1)
    if traced is None:
        return None
    trace, out = traced
    if out.src < 0:  # the lambda returned its argument unchanged
        return core.copy(array, deep=True)
    insns = (TapeInsn * len(trace.insns))()
    for i, (kind, op, a, b) in enumerate(trace.insns):
        b = b if b is not None else (TAPE_CONST, 0.0)
        insns[i] = TapeInsn(kind, op, a[0], b[0], a[1], b[1])
    ptrs = (ArrayP * 1)(array.buffer)
    return core.NDArray(check_ptr(lib.array_tape_eval_f64(ptrs, 1, insns, len(trace.insns), out.src)))


def fused(fn, *arrays):
    """Run `fn` over lazy wrappers of `arrays` and evaluate the resulting chain in one kernel."""
    out = fn(*[lazy(a) for a in arrays])
    return out.eval() if isinstance(out, Expr) else out
