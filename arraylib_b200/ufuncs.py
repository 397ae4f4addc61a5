"""Resolution of Python callables to entries of the device ufunc / reduce tables.

The reference's `apply(fn, a)` and `reduce(fn, a, ...)` wrap ANY Python callable as a C callback
and invoke it once per element from the CPU loop (/root/reference/arraylib/core.py:455-460,
661-673). A Python callable cannot run inside a CUDA kernel, so here the callable has to resolve
to one of the named device kernels (include/arraylib_b200.h: al_ufunc / al_redop):

  1. by identity or name, for `math.*`, `numpy.*` ufuncs, `operator.*` and builtins;
  2. otherwise by fingerprint: the callable is evaluated ON THE HOST on a fixed handful of probe
     scalars (never on the array data) and must agree exactly with one table entry on all of them.

Anything that does not resolve raises NotImplementedError -- there is no silent CPU fallback.
"""
from __future__ import annotations

import math
import operator
import warnings

from ._lib import REDOPS, UFUNCS

_ALIASES = {
    "negative": "neg", "neg": "neg", "absolute": "abs", "fabs": "abs", "abs": "abs",
    "arcsin": "asin", "arccos": "acos", "arctan": "atan", "arcsinh": "asinh", "arccosh": "acosh",
    "arctanh": "atanh",
}
_TRUSTED_MODULES = {"math", "numpy", "operator", "builtins", "_operator"}

_HOST = {
    "neg": operator.neg, "abs": abs, "sqrt": math.sqrt, "exp": math.exp, "log": math.log,
    "sin": math.sin, "cos": math.cos, "tan": math.tan, "asin": math.asin, "acos": math.acos,
    "atan": math.atan, "sinh": math.sinh, "cosh": math.cosh, "tanh": math.tanh,
    "asinh": math.asinh, "acosh": math.acosh, "atanh": math.atanh, "ceil": math.ceil,
    "floor": math.floor,
}
_UNARY_PROBES = (0.5, 1.0, 2.0, -0.75, 0.25, 3.0, 10.0, 1.5, -2.5, 0.0, 7.25, -0.3)

_HOST_BIN = {"sum": operator.add, "max": max, "min": min, "prod": operator.mul}
_BINARY_PROBES = ((0.5, 2.0), (3.0, -1.5), (-2.25, -4.0), (7.0, 7.0), (0.0, 1.25), (10.0, 0.125))


def _safe(fn, *args):
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return float(fn(*args))
    except (ValueError, OverflowError, ZeroDivisionError, ArithmeticError):
        return math.nan
    except TypeError:
        return math.nan


def _same(a: float, b: float) -> bool:
    # Domain errors surface differently (math raises, numpy returns nan/inf with a warning), so a
    # probe that is non-finite on either side carries no information.
    if not (math.isfinite(a) and math.isfinite(b)):
        return not (math.isfinite(a) or math.isfinite(b))
    return a == b or abs(a - b) <= 4e-16 * max(abs(a), abs(b))


def _by_name(fn, table):
    name = getattr(fn, "__name__", None)
    if name is None:
        return None
    module = (getattr(fn, "__module__", None) or "").split(".")[0]
    is_np_ufunc = type(fn).__name__ == "ufunc"
    if not (is_np_ufunc or module in _TRUSTED_MODULES):
        return None
    name = _ALIASES.get(name, name)
    return name if name in table else None


def resolve_by_name(fn):
    """Index into al_ufunc when `fn` is a known function object or name, else None (no probing)."""
    if isinstance(fn, str):
        name = _ALIASES.get(fn, fn)
        return UFUNCS.index(name) if name in UFUNCS else None
    name = _by_name(fn, UFUNCS)
    return UFUNCS.index(name) if name is not None else None


def resolve_unary(fn) -> int:
    """Index into al_ufunc for `fn`, or NotImplementedError."""
    if isinstance(fn, str):
        name = _ALIASES.get(fn, fn)
        if name in UFUNCS:
            return UFUNCS.index(name)
        raise NotImplementedError(f"apply(): no device ufunc named {fn!r}")
    name = _by_name(fn, UFUNCS)
    if name is not None:
        return UFUNCS.index(name)
    if callable(fn):
        got = [_safe(fn, x) for x in _UNARY_PROBES]
        if sum(not math.isnan(g) for g in got) >= 3:
            for cand in UFUNCS:
                want = [_safe(_HOST[cand], x) for x in _UNARY_PROBES]
                if all(_same(g, w) for g, w in zip(got, want)):
                    return UFUNCS.index(cand)
    raise NotImplementedError(
        "apply(): the callable does not resolve to a named device ufunc "
        f"({', '.join(UFUNCS)}); arbitrary Python callables cannot run on the GPU and there is no CPU fallback")


def resolve_reduce(fn) -> int:
    """Index into al_redop for the binary accumulator `fn`, or NotImplementedError."""
    if isinstance(fn, str):
        if fn in REDOPS:
            return REDOPS.index(fn)
        raise NotImplementedError(f"reduce(): no device accumulator named {fn!r}")
    by_id = {operator.add: "sum", operator.mul: "prod", max: "max", min: "min", math.fsum: None}
    if fn in by_id and by_id[fn]:
        return REDOPS.index(by_id[fn])
    name = getattr(fn, "__name__", "")
    if type(fn).__name__ == "ufunc":
        table = {"add": "sum", "multiply": "prod", "maximum": "max", "minimum": "min", "fmax": "max", "fmin": "min"}
        if name in table:
            return REDOPS.index(table[name])
    if callable(fn):
        got = [_safe(fn, x, y) for x, y in _BINARY_PROBES]
        if not any(math.isnan(g) for g in got):
            for cand in REDOPS:
                if all(_same(g, float(_HOST_BIN[cand](x, y))) for g, (x, y) in zip(got, _BINARY_PROBES)):
                    return REDOPS.index(cand)
    raise NotImplementedError(
        "reduce(): the accumulator does not resolve to a named device reduction (sum, max, min, prod); "
        "arbitrary Python callables cannot run on the GPU and there is no CPU fallback")
