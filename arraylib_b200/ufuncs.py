"""Resolution of Python callables to entries of the device ufunc / reduce tables.

The reference's `apply(fn, a)` and `reduce(fn, a, ...)` wrap ANY Python callable as a C callback
and invoke it once per element from the CPU loop (/root/reference/arraylib/core.py:455-460,
661-673). A Python callable cannot run inside a CUDA kernel, so here the callable has to resolve
to one of the named device kernels (include/arraylib_b200.h: al_ufunc / al_redop):

  1. by identity or name, for `math.*`, `numpy.*` ufuncs, `operator.*` and builtins;
  2. otherwise by TRACING: the callable is run once on scalar proxies (lazy.TExpr; `math.*` is
     made traceable for the duration) and must record exactly one instruction over its arguments
     (`lambda x: math.sqrt(x)`, `lambda acc, v: acc + v`). Longer expressions go to the FP64 tape
     (core.apply), never to a guessed table entry.

Anything that does not resolve raises NotImplementedError -- there is no silent CPU fallback, and
no guessing from sample values either: round 1 matched callables against table entries on a handful
of probe scalars, which could silently pick `exp` for `lambda x: math.exp(min(x, 50))`.
"""
from __future__ import annotations

import operator

from ._lib import REDOPS, UFUNCS

_ALIASES = {
    "negative": "neg", "neg": "neg", "absolute": "abs", "fabs": "abs", "abs": "abs",
    "arcsin": "asin", "arccos": "acos", "arctan": "atan", "arcsinh": "asinh", "arccosh": "acosh",
    "arctanh": "atanh",
}
_TRUSTED_MODULES = {"math", "numpy", "operator", "builtins", "_operator"}


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
        from .lazy import single_instruction
        got = single_instruction(fn, 1)
        if got is not None and got[0] == 1:
            return UFUNCS.index(got[1])
    raise NotImplementedError(
        "apply(): the callable does not resolve to a named device ufunc "
        f"({', '.join(UFUNCS)}); arbitrary Python callables cannot run on the GPU and there is no CPU fallback")


_REDUCE_OF_BINOP = {"sum": "sum", "mul": "prod", "max": "max", "min": "min"}


def resolve_reduce(fn) -> int:
    """Index into al_redop for the binary accumulator `fn`, or NotImplementedError."""
    if isinstance(fn, str):
        if fn in REDOPS:
            return REDOPS.index(fn)
        raise NotImplementedError(f"reduce(): no device accumulator named {fn!r}")
    by_id = {operator.add: "sum", operator.mul: "prod", max: "max", min: "min"}
    try:
        if fn in by_id:
            return REDOPS.index(by_id[fn])
    except TypeError:  # unhashable callable
        pass
    name = getattr(fn, "__name__", "")
    if type(fn).__name__ == "ufunc":
        table = {"add": "sum", "multiply": "prod", "maximum": "max", "minimum": "min", "fmax": "max", "fmin": "min"}
        if name in table:
            return REDOPS.index(table[name])
    if callable(fn):
        from .lazy import single_instruction
        got = single_instruction(fn, 2)
        if got is not None and got[0] == 0 and got[1] in _REDUCE_OF_BINOP:
            return REDOPS.index(_REDUCE_OF_BINOP[got[1]])
    raise NotImplementedError(
        "reduce(): the accumulator does not resolve to a named device reduction (sum, max, min, prod); pass "
        "operator.add / max / min / np.maximum ... or a lambda made of one such operation. Arbitrary Python "
        "callables (data-dependent branches included) cannot run on the GPU and there is no CPU fallback")
