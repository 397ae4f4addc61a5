"""CPU tests of host-side logic: callable -> device-ufunc resolution."""
import math
import os
import operator

import numpy as np
import pytest

from arraylib_b200 import ufuncs
from arraylib_b200._lib import REDOPS, UFUNCS


@pytest.mark.parametrize("fn,name", [
    (np.log, "log"), (np.exp, "exp"), (np.negative, "neg"), (np.sin, "sin"), (np.cos, "cos"), (np.tan, "tan"),
    (np.arctan, "atan"), (np.arcsin, "asin"), (math.sqrt, "sqrt"), (abs, "abs"), (operator.neg, "neg"),
    (math.floor, "floor"), (np.ceil, "ceil"), (np.absolute, "abs"), (np.arccosh, "acosh"),
    (lambda x: math.sqrt(x), "sqrt"), (lambda x: -x, "neg"), (lambda x: np.log(x), "log"), (lambda v: math.tanh(v), "tanh"),
    ("sqrt", "sqrt"), ("arctan", "atan"),
])
def test_unary_resolution(fn, name):
    assert UFUNCS[ufuncs.resolve_unary(fn)] == name


@pytest.mark.parametrize("fn", [lambda x: x * x + 1, lambda x: 2 * x, lambda x: "a", print, "nope"])
def test_unary_unresolvable_raises(fn):
    with pytest.raises(NotImplementedError):
        ufuncs.resolve_unary(fn)


@pytest.mark.parametrize("fn,name", [
    (lambda x, y: x + y, "sum"), (operator.add, "sum"), (max, "max"), (min, "min"), (np.maximum, "max"),
    (lambda a, b: a * b, "prod"), (lambda a, b: np.maximum(a, b), "max"), (lambda acc, v: v + acc, "sum"), (np.add, "sum"), ("min", "min"),
])
def test_reduce_resolution(fn, name):
    assert REDOPS[ufuncs.resolve_reduce(fn)] == name


@pytest.mark.parametrize("fn", [
    lambda a, b: a - b,                      # not an accumulator the device has
    lambda a, b: a if a > b else b,          # data-dependent branch: cannot be traced, and is NOT guessed from samples
    lambda a, b: max(a, b),                  # builtin max compares and branches
    lambda a, b: a + b + 0.0,                # two instructions
    lambda a, b: a + a,                      # ignores one argument
])
def test_reduce_unresolvable_raises(fn):
    with pytest.raises(NotImplementedError):
        ufuncs.resolve_reduce(fn)


@pytest.mark.parametrize("fn", [
    lambda x: math.exp(min(x, 50)),                      # equals exp on any small probe set, differs for x > 50
    lambda x: math.sqrt(x) if x < 100 else 10.0,         # piecewise
    lambda x: math.sqrt(float(x)),                       # float() of the proxy
])
def test_piecewise_callables_are_refused_not_guessed(fn):
    with pytest.raises(NotImplementedError):
        ufuncs.resolve_unary(fn)


def test_math_module_is_restored_after_tracing():
    before = (math.sqrt, math.exp, math.pow, math.fmod, math.floor)
    ufuncs.resolve_unary(lambda x: math.sqrt(x))
    with pytest.raises(NotImplementedError):
        ufuncs.resolve_unary(lambda x: math.sqrt(x) if x > 0 else 0.0)
    assert (math.sqrt, math.exp, math.pow, math.fmod, math.floor) == before
    assert math.sqrt(4.0) == 2.0


def test_traced_comparisons_and_mod_build_instructions():
    """`x == c` / `x != c` / `x % c` inside a traced lambda must become tape instructions, not Python
    bools baked in as constants."""
    import importlib
    lazy = importlib.import_module("arraylib_b200.lazy")  # (the package attribute `lazy` is the function)
    from arraylib_b200._lib import BINOPS
    for fn, op in ((lambda x: x == 2, "eq"), (lambda x: x != 0, "ne"), (lambda x: x % 3, "mod"), (lambda x: 7 % x, "mod"),
                   (lambda x: math.fmod(x, 3), "mod"), (lambda x: math.pow(x, 2), "pow")):
        trace, out = lazy._run_traced(fn, 1)
        assert len(trace.insns) == 1 and trace.insns[0][0] == 0 and BINOPS[trace.insns[0][1]] == op
    trace, out = lazy._run_traced(lambda x: x * (x != 0), 1)
    assert [BINOPS[i[1]] for i in trace.insns] == ["ne", "mul"]
    with pytest.raises(TypeError):
        hash(lazy.TExpr(trace, -1))


def test_streaming_host_copy_matches_memcpy(tmp_path):
    """csrc/hostcopy.cpp: the staging ring's copy with non-temporal stores (head / 128-byte body / tail split) must move
    exactly the bytes memcpy moves, for any size and any source / destination alignment, and touch nothing else."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "hostcopy_harness"
    subprocess.check_call(["g++", "-O2", "-mavx2", "-o", str(exe), os.path.join(root, "tests", "csrc", "hostcopy_harness.cpp"),
                           os.path.join(root, "arraylib_b200", "csrc", "hostcopy.cpp")])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "hostcopy ok" in out.stdout, out.stdout[-500:]
