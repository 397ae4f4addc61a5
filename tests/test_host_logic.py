"""CPU tests of host-side logic: callable -> device-ufunc resolution."""
import math
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
    (lambda a, b: a * b, "prod"), (lambda a, b: a if a > b else b, "max"), (np.add, "sum"), ("min", "min"),
])
def test_reduce_resolution(fn, name):
    assert REDOPS[ufuncs.resolve_reduce(fn)] == name


def test_reduce_unresolvable_raises():
    with pytest.raises(NotImplementedError):
        ufuncs.resolve_reduce(lambda a, b: a - b)
