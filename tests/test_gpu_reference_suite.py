"""The reference's own test-suite cases (/root/reference/tests/test_array.py), restated against the
new package: same inputs, same NumPy oracle, exact equality. The file is not copied -- each case
names the reference lines it covers."""
import operator
from itertools import product

import numpy as np
import numpy.testing as npt
import pytest

pytestmark = pytest.mark.gpu

SHAPE4 = (3, 4, 5, 6)
N4 = 360


def same(al, got, want):
    npt.assert_array_equal(al.to_numpy(got), np.asarray(want).astype(np.float32))


def base(al):
    return al.reshape(al.arange(1, 1 + N4), SHAPE4), np.arange(1, 1 + N4).reshape(SHAPE4)


@pytest.mark.parametrize("data", [[[1, 2, 3], [4, 5, 6]], [[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]], [[1]]])
def test_creation(al, data):  # test_array.py:31-42
    same(al, al.array(data), np.array(data))


@pytest.mark.parametrize("shape", [(2, 3), (3, 2), (1, 1)])
def test_ones_zeros(al, shape):  # :45-70
    same(al, al.ones(shape), np.ones(shape))
    same(al, al.zeros(shape), np.zeros(shape))


@pytest.mark.parametrize("args", [(0, 5, 1), (1, 10, 2)])
def test_arange(al, args):  # :73-77
    same(al, al.arange(*args), np.arange(*args))


@pytest.mark.parametrize("args", [(0, 10, 5), (1, 100, 10)])
def test_linspace(al, args):  # :80-84
    same(al, al.linspace(*args), np.linspace(*args))


@pytest.mark.parametrize("fn", [operator.add, operator.sub, operator.mul, operator.truediv])
def test_elementwise(al, fn):  # :87-103
    a, b = [1, 2, 3], [4, 5, 6]
    same(al, fn(al.array(a), al.array(b)), fn(np.array(a, np.float32), np.array(b, np.float32)))


@pytest.mark.parametrize("a,b", [([[1, 2], [3, 4]], [[5, 6], [7, 8]]), ([[1, 2, 3], [4, 5, 6]], [[7, 8], [9, 10], [11, 12]])])
def test_matmul(al, a, b):  # :106-122
    same(al, al.array(a) @ al.array(b), np.array(a) @ np.array(b))


@pytest.mark.parametrize("a,b", [([1, 2, 3], [4, 5, 6]), ([-1, 2, 3], [4, -5, 0])])
def test_dot(al, a, b):  # :125-138
    same(al, al.dot(al.array(a), al.array(b)), np.dot(np.array(a), np.array(b)).reshape(1))


@pytest.mark.parametrize("data,shape", [([[1, 2, 3], [4, 5, 6]], (3, 2)), ([[1, 2], [3, 4], [5, 6]], (2, 3))])
def test_reshape_ravel(al, data, shape):  # :141-155, :188-202
    same(al, al.reshape(al.array(data), shape), np.reshape(np.array(data), shape))
    same(al, al.ravel(al.array(data)), np.ravel(np.array(data)))


def test_transpose_move_axis(al):  # :158-185
    a, n = base(al)
    same(al, al.transpose(a, (1, 2, 0, 3)), np.transpose(n, (1, 2, 0, 3)))
    for src, dst in [((1, 2), (2, 1)), ((3, 0), (1, 2))]:
        same(al, al.move_dim(a, src, dst), np.moveaxis(n, src, dst))


@pytest.mark.parametrize("data,index,value", [([[1, 2, 3], [4, 5, 6]], (0, 1), 10), ([[1, 2], [3, 4]], (1, 0), 5)])
def test_getitem_setitem(al, data, index, value):  # :205-219
    a = al.array(data)
    assert a[index] == data[index[0]][index[1]]
    a[index] = value
    data[index[0]][index[1]] = value
    same(al, a, np.array(data))


def test_get_and_set_every_index(al):  # :255-270 (360 cases each, folded into one test)
    a, n = base(al)
    for index in product(range(3), range(4), range(5), range(6)):
        assert a[index] == n[index]
    for index in list(product(range(3), range(4), range(5), range(6)))[::7]:
        b, m = base(al)
        b[index] = 10
        m[index] = 10
        same(al, b, m)


@pytest.mark.parametrize("fn", [np.log, np.exp, np.negative, np.sin, np.cos, np.tan, np.arctan, np.arcsin])
def test_unary_through_apply(al, fn):  # :273-283 -- callables resolve to the named device ufuncs
    a, n = base(al)
    with np.errstate(all="ignore"):
        same(al, al.apply(fn, a), fn(n))


def test_reduce_with_lambda(al):  # :286-293
    a, n = base(al)
    same(al, al.reduce(lambda x, y: x + y, a, dims=(1, 2), init=0), n.sum(axis=(1, 2), keepdims=True))


def test_where(al):  # :296-315
    a, n = base(al)
    cond = al.lt(a, 10.0)
    same(al, al.where(cond, a - 10.0, a + 10.0), np.where(n < 10.0, n - 10.0, n + 10.0))
    same(al, al.where(cond, a - 10.0, 10.0), np.where(n < 10.0, n - 10.0, 10.0))
    same(al, al.where(cond, 10.0, a - 10.0), np.where(n < 10.0, 10.0, n - 10.0))


@pytest.mark.parametrize("ls,rs", [((1, 2, 3), (1, 1, 2, 3)), ((1, 2, 3), (1, 2, 3)), ((1, 2, 3), (2, 3)), ((1, 2, 3), (3,)),
                                   ((1, 2, 3), (1, 3)), ((1, 2, 3), (2, 1)), ((1, 2, 3), (1, 1))])
def test_broadcast(al, ls, rs):  # :318-334
    same(al, al.ones(ls) + al.ones(rs), np.ones(ls) + np.ones(rs))


def test_numpy_round_trip(al):  # :337-348
    n = (np.arange(1, 1 + N4).reshape(SHAPE4) / 10.0).astype(np.float32)
    same(al, al.from_numpy(n), n)
    a = al.reshape(al.arange(1, 1 + N4) / 10.0, SHAPE4)
    same(al, a, al.to_numpy(a))
    npt.assert_array_equal(np.asarray(a), n)


@pytest.mark.parametrize("lhs,rhs,dim", [((2, 4), (3, 4), 0), ((1, 6, 3), (1, 4, 3), 1), ((1, 2, 3), (1, 2, 4), 2)])
def test_cat(al, lhs, rhs, dim):  # :351-384
    pl, pr = int(np.prod(lhs)), int(np.prod(rhs))
    got = al.cat([al.reshape(al.arange(pl), lhs), al.reshape(al.arange(pr), rhs)], [dim])
    same(al, got, np.concatenate([np.arange(pl).reshape(lhs), np.arange(pr).reshape(rhs)], axis=dim))


def test_nd_cat(al):
    a, b = al.reshape(al.arange(1, 4), (1, 3)), al.reshape(al.arange(1, 5), (2, 2))
    an, bn = np.arange(1, 4).reshape(1, 3), np.arange(1, 5).reshape(2, 2)
    same(al, al.cat([a, b], [0, 1]), np.block([[an, np.zeros((1, 2))], [np.zeros((2, 3)), bn]]))
