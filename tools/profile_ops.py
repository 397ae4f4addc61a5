"""Runs every op of the bench suite a few times at the BASELINE shapes (for `ncu`; not a benchmark)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import arraylib as al  # noqa: E402
from tools.microbench import rnd  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = (1 << 28) // scale
a, b = rnd((n,), 1), rnd((n,), 2)
L = 1024 // scale
A, B, c = rnd((L, 1, 4096), 1), rnd((1, 1024, 4096), 2), rnd((4096,), 3)
X = rnd((64 // scale, 512, 512, 64), 4)
m = 16384 if scale == 1 else 8192
X5, Y5 = rnd((m, m), 5), rnd((m, m), 6)
# shapes off the BASELINE configs that exercise the fallback kernels (ew_run modes, reduce_inner_peel, reduce_window_any)
P3, p3 = rnd(((1 << 26) // scale, 3), 7), rnd((3,), 8)
R255 = rnd(((1 << 20) // scale, 255), 9)
# round 3: period-table kernel (short inner axes), odd-extent transpose (4-byte tile), FP64 ufuncs and pow
PA, PB = rnd(((1 << 22) // scale, 1, 7), 10), rnd((1, 9, 7), 11)
QA, QB = rnd((4096 // scale, 1, 5, 3), 12), rnd((1, 2048, 1, 3), 13)
mo = m - 1
XO = rnd((mo, mo), 14)
pos = al.abs(a) + 0.5
for rep in range(2):
    r = a + b; del r
    r = a * b; del r
    r = a + 1.0; del r
    t = A + B
    r = t + c; del r, t
    r = X.reduce_sum(dims=(1, 2)); del r
    r = X.reduce_sum(dims=(3,)); del r
    r = X.reduce_sum(dims=(0,)); del r
    r = a.as_strided(shape=(n - 63, 64), stride=(1, 1)).reduce_sum(dims=[1]); del r
    r = X5.transpose((1, 0)) + Y5; del r
    r = X5.transpose((1, 0)).reshape((m * m,)); del r
    r = (al.lazy(A) + B + c).eval(); del r
    r = P3 + p3; del r
    r = R255.reduce_sum(dims=(1,)); del r
    r = a.as_strided(shape=(n - 99, 100), stride=(1, 1)).reduce_sum(dims=[1]); del r
    r = PA + PB; del r
    r = QA + QB; del r
    r = XO.transpose((1, 0)).reshape((mo * mo,)); del r
    r = al.exp(a); del r
    r = al.log(pos); del r
    r = al.sin(a); del r
    r = al.pow(pos, b); del r
al.synchronize()
print("profile_ops done")
