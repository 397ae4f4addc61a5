"""Round-3 ncu targets: the period-table kernel on two write-bound broadcasts and four FP64 ufuncs (not a benchmark)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import arraylib as al  # noqa: E402
from tools.microbench import rnd  # noqa: E402

for sa, sb in (((1 << 22, 1, 7), (1, 9, 7)), ((1 << 12, 1, 5, 3), (1, 1 << 11, 1, 3)), ((1 << 14, 1, 3), (1, 1 << 12, 3))):
    A, B = rnd(sa, 1), rnd(sb, 2)
    for _ in range(2):
        r = A + B
        del r
    del A, B
n = 1 << 27
a = rnd((n,), 1)
pos = al.abs(a) + 0.5
for _ in range(2):
    for name, x in (("exp", a), ("log", pos), ("sin", a), ("asinh", a)):
        r = getattr(al, name)(x)
        del r
al.synchronize()
print("profile_r3 done")
