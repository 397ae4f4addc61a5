"""binary max / min against add on 2^28 elements (for `ncu`; not a benchmark)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import arraylib as al  # noqa: E402
from tools.microbench import rnd  # noqa: E402

n = 1 << 28
a, b = rnd((n,), 1), rnd((n,), 2)
for _ in range(2):
    r = al.add(a, b); del r
    r = al.max(a, b); del r
    r = al.min(a, b); del r
al.synchronize()
print("profile_maxmin done")
