"""pow / log / exp on 2^27 in-domain elements (for `ncu`; not a benchmark)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import arraylib as al  # noqa: E402
from tools.microbench import rnd  # noqa: E402

n = 1 << 27
a, b = rnd((n,), 1), rnd((n,), 2)
pos = al.abs(a) + 0.5
for _ in range(2):
    r = al.pow(pos, b); del r
    r = al.log(pos); del r
    r = al.exp(a); del r
    r = al.pow(pos, 2.0); del r
al.synchronize()
print("profile_pow done")
