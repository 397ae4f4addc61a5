"""Runs the C5a window reduction under each kernel variant (for `ncu`; not a benchmark)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import arraylib as al  # noqa: E402
from tools.microbench import rnd, tune  # noqa: E402

n, w = 1 << 28, int(sys.argv[1]) if len(sys.argv) > 1 else 64
base = rnd((n,), 1)
for variant, per_sm in ((1, 2), (1, 3), (0, 0)):
    tune(window_shfl=variant, window_ctas_per_sm=per_sm)
    for _ in range(2):
        r = base.as_strided(shape=(n - w + 1, w), stride=(1, 1)).reduce_sum(dims=[1])
        del r
r = base + 1.0   # the plain 1:1 stream for comparison (ew_nd)
del r
al.synchronize()
print("profile_window done")
