"""ncu target: the fused C3 chain at a quarter of the BASELINE leading extent."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import arraylib as al  # noqa: E402
from tools.microbench import rnd  # noqa: E402

A, B, c = rnd((256, 1, 4096), 1), rnd((1, 1024, 4096), 2), rnd((4096,), 3)
for _ in range(3):
    r = (al.lazy(A) + B + c).eval()
    del r
al.synchronize()
print("done")
