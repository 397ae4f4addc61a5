"""arraylib_b200: B200 (sm_100a) implementation of the arraylib array core.

`import arraylib_b200 as al` (or `import arraylib as al` through the shim package) gives the same
names as the reference's `arraylib` package; buffers live in GPU HBM behind libarraylib_b200.so.
"""
from arraylib_b200.core import *  # noqa: F401,F403
from arraylib_b200.core import NDArray, Layout, synchronize, full, mod  # noqa: F401
from arraylib_b200 import core, ufuncs  # noqa: F401
from arraylib_b200.lazy import Expr, fused, lazy  # noqa: F401,E402
