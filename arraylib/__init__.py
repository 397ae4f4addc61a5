"""Drop-in name for the reference package: `import arraylib as al` resolves to arraylib_b200."""
from arraylib_b200.core import *  # noqa: F401,F403
from arraylib_b200.core import NDArray, Layout, synchronize, full, mod  # noqa: F401
from arraylib_b200 import core  # noqa: F401  (so `arraylib.core` keeps working)
from arraylib_b200.lazy import Expr, fused, lazy  # noqa: F401,E402
