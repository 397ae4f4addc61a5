"""Drop-in name for the reference package: `import arraylib as al` resolves to arraylib_b200.

One GPU (default): the single-device API (arraylib_b200.core). Several GPUs of one box, opt-in through the
environment (env knobs only, `al.*` keeps no config API):

  ARRAYLIB_DEVICES=N python script.py     this import re-launches the command once per GPU
                                          (arraylib_b200.launch) and exits with the workers' status;
  ARRAYLIB_SPMD=1 under torchrun / `python -m arraylib_b200.launch --spmd`: the ranks already exist.

In both cases the names below come from arraylib_b200.spmd: global arrays sharded along their leading axis."""
import os as _os
import sys as _sys


def _devices() -> int:
    try:
        return int(_os.environ.get("ARRAYLIB_DEVICES", "1") or 1)
    except ValueError:
        return 1


if _devices() > 1 and "RANK" not in _os.environ:
    from arraylib_b200.launch import launch as _launch

    _argv = list(getattr(_sys, "orig_argv", [_sys.executable] + _sys.argv))[1:]
    _sys.stdout.flush()
    _sys.exit(_launch(_devices(), _argv, quiet_workers=_os.environ.get("ARRAYLIB_QUIET_WORKERS", "1") != "0",
                      env_extra={"ARRAYLIB_SPMD": "1"}))

if int(_os.environ.get("WORLD_SIZE", "1") or 1) > 1 and _os.environ.get("ARRAYLIB_SPMD") == "1":
    from arraylib_b200 import dist as _dist

    if _dist.world() == 1:
        _dist.init_from_env()
    from arraylib_b200.spmd import *  # noqa: F401,F403
    from arraylib_b200.spmd import NDArray, Layout  # noqa: F401
    from arraylib_b200 import core, dist, spmd  # noqa: F401
else:
    from arraylib_b200.core import *  # noqa: F401,F403
    from arraylib_b200.core import NDArray, Layout, synchronize, full, mod  # noqa: F401
    from arraylib_b200 import core  # noqa: F401  (so `arraylib.core` keeps working)
    from arraylib_b200.lazy import Expr, fused, lazy  # noqa: F401,E402
