"""Start one process per GPU without PyTorch:

    python -m arraylib_b200.launch --nproc N script.py [args...]

Every worker gets the torchrun-style variables RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT and a
private rendezvous directory (ARRAYLIB_RDZV_DIR) through which `arraylib_b200.dist.init_from_env()` ships the
NCCL id. The launcher exits with the first non-zero worker status and stops the other workers (its own process
group only). `torchrun` works just as well: `dist.init_from_env()` only reads the environment.
"""
from __future__ import annotations

import argparse
import os
import shutil
import signal
import subprocess
import sys
import tempfile
import time


def launch(nproc: int, argv: list, *, timeout: float | None = None, env_extra: dict | None = None, quiet_workers: bool = False) -> int:
    """Run `python argv...` once per rank. `quiet_workers`: only rank 0 keeps its stdout (SPMD scripts print the same
    thing on every rank)."""
    rdzv = tempfile.mkdtemp(prefix="arraylib_b200_rdzv_")
    procs = []
    try:
        for rank in range(nproc):
            env = dict(os.environ)
            env.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(nproc), LOCAL_WORLD_SIZE=str(nproc),
                       MASTER_ADDR="127.0.0.1", MASTER_PORT=env.get("MASTER_PORT", "29511"), ARRAYLIB_RDZV_DIR=rdzv)
            env.update(env_extra or {})
            out = subprocess.DEVNULL if quiet_workers and rank > 0 else None
            procs.append(subprocess.Popen([sys.executable] + list(argv), env=env, start_new_session=True, stdout=out))
        deadline = None if timeout is None else time.monotonic() + timeout
        status = 0
        live = set(range(nproc))
        while live:
            for r in sorted(live):
                rc = procs[r].poll()
                if rc is None:
                    continue
                live.discard(r)
                if rc != 0 and status == 0:
                    status = rc
            if status != 0 or (deadline is not None and time.monotonic() > deadline):
                if status == 0:
                    status = 124
                break
            time.sleep(0.05)
        return status
    finally:
        for p in procs:
            if p.poll() is None:
                try:
                    os.killpg(p.pid, signal.SIGTERM)  # each worker leads its own session: only our children
                except ProcessLookupError:
                    pass
        for p in procs:
            try:
                p.wait(timeout=10)
            except subprocess.TimeoutExpired:
                try:
                    os.killpg(p.pid, signal.SIGKILL)
                except ProcessLookupError:
                    pass
        shutil.rmtree(rdzv, ignore_errors=True)


def main() -> None:
    ap = argparse.ArgumentParser(prog="python -m arraylib_b200.launch")
    ap.add_argument("--nproc", type=int, required=True)
    ap.add_argument("--timeout", type=float, default=None)
    ap.add_argument("--quiet-workers", action="store_true", help="only rank 0 keeps its stdout")
    ap.add_argument("--spmd", action="store_true", help="export ARRAYLIB_SPMD=1: `import arraylib` gives the sharded al.* front end (arraylib_b200.spmd)")
    ap.add_argument("script")
    ap.add_argument("args", nargs=argparse.REMAINDER)
    ns = ap.parse_args()
    sys.exit(launch(ns.nproc, [ns.script] + ns.args, timeout=ns.timeout, quiet_workers=ns.quiet_workers,
                    env_extra={"ARRAYLIB_SPMD": "1"} if ns.spmd else None))


if __name__ == "__main__":
    main()
