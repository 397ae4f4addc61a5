#!/usr/bin/env python
"""bench.py -- effective HBM GB/s of the array-core hot path on N B200s (BASELINE.json metric).

One STEP = one pass of the hot-path suite over synthetic float32 arrays of the BASELINE shapes:
  C2  a+b, a*b, a+1.0                      two 2^28-element contiguous arrays
  C3  t = A[1024,1,4096] + B[1,1024,4096]; u = t + c[4096]      (2^32-element outputs, eager)
  C4  X[64,512,512,64]: reduce_sum(dims=(1,2)), (3,), (0,)
  C5  window-64 as_strided view of 2^28 elements reduce_sum(dims=[1]);  X[16384,16384].T + Y;
      X.T.reshape(-1) (forces the strided-copy kernel)
`value` = algorithmic bytes of all ops (SURVEY.md section 8d / BASELINE.md section 5) of the WHOLE
job / device time of the slowest rank (CUDA events on the library's stream).

Multi-GPU (`--gpus N`, one process per GPU, torchrun or `python -m arraylib_b200.launch`; no PyTorch
anywhere in this file): `--scaling strong` (default) cuts the BASELINE shapes along their leading axis
over the N ranks (BASELINE.json config 3: "sharded by leading dim over 8 GPUs"): C2 2^28/N elements,
C3 1024/N rows of A with B and c replicated, C4 64/N slabs, C5a (2^28-63)/N windows + a 63-element halo,
C5b 16384/N output rows (X replicated: every rank reads its own COLUMN block of it). The only exchange is
reduce_sum(dims=(0,)) of C4: every rank reduces its slab and the partials meet over NVLink peer memory,
the result staying sharded along the leading non-reduced axis (reduce-scatter); the replicated
(allreduce) variant and the local part are timed next to it. `--scaling weak` gives every rank a full
copy of the named shapes (round-1 behaviour); a short weak run is also reported under "weak".

`--impl reference` times the UNMODIFIED reference (oracle/_ref/arraylib_ref_omp.so, built with
the reference Makefile's flags + -fopenmp) on the host cores, on a bounded sample of the same
suite (BASELINE.md section 4.5 sizes when the time budget allows). The same sample is reported as
`cpu_baseline` in the default arm.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "broadcast-add & reduce_sum effective HBM GB/s (% of 8 TB/s) at 1/2/4/8 B200"
NOMINAL_GBS = 8000.0
OP_NAMES = ("c2_add", "c2_mul", "c2_add_scalar", "c3_bcast_add", "c3_bcast_add_vec", "c4_reduce_dims12", "c4_reduce_dim3",
            "c4_reduce_dim0", "c5a_window_reduce", "c5b_transposed_add", "c5b_transposed_copy")


# ---------------------------------------------------------------------------------------------------
# workload definition (shared by the GPU arm, the e2e leg and the CPU reference arm)
# ---------------------------------------------------------------------------------------------------
def suite_shapes(scale: int):
    """GLOBAL shapes of the suite; scale=1 is the BASELINE configuration, larger values shrink it."""
    s = max(scale, 1)
    n2 = (1 << 28) // s
    l3 = max(1024 // s, 8)
    l4 = max(64 // s, 8)
    n5 = (1 << 28) // s
    m5 = 16384 if s == 1 else max(int(16384 / (s ** 0.5)) // 64 * 64, 64)
    return dict(n2=n2, c3=(l3, 1024 if s == 1 else max(1024 // s, 8), 4096), c4=(l4, 512, 512, 64) if s == 1 else (l4, max(512 // s, 4), 512, 64),
                n5=n5, m5=m5, window=64)


def algorithmic_bytes(sh):
    n2, (l3, j3, w3), c4, n5, m5, w = sh["n2"], sh["c3"], sh["c4"], sh["n5"], sh["m5"], sh["window"]
    x4 = c4[0] * c4[1] * c4[2] * c4[3]
    out3 = l3 * j3 * w3
    return {
        "c2_add": 12 * n2, "c2_mul": 12 * n2, "c2_add_scalar": 8 * n2,
        "c3_bcast_add": 4 * out3 + 4 * (l3 * w3 + j3 * w3),
        "c3_bcast_add_vec": 8 * out3 + 4 * w3,
        "c4_reduce_dims12": 4 * x4 + 4 * c4[0] * c4[3],
        "c4_reduce_dim3": 4 * x4 + 4 * (x4 // c4[3]),
        "c4_reduce_dim0": 4 * x4 + 4 * (x4 // c4[0]),
        "c5a_window_reduce": 4 * n5 + 4 * (n5 - w + 1),
        "c5b_transposed_add": 12 * m5 * m5,
        "c5b_transposed_copy": 8 * m5 * m5,
    }


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.device)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=3)
            except Exception:
                self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
class Suite:
    """The eleven ops over one rank's share of the suite. `mode` is "strong" (global shapes cut over the ranks)
    or "weak" (every rank owns a full copy of the named shapes)."""

    def __init__(self, al, dist, sh, rank, world, mode):
        import numpy as np
        self.al, self.dist, self.np = al, dist, np
        self.rank, self.world, self.mode = rank, world, mode
        self.sh = sh
        div = world if mode == "strong" else 1
        n2, c3, c4, n5, m5, w = sh["n2"], sh["c3"], sh["c4"], sh["n5"], sh["m5"], sh["window"]
        for what, extent in (("c2_n", n2), ("c3 rows", c3[0]), ("c4 slabs", c4[0]), ("c5b rows", m5)):
            assert extent % div == 0, f"{what} = {extent} does not divide over {div} ranks"
        self.n2 = n2 // div
        self.c3 = (c3[0] // div, c3[1], c3[2])
        self.c4 = (c4[0] // div,) + tuple(c4[1:])
        if div > 1:  # windows that START in this rank's block + the halo behind them
            lo, hi = dist.window_halo_bounds(n5 - w + 1, w, rank, world)
            self.n5_local, self.win_out = hi - lo, hi - lo - w + 1
        else:
            self.n5_local, self.win_out = n5, n5 - w + 1
        self.m5, self.m5_rows = m5, m5 // div
        self.row_lo = rank * self.m5_rows if div > 1 else 0
        self.w = w
        self.sharded = div > 1
        self.global_bytes = algorithmic_bytes(sh)
        self.step_bytes = sum(self.global_bytes.values()) * (1 if mode == "strong" else world)

    # -- synthetic inputs: one pinned host pool of random normals, sliced for every operand --------------
    def make_host_inputs(self):
        np, al = self.np, self.al
        x4 = int(np.prod(self.c4))
        pool_n = max(x4, 2 * self.n2, 2 * self.m5 * self.m5 if not self.sharded else self.m5 * self.m5 + self.m5_rows * self.m5,
                     self.n5_local, 2 * self.c3[0] * self.c3[2] + self.c3[1] * self.c3[2])
        hin = al.host_buffer((pool_n,))
        rng = np.random.default_rng(self.rank)
        block = rng.standard_normal(min(pool_n, 1 << 24), dtype=np.float32)
        for off in range(0, pool_n, block.size):
            hin[off:off + block.size] = block[: min(block.size, pool_n - off)]

        def host_in(offset, shape):
            n = int(np.prod(shape))
            return hin[offset:offset + n].reshape(shape)

        c3, m5 = self.c3, self.m5
        self.host_inputs = {
            "a": host_in(0, (self.n2,)), "b": host_in(self.n2, (self.n2,)),
            "A": host_in(0, (c3[0], 1, c3[2])), "B": host_in(c3[0] * c3[2], (1, c3[1], c3[2])), "c": host_in(7, (c3[2],)),
            "X": host_in(0, self.c4), "base": host_in(0, (self.n5_local,)),
            "X5": host_in(0, (m5, m5)), "Y5": host_in(m5 * m5 if self.sharded else m5 * m5, (self.m5_rows, m5)),
        }
        self.h2d_bytes = sum(v.nbytes for v in self.host_inputs.values())
        return self.host_inputs

    def out_elems(self):
        np = self.np
        x4 = int(np.prod(self.c4))
        c3 = self.c3
        dim0 = x4 // self.c4[0]
        if self.sharded and self.mode == "strong":
            dim0 //= self.world  # the result stays sharded
        return {"c2_add": self.n2, "c2_mul": self.n2, "c2_add_scalar": self.n2, "c3_bcast_add_vec": c3[0] * c3[1] * c3[2],
                "c4_reduce_dims12": self.c4[0] * self.c4[3], "c4_reduce_dim3": x4 // self.c4[3], "c4_reduce_dim0": dim0,
                "c5a_window_reduce": self.win_out, "c5b_transposed_add": self.m5_rows * self.m5, "c5b_transposed_copy": self.m5_rows * self.m5}

    def make_ops(self, d):
        """name -> thunk returning the op's result (inputs are device arrays in `d`)."""
        al, dist = self.al, self.dist
        keep = {}
        m5, rows, lo, w = self.m5, self.m5_rows, self.row_lo, self.w

        def c3_first():
            keep["t"] = d["A"] + d["B"]
            return keep["t"]

        def c3_second():
            return keep.pop("t") + d["c"]

        def xt():  # this rank's rows of X.T = its column block of X (the whole of X.T on one GPU)
            t = d["X5"].transpose((1, 0))
            return t[lo:lo + rows, 0:m5] if self.sharded else t

        if self.mode == "strong" and self.world > 1:
            dim0 = lambda: dist.reduce_sharded(d["X"], "sum", dims=(0,), output="sharded")  # noqa: E731
        elif self.world > 1:
            dim0 = lambda: dist.reduce_sharded(d["X"], "sum", dims=(0,))  # noqa: E731  (weak: replicated result, as in round 1)
        else:
            dim0 = lambda: d["X"].reduce_sum(dims=(0,))  # noqa: E731
        copy_t = (lambda: al.copy(xt(), deep=True)) if self.sharded else (lambda: xt().reshape((m5 * m5,)))
        return [
            ("c2_add", lambda: d["a"] + d["b"]),
            ("c2_mul", lambda: d["a"] * d["b"]),
            ("c2_add_scalar", lambda: d["a"] + 1.0),
            ("c3_bcast_add", c3_first),
            ("c3_bcast_add_vec", c3_second),
            ("c4_reduce_dims12", lambda: d["X"].reduce_sum(dims=(1, 2))),
            ("c4_reduce_dim3", lambda: d["X"].reduce_sum(dims=(3,))),
            ("c4_reduce_dim0", dim0),
            ("c5a_window_reduce", lambda: d["base"].as_strided(shape=(self.win_out, w), stride=(1, 1)).reduce_sum(dims=[1])),
            ("c5b_transposed_add", lambda: xt() + d["Y5"]),
            ("c5b_transposed_copy", copy_t),
        ]


def parity_check(al, dist, rank, world):
    """Before anything is timed at N > 1: the cross-GPU reduce (both variants), a window halo shard and a sharded
    broadcast chain on integer-valued data must equal the host results bit for bit on every rank."""
    import numpy as np
    rng = np.random.default_rng(1234)  # every rank builds the same global arrays
    G = rng.integers(-8, 9, size=(4 * world, 24, 16 * world, 8)).astype(np.float32)
    lo, hi = dist.shard_bounds(G.shape[0], rank, world)
    shard = al.from_numpy(G[lo:hi])
    want = G.sum(axis=0, keepdims=True, dtype=np.float64).astype(np.float32)
    ok = True
    for _ in range(2):  # twice: both slot buffers of the peer windows
        ok &= bool(np.array_equal(al.to_numpy(dist.reduce_sharded(shard, "sum", dims=(0,))), want))
        blk = al.to_numpy(dist.reduce_sharded(shard, "sum", dims=(0,), output="sharded"))
        rows = want.shape[1] // world
        ok &= bool(np.array_equal(blk, want[:, rank * rows:(rank + 1) * rows]))
    want02 = G.max(axis=(0, 2), keepdims=True)
    ok &= bool(np.array_equal(al.to_numpy(dist.reduce_sharded(shard, "max", dims=(0, 2))), want02))
    base = rng.integers(-8, 9, size=(100000,)).astype(np.float32)
    w = 64
    part, n_local = dist.window_shard_from_numpy(base, w)
    got = al.to_numpy(part.as_strided(shape=(n_local, w), stride=(1, 1)).reduce_sum(dims=[1]))
    full = np.lib.stride_tricks.sliding_window_view(base, w).sum(axis=1, dtype=np.float64).astype(np.float32)
    wlo, whi = dist.shard_bounds(base.size - w + 1, rank, world)
    ok &= bool(np.array_equal(got.reshape(-1), full[wlo:whi]))
    A = rng.integers(-8, 9, size=(4 * world, 1, 64)).astype(np.float32)
    B = rng.integers(-8, 9, size=(1, 12, 64)).astype(np.float32)
    got = al.to_numpy(al.from_numpy(A[4 * rank:4 * rank + 4]) + al.from_numpy(B) + 1.0)
    ok &= bool(np.array_equal(got, (A + B + 1.0)[4 * rank:4 * rank + 4]))
    return dist.allreduce_scalar(1.0 if ok else 0.0, "min") == 1.0


def run_gpu(args):
    import ctypes as C

    import numpy as np

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    os.environ.setdefault("ARRAYLIB_DEVICE", str(local_rank))
    import arraylib as al
    from arraylib_b200 import dist
    from arraylib_b200._lib import lib

    assert lib.al_init(local_rank) == 0, lib.al_last_error()
    if world > 1:
        dist.init(rank, world)  # NCCL id through the file rendezvous; no torch.distributed

    barrier = dist.barrier

    def max_over_ranks(x: float) -> float:
        return dist.allreduce_scalar(x, "max")

    def max_vector(values):
        if world == 1:
            return list(values)
        arr = al.from_numpy(np.asarray(list(values) + [0.0] * (-len(values) % 4), dtype=np.float32))
        dist.allreduce(arr, "max")
        return [float(v) for v in al.to_numpy(arr)[: len(values)]]

    parity = parity_check(al, dist, rank, world) if world > 1 else None
    if parity is False:
        raise SystemExit("cross-GPU parity check FAILED; nothing was timed")

    sh = suite_shapes(args.scale)
    nbytes = algorithmic_bytes(sh)

    def timed_run(mode, steps, warmup, sample_clocks):
        """Upload one suite, warm up, time exactly `steps` steps (CUDA events on the library stream, max over ranks)."""
        suite = Suite(al, dist, sh, rank, world, mode)
        host_inputs = suite.make_host_inputs()
        dev = {k: al.from_numpy(v, blocking=False) for k, v in host_inputs.items()}
        al.synchronize()
        ops = suite.make_ops(dev)
        kernel_of = {}

        def step(record_base=None):
            for i, (name, thunk) in enumerate(ops):
                if record_base is not None:
                    lib.al_event_record(record_base + 2 * i)
                out = thunk()
                if record_base is not None:
                    lib.al_event_record(record_base + 2 * i + 1)
                kernel_of.setdefault(name, lib.al_last_kernel().decode())
                if name != "c3_bcast_add":
                    del out

        for _ in range(max(warmup, 0)):
            step()
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0 and sample_clocks:
            sampler.start()
            time.sleep(0.15)
        barrier()
        launches_before = lib.al_launch_count()
        assert lib.al_timer_start() == 0
        for k in range(steps):
            step(record_base=2 * len(ops) * k)
        ms = C.c_float()
        assert lib.al_timer_stop(C.byref(ms)) == 0, lib.al_last_error()
        barrier()
        launches = int(lib.al_launch_count() - launches_before)
        clocks = sampler.stop() if rank == 0 and sample_clocks else None
        total_ms = max_over_ranks(float(ms.value))
        meds = []
        for i in range(len(ops)):
            ts = []
            for k in range(steps):
                e = C.c_float()
                b = 2 * len(ops) * k + 2 * i
                assert lib.al_event_elapsed(b, b + 1, C.byref(e)) == 0, lib.al_last_error()
                ts.append(float(e.value))
            meds.append(statistics.median(ts))
        meds = max_vector(meds)  # the slowest rank's median, per op
        return dict(suite=suite, dev=dev, ops=ops, ms_per_step=total_ms / steps, launches=launches, clocks=clocks,
                    per_op_ms=dict(zip([n for n, _ in ops], meds)), kernel_of=kernel_of)

    mode = args.scaling if world > 1 else "strong"
    run = timed_run(mode, args.steps, args.warmup, True)
    suite, dev = run["suite"], run["dev"]
    ms_per_step, launches, clocks = run["ms_per_step"], run["launches"], run["clocks"]
    step_bytes = suite.step_bytes
    value = step_bytes / (ms_per_step * 1e-3) / 1e9
    bytes_scale = world if mode == "weak" else 1  # per-op bytes of the whole job

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    per_op = {}
    for name in OP_NAMES:
        med = run["per_op_ms"][name]
        gbs = bytes_scale * nbytes[name] / (med * 1e-3) / 1e9
        per_op[name] = {"ms": round(med, 4), "gbs": round(gbs, 1), "kernel": run["kernel_of"].get(name, "?"),
                        "frac_of_peak_per_gpu": round(gbs / world / peak, 3), "pct_of_8TBs_per_gpu": round(100 * gbs / world / NOMINAL_GBS, 1)}

    # dominant kernel by time: ew_nd_kernel<VEC=4, UNROLL=4, BinF<AL_SUM>>, launched by c2_add
    # (contiguous) and c3_bcast_add_vec (row-broadcast); together ~45% of the step. Per launch = per rank.
    dom = ["c2_add", "c3_bcast_add_vec"]
    dom_ms = sum(per_op[n]["ms"] for n in dom) / len(dom)
    dom_bytes = sum(nbytes[n] for n in dom) / len(dom) / (world if mode == "strong" else 1)
    traffic = traffic_src = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and world == 1:
        tj = json.load(open(tpath))
        traffic, traffic_src = tj.get("ew_nd_vec4_add_bytes_per_launch"), tj.get("source")
    roofline = {"bound": "hbm", "kernel": "ew_nd_kernel<VEC=4, UNROLL=4, BinF<AL_SUM>> (launched by c2_add and c3_bcast_add_vec)",
                "achieved": round(dom_bytes / (dom_ms * 1e-3) / 1e9, 1), "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": round(dom_bytes / (dom_ms * 1e-3) / 1e9 / peak, 3), "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": int(dom_bytes), "avg_launch_ms": round(dom_ms, 4), "launches_per_step": len(dom)}

    def time_thunk(thunk, reps=7, warm=2):
        for _ in range(warm):
            r = thunk()
            del r
        ts = []
        for _ in range(reps):
            lib.al_event_record(60000)
            r = thunk()
            lib.al_event_record(60001)
            e = C.c_float()
            assert lib.al_event_elapsed(60000, 60001, C.byref(e)) == 0
            ts.append(float(e.value))
            del r
        return max_over_ranks(statistics.median(ts))

    # ---- the cross-GPU reduce taken apart (collective time reported separately) ----
    cross = None
    if world > 1:
        barrier()
        local_ms = time_thunk(lambda: dev["X"].reduce_sum(dims=(0,)))
        barrier()
        ar_ms = time_thunk(lambda: dist.reduce_sharded(dev["X"], "sum", dims=(0,)))
        barrier()
        rs_ms = time_thunk(lambda: dist.reduce_sharded(dev["X"], "sum", dims=(0,), output="sharded"))
        cross = {"slab_per_rank": list(suite.c4), "outputs": int(np.prod(suite.c4[1:])),
                 "local_reduce_only_ms": round(local_ms, 4), "reduce_scatter_ms": round(rs_ms, 4), "allreduce_ms": round(ar_ms, 4),
                 "collective_part_ms": {"reduce_scatter": round(rs_ms - local_ms, 4), "allreduce": round(ar_ms - local_ms, 4)},
                 "in_step": "reduce_scatter (result sharded along the leading non-reduced axis)" if mode == "strong" else "allreduce (result replicated)",
                 "note": "timed alone, back to back, max over ranks; local_reduce_only is the same slab reduced without any exchange"}

    # ---- fused chain (reported separately; NOT part of `value`, which is the eager suite) ----
    fused = None
    try:
        c3 = suite.c3
        fms = time_thunk(lambda: (al.lazy(dev["A"]) + dev["B"] + dev["c"]).eval(), reps=5)
        eager_bytes = (nbytes["c3_bcast_add"] + nbytes["c3_bcast_add_vec"]) * bytes_scale
        fused_bytes = (4 * c3[0] * c3[1] * c3[2] + 4 * (c3[0] * c3[2] + c3[1] * c3[2] + c3[2])) * world
        fused = {"op": "c3 chain A+B+c as ONE kernel (al.lazy(...).eval())", "ms": round(fms, 4), "kernel": lib.al_last_kernel().decode(),
                 "eager_ms": round(per_op["c3_bcast_add"]["ms"] + per_op["c3_bcast_add_vec"]["ms"], 4),
                 "gbs_vs_eager_algorithmic_bytes": round(eager_bytes / (fms * 1e-3) / 1e9, 1),
                 "gbs_vs_fused_algorithmic_bytes": round(fused_bytes / (fms * 1e-3) / 1e9, 1)}
    except Exception as exc:
        if world > 1:  # a rank that skips the rest of a section would leave the other ranks' collectives unmatched
            raise
        fused = {"error": repr(exc)}

    # ---- e2e: host buffers in, host buffers out, through the public API ----
    e2e = None
    if not args.no_e2e:
        host_inputs = suite.host_inputs
        make_ops = suite.make_ops
        run["ops"] = None
        dev.clear()
        out_elems = suite.out_elems()
        hout_n = min(max(out_elems.values()), 1 << 28)  # <= 1 GiB pinned staging; larger results leave in slabs
        hout = al.host_buffer((hout_n,))
        d2h_bytes = 4 * sum(out_elems.values())

        def download(out):
            """D2H of a result through the pinned staging buffer (leading-axis slabs when it is larger)."""
            shape = out.shape
            row = int(np.prod(shape[1:])) if len(shape) > 1 else 1
            rows_per = max(hout_n // row, 1)
            if shape[0] * row <= hout_n:
                al.to_numpy(out, out=hout[: shape[0] * row].reshape(shape), blocking=False)
                return
            for lo in range(0, shape[0], rows_per):
                hi = min(lo + rows_per, shape[0])
                view = out[tuple([slice(lo, hi)] + [slice(0, e) for e in shape[1:]])]
                al.to_numpy(view, out=hout[: (hi - lo) * row].reshape((hi - lo,) + tuple(shape[1:])), blocking=False)

        class JustInTime(dict):
            """Device inputs uploaded (asynchronously, from pinned memory) when an op first asks for them, so that the
            compute stream only ever waits for the operands it is about to use and the H2D of later operands runs
            under the D2H of earlier results (PCIe is full duplex)."""

            def __missing__(self, key):
                self[key] = al.from_numpy(host_inputs[key], blocking=False)
                return self[key]

        def e2e_order(ops):
            """The same eleven ops, broadcast adds first: they need 32 MB of inputs and produce 16 GiB of results, so the
            D2H stream (the bottleneck: 23.8 GB leave, 9.7 GB arrive) is busy from the first millisecond and every other
            operand travels underneath it. In suite order the downloads could not start before 2 GiB had been uploaded."""
            first = [o for o in ops if o[0].startswith("c3_")]
            return first + [o for o in ops if not o[0].startswith("c3_")]

        def e2e_step():
            # uploads (H2D stream), kernels (compute stream) and downloads (D2H stream) overlap;
            # the step ends when every result has landed in host memory
            d = JustInTime()
            for name, thunk in e2e_order(make_ops(d)):
                out = thunk()
                if name == "c3_bcast_add":
                    continue
                download(out)
                del out
            al.synchronize()

        e2e_steps = max(1, min(args.steps, 3))
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0) / e2e_steps
        e2e = {"value": round(step_bytes / e2e_s / 1e9, 2), "unit": "GB/s", "h2d_bytes_per_step": int(suite.h2d_bytes) * world,
               "d2h_bytes_per_step": int(d2h_bytes) * world, "steps": e2e_steps, "s_per_step": round(e2e_s, 4),
               "note": "pinned host buffers -> from_numpy(blocking=False) per op, just in time -> op -> to_numpy(out=pinned staging, blocking=False) -> synchronize at the end of the step; the eleven ops run broadcast adds first (their 16 GiB result keeps the D2H stream busy while the other operands upload); H2D, kernels and D2H on separate streams; every input is uploaded and every result downloaded once per step on the rank that owns it; same algorithmic bytes as `value`; byte counts are whole-job"}

        # the same step through the STOCK signatures from_numpy(x) / to_numpy(y): pageable numpy arrays in and out,
        # staged by the library through its pinned ring; downloads run on a second host thread so that both PCIe
        # directions stay busy (ctypes releases the GIL)
        try:
            from concurrent.futures import ThreadPoolExecutor
            pageable = {k: np.array(v, copy=True) for k, v in host_inputs.items()}
            outs_pageable = {k: np.empty(n, dtype=np.float32) for k, n in out_elems.items()}

            def stock_step(pool):
                d, pending = {}, []
                for name, thunk in e2e_order(make_ops(type("Lazy", (dict,), {"__missing__": lambda self, k: self.setdefault(k, al.from_numpy(pageable[k]))})())):
                    out = thunk()
                    if name == "c3_bcast_add":
                        continue
                    pending.append(pool.submit(lambda o=out, n=name: al.to_numpy(o, out=outs_pageable[n].reshape(o.shape))))
                    del out
                for f in pending:
                    f.result()
                al.synchronize()
                del d

            with ThreadPoolExecutor(max_workers=1) as pool:
                stock_step(pool)
                barrier()
                t0 = time.perf_counter()
                for _ in range(e2e_steps):
                    stock_step(pool)
                barrier()
                stock_s = max_over_ranks(time.perf_counter() - t0) / e2e_steps
            e2e["stock_api"] = {"value": round(step_bytes / stock_s / 1e9, 2), "unit": "GB/s", "s_per_step": round(stock_s, 4),
                                "fraction_of_pinned": round(e2e_s / stock_s, 3),
                                "note": "unmodified reference signatures from_numpy(x) / to_numpy(y) on pageable numpy arrays (to_numpy given a pageable `out` to avoid 24 GB of fresh allocations per step); downloads on one extra host thread"}
            del pageable, outs_pageable
        except Exception as exc:  # noqa: BLE001
            if world > 1:
                raise
            e2e["stock_api"] = {"error": repr(exc)}
        # what the host side of this box can move while ALL ranks copy at once (the ceiling of the e2e figure):
        # 1 GiB pinned each way per rank, one direction at a time and then both together
        try:
            del hout
            pn = 1 << 28
            ph_in, ph_out = al.host_buffer((pn,)), al.host_buffer((pn,))
            pd = al.from_numpy(ph_in, blocking=False)
            al.synchronize()

            def copy_rate(up, down, reps=3):
                barrier()
                t0 = time.perf_counter()
                for _ in range(reps):
                    if up:
                        keep = al.from_numpy(ph_in, blocking=False)
                    if down:
                        al.to_numpy(pd, out=ph_out, blocking=False)
                    al.synchronize()
                    if up:
                        del keep
                barrier()
                dt = max_over_ranks(time.perf_counter() - t0) / reps
                return round(4 * pn * world * (int(up) + int(down)) / dt / 1e9, 1)

            copy_rate(True, True, 1)
            e2e["pcie_probe"] = {"h2d_gbs": copy_rate(True, False), "d2h_gbs": copy_rate(False, True), "both_gbs": copy_rate(True, True),
                                 "wire_gbs_of_the_e2e_step": round((int(suite.h2d_bytes) + int(d2h_bytes)) * world / e2e_s / 1e9, 1),
                                 "note": "aggregate over all ranks copying AT THE SAME TIME, 1 GiB pinned per rank and direction; the e2e step cannot move its bytes faster than this"}
            del pd, ph_in, ph_out
        except Exception as exc:  # noqa: BLE001
            if world > 1:
                raise
            e2e["pcie_probe"] = {"error": repr(exc)}

    # ---- the other scaling mode, briefly (N > 1 only) ----
    other = None
    if world > 1 and args.other_steps > 0:
        del run, dev
        lib.al_trim_pool()
        omode = "weak" if mode == "strong" else "strong"
        o = timed_run(omode, args.other_steps, 2, False)
        oval = o["suite"].step_bytes / (o["ms_per_step"] * 1e-3) / 1e9
        other = {"scaling": omode, "value": round(oval, 1), "unit": "GB/s", "ms_per_step": round(o["ms_per_step"], 4), "steps": args.other_steps,
                 "pct_of_8TBs_per_gpu": round(100 * oval / world / NOMINAL_GBS, 1),
                 "c4_reduce_dim0_ms": round(o["per_op_ms"]["c4_reduce_dim0"], 4),
                 "note": "every rank owns a full copy of the named shapes (global leading extent = N x the named one)" if omode == "weak" else "BASELINE shapes cut over the ranks"}
        del o

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cpu_baseline = reference_sample(steps=1, warmup=0, budget_s=40.0)["cpu_baseline"]
        except Exception as exc:  # the checker is optional for the GPU number
            cpu_baseline = {"error": repr(exc)}

    transport = dist.transport()
    if world > 1:
        dist.destroy()
    if rank != 0:
        return None
    n2, c3, c4, n5, m5, w = sh["n2"], sh["c3"], sh["c4"], sh["n5"], sh["m5"], sh["window"]
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": mode if world > 1 else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (standard-normal float32, random-init, no datasets)",
        "config": {"workload": "C2+C3+C4+C5 hot-path suite at the BASELINE shapes" + ("" if args.scale == 1 else f" / scale {args.scale}"),
                   "shapes": {"c2_n": n2, "c3": f"[{c3[0]},1,{c3[2]}]+[1,{c3[1]},{c3[2]}]+[{c3[2]}]", "c4": list(c4), "c5a_n": n5, "c5a_window": w, "c5b": [m5, m5]},
                   "sharding": ("single GPU" if world == 1 else
                                "strong: the BASELINE shapes cut along their leading axis over the ranks (C3: B and c replicated; C5a: 63-element halo; C5b: X replicated, each rank reads its column block; c4_reduce_dim0: peer-memory reduce-scatter, result sharded along the leading non-reduced axis)"
                                if mode == "strong" else "weak: one full copy of the named shapes per GPU"),
                   "cross_gpu_reduce": transport,
                   "l2_policy": "every operand is >= 128 MiB (larger than the 126 MB L2) and ops alternate operands; no flush",
                   "algorithmic_bytes_per_step": int(step_bytes), "process_group": "file rendezvous + NCCL id, no torch.distributed"},
        "pct_of_8TBs_per_gpu": round(100 * value / world / NOMINAL_GBS, 1),
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "ops": per_op, "cross_gpu_reduce": cross,
        "fused_chain": fused, "e2e": e2e, "cpu_baseline": cpu_baseline,
    }
    if world > 1:
        line["parity_checked"] = bool(parity)
        line[other["scaling"] if other else "other_scaling"] = other
    return line


# ---------------------------------------------------------------------------------------------------
# reference arm: the unmodified reference on the host cores, bounded sample of the same suite
# ---------------------------------------------------------------------------------------------------
# Sample tiers, largest first. "section 4.5" is what BASELINE.md section 4.5 prescribes for the CPU side: C2 and C5b at
# full size, C3 [256,1,4096]+[1,256,4096], C4 [16,256,256,64], C5a N=2^24 (the reference needs minutes and tens of GiB
# beyond that: C3 materialises 2^32-element temporaries, C5a allocates its whole output once per OpenMP thread,
# arraylib.c:732). The smaller tiers exist because `--steps K --warmup W` must finish within a few minutes.
CPU_TIERS = [
    ("section 4.5", dict(n2=1 << 28, c3=(256, 256, 4096), c4=(16, 256, 256, 64), n5=1 << 24, m5=16384, window=64)),
    ("section 4.5 / 4", dict(n2=1 << 26, c3=(128, 128, 4096), c4=(8, 128, 256, 64), n5=1 << 22, m5=8192, window=64)),
    ("section 4.5 / 16", dict(n2=1 << 24, c3=(64, 64, 4096), c4=(8, 64, 128, 64), n5=1 << 20, m5=4096, window=64)),
    ("round-1 sample", dict(n2=1 << 22, c3=(128, 128, 2048), c4=(8, 128, 128, 64), n5=1 << 20, m5=2048, window=64)),
]


def describe_sample(name, sh):
    full = algorithmic_bytes(suite_shapes(1))
    mine = algorithmic_bytes(sh)
    c3, c4 = sh["c3"], sh["c4"]
    text = (f"{name}: C2 n=2^{sh['n2'].bit_length() - 1}; C3 [{c3[0]},1,{c3[2]}]+[1,{c3[1]},{c3[2]}]+[{c3[2]}]; C4 {list(c4)}; "
            f"C5a n=2^{sh['n5'].bit_length() - 1} window {sh['window']}; C5b {sh['m5']}x{sh['m5']} (same ops and byte formulas as the GPU suite)")
    factors = {k: round(full[k] / mine[k], 2) for k in mine}
    return text, factors


def make_cpu_ops(sh):
    """The suite on the host: (kind, ops). The unmodified reference when oracle/_ref holds it, else the oracle port."""
    import numpy as np
    from oracle import ref_binding

    n2, c3, c4, n5, m5, w = sh["n2"], sh["c3"], sh["c4"], sh["n5"], sh["m5"], sh["window"]
    block = np.random.default_rng(0).standard_normal(1 << 22, dtype=np.float32)

    def rnd(*shape):  # tiled random block: the values do not matter for the timing, only that they are not constant
        return np.resize(block, int(np.prod(shape))).reshape(shape)

    keep = {}
    if ref_binding.available(omp=True):
        R = ref_binding.RefLib(omp=True)
        up = R.from_numpy
        a, b = up(rnd(n2)), up(rnd(n2))
        A, B, c = up(rnd(c3[0], 1, c3[2])), up(rnd(1, c3[1], c3[2])), up(rnd(c3[2]))
        X, base, X5, Y5 = up(rnd(*c4)), up(rnd(n5)), up(rnd(m5, m5)), up(rnd(m5, m5))
        return "reference", [
            ("c2_add", lambda: R.binary("sum", a, b)), ("c2_mul", lambda: R.binary("mul", a, b)),
            ("c2_add_scalar", lambda: R.binary("sum", a, 1.0)),
            ("c3_bcast_add", lambda: keep.__setitem__("t", R.binary("sum", A, B))),
            ("c3_bcast_add_vec", lambda: R.binary("sum", keep.pop("t"), c)),
            ("c4_reduce_dims12", lambda: R.reduce("sum", X, (1, 2))), ("c4_reduce_dim3", lambda: R.reduce("sum", X, (3,))),
            ("c4_reduce_dim0", lambda: R.reduce("sum", X, (0,))),
            ("c5a_window_reduce", lambda: R.reduce("sum", R.as_strided(base, (n5 - w + 1, w), (1, 1)), (1,))),
            ("c5b_transposed_add", lambda: R.binary("sum", R.transpose(X5, (1, 0)), Y5)),
            ("c5b_transposed_copy", lambda: R.reshape(R.transpose(X5, (1, 0)), (m5 * m5,))),
        ]
    from oracle import oracle as orc  # our C restatement; OpenMP on the elementwise loops only
    orc.build()
    a, b = rnd(n2), rnd(n2)
    A, B, c = rnd(c3[0], 1, c3[2]), rnd(1, c3[1], c3[2]), rnd(c3[2])
    X, base, X5, Y5 = rnd(*c4), rnd(n5), rnd(m5, m5), rnd(m5, m5)
    win = np.lib.stride_tricks.as_strided(base, shape=(n5 - w + 1, w), strides=(4, 4))
    return "port", [
        ("c2_add", lambda: orc.binary("sum", a, b)), ("c2_mul", lambda: orc.binary("mul", a, b)),
        ("c2_add_scalar", lambda: orc.binary("sum", a, 1.0)),
        ("c3_bcast_add", lambda: keep.__setitem__("t", orc.binary("sum", A, B))),
        ("c3_bcast_add_vec", lambda: orc.binary("sum", keep.pop("t"), c)),
        ("c4_reduce_dims12", lambda: orc.reduce("sum", X, (1, 2))), ("c4_reduce_dim3", lambda: orc.reduce("sum", X, (3,))),
        ("c4_reduce_dim0", lambda: orc.reduce("sum", X, (0,))),
        ("c5a_window_reduce", lambda: orc.reduce("sum", win, (1,))),
        ("c5b_transposed_add", lambda: orc.binary("sum", X5.T, Y5)),
        ("c5b_transposed_copy", lambda: orc.gather(X5.T)),
    ]


def run_cpu_ops(ops, steps, warmup):
    per_op = {n: [] for n, _ in ops}

    def step(record):
        for name, thunk in ops:
            t0 = time.perf_counter()
            r = thunk()
            dt = time.perf_counter() - t0
            del r
            if record:
                per_op[name].append(dt)

    for _ in range(max(warmup, 0)):
        step(False)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(True)
    return (time.perf_counter() - t0) / steps, per_op


def reference_sample(steps: int, warmup: int, budget_s: float = 240.0, tier: str | None = None):
    """Times `steps` steps (+ `warmup`) of the suite on the host cores, on the largest sample tier whose estimated
    run time fits `budget_s`. The estimate comes from one calibration step two tiers down (host loops here are
    index-arithmetic bound, so time scales with the element count)."""
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(cores)
    chosen = None
    calibration = None
    if tier is not None:
        chosen = next(i for i, (n, _) in enumerate(CPU_TIERS) if n == tier)
    else:
        cal_name, cal_sh = CPU_TIERS[2]
        _, cal_ops = make_cpu_ops(cal_sh)
        run_cpu_ops(cal_ops, 1, 0)  # first touch
        _, cal_times = run_cpu_ops(cal_ops, 1, 0)
        cal_bytes = algorithmic_bytes(cal_sh)
        del cal_ops
        calibration = {}
        for i, (name, sh) in enumerate(CPU_TIERS):
            b = algorithmic_bytes(sh)
            est = sum(cal_times[k][0] * b[k] / cal_bytes[k] for k in b) * (1.5 if i < 2 else 1.0)  # larger tiers fall out of cache / TLB reach
            calibration[name] = round(est, 2)
            if chosen is None and est * (steps + max(warmup, 0)) <= budget_s:
                chosen = i
        if chosen is None:
            chosen = len(CPU_TIERS) - 1
    name, sh = CPU_TIERS[chosen]
    nbytes = algorithmic_bytes(sh)
    kind, ops = make_cpu_ops(sh)
    s_per_step, per_op = run_cpu_ops(ops, steps, warmup)
    step_bytes = sum(nbytes.values())
    value = step_bytes / s_per_step / 1e9
    text, factors = describe_sample(name, sh)
    ops_out = {n: {"ms": round(1e3 * statistics.median(v), 3), "gbs": round(nbytes[n] / statistics.median(v) / 1e9, 3),
                   "baseline_bytes_over_sample_bytes": factors[n]} for n, v in per_op.items()}
    base_obj = {"value": round(value, 3), "unit": "GB/s", "cores": cores, "kind": kind, "sample": text, "tier": name,
                "estimated_s_per_step_by_tier": calibration, "time_budget_s": budget_s,
                "build": "oracle/_ref/arraylib_ref_omp.so: unmodified arraylib.c, reference Makefile flags + -fopenmp, -march=x86-64-v3 instead of -march=native (built off-box)",
                "ops": ops_out}
    return {"value": value, "ms_per_step": 1e3 * s_per_step, "cpu_baseline": base_obj, "step_bytes": step_bytes, "sample": text}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    r = reference_sample(args.steps, args.warmup, budget_s=args.cpu_budget, tier=args.cpu_tier)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(r["value"], 3), "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(r["ms_per_step"], 3), "higher_is_better": True,
        "scaling": "weak" if args.gpus == 1 else args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (standard-normal float32)",
        "config": {"workload": "C2+C3+C4+C5 hot-path suite, " + r["sample"], "algorithmic_bytes_per_step": int(r["step_bytes"])},
        "cpu_baseline": r["cpu_baseline"],
        "e2e": {"value": round(r["value"], 3), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    return line


class StdoutGuard:
    """NCCL (and anything else) may print to fd 1; the contract is ONE JSON line on stdout. Route
    fd 1 to stderr for the duration of the run and keep the real stdout for the final line."""

    def __enter__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text: str):
        sys.stdout.flush()
        os.write(self.real, (text + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.real, 1)
        os.close(self.real)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=int, default=1, help="shrink the suite (development only; 1 = BASELINE shapes)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: cut the BASELINE shapes over the ranks (strong, default) or give every rank a full copy (weak)")
    ap.add_argument("--other-steps", type=int, default=3, help="N > 1: steps of the OTHER scaling mode reported next to the main one (0 = skip)")
    ap.add_argument("--cpu-budget", type=float, default=240.0, help="--impl reference: seconds the K+W steps may take (picks the sample tier)")
    ap.add_argument("--cpu-tier", default=None, help="--impl reference: force a sample tier by name")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    with StdoutGuard() as out:
        line = run_reference(args) if args.impl == "reference" else run_gpu(args)
        if line is not None:
            out.emit(json.dumps(line))


if __name__ == "__main__":
    main()
