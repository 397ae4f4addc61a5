#!/usr/bin/env python
"""bench.py -- effective HBM GB/s of the array-core hot path on N B200s (BASELINE.json metric).

One STEP = one pass of the hot-path suite over synthetic float32 arrays of the BASELINE shapes:
  C2  a+b, a*b, a+1.0                      two 2^28-element contiguous arrays
  C3  t = A[1024,1,4096] + B[1,1024,4096]; u = t + c[4096]      (2^32-element outputs, eager)
  C4  X[64,512,512,64]: reduce_sum(dims=(1,2)), (3,), (0,)     ((0,) + NCCL allreduce when N>1)
  C5  window-64 as_strided view of 2^28 elements reduce_sum(dims=[1]);  X[16384,16384].T + Y;
      X.T.reshape(-1) (forces the strided-copy kernel)
`value` = algorithmic bytes of all ops (SURVEY.md section 8d / BASELINE.md section 5) summed over
all ranks / device time of the slowest rank (CUDA events on the library's stream). Multi-GPU runs
are WEAK scaling: every rank owns one leading-axis shard of the named shape (global leading
extent = N x the named one), so elementwise/broadcast/copy ops need no communication and only
reduce_sum(dims=(0,)) runs a collective.

`--impl reference` times the UNMODIFIED reference (oracle/_ref/arraylib_ref_omp.so, built with
the reference Makefile's flags + -fopenmp) on the host cores, on a bounded sample of the same
suite. The same sample is reported as `cpu_baseline` in the default arm.
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


# ---------------------------------------------------------------------------------------------------
# workload definition (shared by the GPU arm, the e2e leg and the CPU reference arm)
# ---------------------------------------------------------------------------------------------------
def suite_shapes(scale: int):
    """Shapes of the suite; scale=1 is the BASELINE configuration, larger values shrink it."""
    s = max(scale, 1)
    n2 = (1 << 28) // s
    l3 = max(1024 // s, 2)
    l4 = max(64 // s, 2)
    n5 = (1 << 28) // s
    m5 = 16384 if s == 1 else max(int(16384 / (s ** 0.5)) // 32 * 32, 64)
    return dict(n2=n2, c3=(l3, 1024 if s == 1 else max(1024 // s, 2), 4096), c4=(l4, 512, 512, 64) if s == 1 else (l4, max(512 // s, 4), 512, 64),
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
def run_gpu(args):
    import numpy as np

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    os.environ.setdefault("ARRAYLIB_DEVICE", str(local_rank))
    import arraylib as al
    from arraylib_b200 import dist
    from arraylib_b200._lib import lib
    import ctypes as C

    assert lib.al_init(local_rank) == 0, lib.al_last_error()

    tdist = None
    if world > 1:
        import torch
        import torch.distributed as tdist
        torch.cuda.set_device(local_rank)
        tdist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

        def bcast(obj, src):
            box = [obj]
            tdist.broadcast_object_list(box, src=src)
            return box[0]

        dist.init(rank, world, broadcast=bcast)

    def barrier():
        if tdist is not None:
            import torch
            tdist.barrier()
            torch.cuda.synchronize()
        al.synchronize()

    def max_over_ranks(x: float) -> float:
        if tdist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        return float(t.item())

    sh = suite_shapes(args.scale)
    nbytes = algorithmic_bytes(sh)
    n2, c3, c4, n5, m5, w = sh["n2"], sh["c3"], sh["c4"], sh["n5"], sh["m5"], sh["window"]
    x4 = c4[0] * c4[1] * c4[2] * c4[3]

    # ---- synthetic inputs: one pinned host pool of random normals, sliced for every operand ----
    pool_n = max(x4, 2 * n2, 2 * m5 * m5)
    hin = al.host_buffer((pool_n,))
    rng = np.random.default_rng(rank)
    block = rng.standard_normal(min(pool_n, 1 << 24), dtype=np.float32)
    for off in range(0, pool_n, block.size):
        hin[off:off + block.size] = block[: min(block.size, pool_n - off)]

    def host_in(offset, shape):
        n = int(np.prod(shape))
        return hin[offset:offset + n].reshape(shape)

    host_inputs = {
        "a": host_in(0, (n2,)), "b": host_in(n2, (n2,)),
        "A": host_in(0, (c3[0], 1, c3[2])), "B": host_in(c3[0] * c3[2], (1, c3[1], c3[2])), "c": host_in(7, (c3[2],)),
        "X": host_in(0, c4), "base": host_in(0, (n5,)), "X5": host_in(0, (m5, m5)), "Y5": host_in(m5 * m5, (m5, m5)),
    }
    h2d_bytes = sum(v.nbytes for v in host_inputs.values())

    def upload(blocking=True):
        return {k: al.from_numpy(v, blocking=blocking) for k, v in host_inputs.items()}

    def make_ops(d):
        """name -> thunk returning the op's result (inputs are device arrays in `d`)."""
        keep = {}

        def c3_first():
            keep["t"] = d["A"] + d["B"]
            return keep["t"]

        def c3_second():
            return keep.pop("t") + d["c"]

        return [
            ("c2_add", lambda: d["a"] + d["b"]),
            ("c2_mul", lambda: d["a"] * d["b"]),
            ("c2_add_scalar", lambda: d["a"] + 1.0),
            ("c3_bcast_add", c3_first),
            ("c3_bcast_add_vec", c3_second),
            ("c4_reduce_dims12", lambda: d["X"].reduce_sum(dims=(1, 2))),
            ("c4_reduce_dim3", lambda: d["X"].reduce_sum(dims=(3,))),
            ("c4_reduce_dim0", lambda: dist.reduce_sharded(d["X"], "sum", dims=(0,))),
            ("c5a_window_reduce", lambda: d["base"].as_strided(shape=(n5 - w + 1, w), stride=(1, 1)).reduce_sum(dims=[1])),
            ("c5b_transposed_add", lambda: d["X5"].transpose((1, 0)) + d["Y5"]),
            ("c5b_transposed_copy", lambda: d["X5"].transpose((1, 0)).reshape((m5 * m5,))),
        ]

    dev = upload()
    ops = make_ops(dev)
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

    # ---- warm-up ----
    for _ in range(max(args.warmup, 0)):
        step()
    barrier()

    # ---- timed region: exactly K steps, CUDA events on the library stream ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    barrier()
    launches_before = lib.al_launch_count()
    assert lib.al_timer_start() == 0
    for k in range(args.steps):
        step(record_base=2 * len(ops) * k)
    ms = C.c_float()
    assert lib.al_timer_stop(C.byref(ms)) == 0, lib.al_last_error()
    barrier()
    launches = int(lib.al_launch_count() - launches_before)
    clocks = sampler.stop() if rank == 0 else None
    total_ms = max_over_ranks(float(ms.value))
    ms_per_step = total_ms / args.steps
    step_bytes = sum(nbytes.values())
    value = world * step_bytes / (ms_per_step * 1e-3) / 1e9

    # per-op device times from the events recorded inside the timed region
    per_op = {}
    for i, (name, _) in enumerate(ops):
        ts = []
        for k in range(args.steps):
            e = C.c_float()
            b = 2 * len(ops) * k + 2 * i
            assert lib.al_event_elapsed(b, b + 1, C.byref(e)) == 0, lib.al_last_error()
            ts.append(float(e.value))
        med = statistics.median(ts)
        per_op[name] = {"ms": round(med, 4), "gbs": round(nbytes[name] / (med * 1e-3) / 1e9, 1),
                        "kernel": kernel_of.get(name, "?")}

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    for v in per_op.values():
        v["frac_of_peak"] = round(v["gbs"] / peak, 3)
        v["pct_of_8TBs"] = round(100 * v["gbs"] / NOMINAL_GBS, 1)

    # dominant kernel by time: ew_nd_kernel<VEC=4, UNROLL=4, BinF<AL_SUM>>, launched by c2_add
    # (contiguous) and c3_bcast_add_vec (row-broadcast); together ~45% of the step.
    dom = ["c2_add", "c3_bcast_add_vec"]
    dom_ms = sum(per_op[n]["ms"] for n in dom) / len(dom)
    dom_bytes = sum(nbytes[n] for n in dom) / len(dom)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("ew_nd_vec4_add_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "ew_nd_kernel<VEC=4, UNROLL=4, BinF<AL_SUM>> (launched by c2_add and c3_bcast_add_vec)",
                "achieved": round(dom_bytes / (dom_ms * 1e-3) / 1e9, 1), "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": round(dom_bytes / (dom_ms * 1e-3) / 1e9 / peak, 3), "traffic": traffic,
                "algorithmic_bytes_per_launch": int(dom_bytes), "avg_launch_ms": round(dom_ms, 4), "launches_per_step": len(dom)}

    # ---- fused chain (reported separately; NOT part of `value`, which is the eager suite) ----
    fused = None
    try:
        def chain():
            return (al.lazy(dev["A"]) + dev["B"] + dev["c"]).eval()
        for _ in range(2):
            r = chain()
            del r
        ts = []
        for _ in range(5):
            lib.al_event_record(60000)
            r = chain()
            lib.al_event_record(60001)
            e = C.c_float()
            assert lib.al_event_elapsed(60000, 60001, C.byref(e)) == 0
            ts.append(float(e.value))
            del r
        fms = statistics.median(ts)
        eager_bytes = nbytes["c3_bcast_add"] + nbytes["c3_bcast_add_vec"]
        fused_bytes = 4 * c3[0] * c3[1] * c3[2] + 4 * (c3[0] * c3[2] + c3[1] * c3[2] + c3[2])
        fused = {"op": "c3 chain A+B+c as ONE kernel (al.lazy(...).eval())", "ms": round(fms, 4), "kernel": lib.al_last_kernel().decode(),
                 "eager_ms": round(per_op["c3_bcast_add"]["ms"] + per_op["c3_bcast_add_vec"]["ms"], 4),
                 "gbs_vs_eager_algorithmic_bytes": round(eager_bytes / (fms * 1e-3) / 1e9, 1),
                 "gbs_vs_fused_algorithmic_bytes": round(fused_bytes / (fms * 1e-3) / 1e9, 1)}
    except Exception as exc:
        fused = {"error": repr(exc)}

    # ---- e2e: host buffers in, host buffers out, through the public API ----
    e2e = None
    if not args.no_e2e:
        del ops
        dev.clear()
        del dev
        out_elems = {"c2_add": n2, "c2_mul": n2, "c2_add_scalar": n2, "c3_bcast_add_vec": c3[0] * c3[1] * c3[2],
                     "c4_reduce_dims12": c4[0] * c4[3], "c4_reduce_dim3": x4 // c4[3], "c4_reduce_dim0": x4 // c4[0],
                     "c5a_window_reduce": n5 - w + 1, "c5b_transposed_add": m5 * m5, "c5b_transposed_copy": m5 * m5}
        hout_n = min(max(out_elems.values()), 1 << 28)  # 1 GiB pinned staging; larger results leave in slabs
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

        def e2e_step():
            # uploads (H2D stream), kernels (compute stream) and downloads (D2H stream) overlap;
            # the step ends when every result has landed in host memory
            d = JustInTime()
            for name, thunk in make_ops(d):
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
        e2e = {"value": round(world * step_bytes / e2e_s / 1e9, 2), "unit": "GB/s", "h2d_bytes_per_step": int(h2d_bytes),
               "d2h_bytes_per_step": int(d2h_bytes), "steps": e2e_steps, "s_per_step": round(e2e_s, 4),
               "note": "pinned host buffers -> from_numpy(blocking=False) per op, just in time -> op -> to_numpy(out=pinned 1 GiB staging, blocking=False) -> synchronize at the end of the step; H2D, kernels and D2H on separate streams; every input is uploaded and every result downloaded once per step; same algorithmic bytes as `value`"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cpu_baseline = reference_sample(steps=1, warmup=1)["cpu_baseline"]
        except Exception as exc:  # the checker is optional for the GPU number
            cpu_baseline = {"error": repr(exc)}

    transport = dist.transport()
    if tdist is not None:
        dist.destroy()
        tdist.destroy_process_group()
    if rank != 0:
        return None
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (standard-normal float32, random-init, no datasets)",
        "config": {"workload": "C2+C3+C4+C5 hot-path suite at the BASELINE shapes" + ("" if args.scale == 1 else f" / scale {args.scale}"),
                   "shapes": {"c2_n": n2, "c3": f"[{c3[0]},1,{c3[2]}]+[1,{c3[1]},{c3[2]}]+[{c3[2]}]", "c4": list(c4), "c5a_n": n5, "c5a_window": w, "c5b": [m5, m5]},
                   "sharding": "leading axis, one shard of the named shape per GPU",
                   "cross_gpu_reduce": transport, "l2_policy": "inputs larger than L2 (>=1 GiB per operand); no flush",
                   "algorithmic_bytes_per_step_per_gpu": int(step_bytes)},
        "pct_of_8TBs_per_gpu": round(100 * value / world / NOMINAL_GBS, 1),
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "ops": per_op, "fused_chain": fused, "e2e": e2e, "cpu_baseline": cpu_baseline,
    }
    return line


# ---------------------------------------------------------------------------------------------------
# reference arm: the unmodified reference on the host cores, bounded sample of the same suite
# ---------------------------------------------------------------------------------------------------
CPU_SCALE_NOTE = ("bounded sample: C2 n=2^22; C3 [128,1,2048]+[1,128,2048]+[2048]; C4 [8,128,128,64]; "
                  "C5a n=2^20 window 64; C5b 2048x2048 (same ops and byte formulas as the GPU suite)")


def cpu_sample_shapes():
    return dict(n2=1 << 22, c3=(128, 128, 2048), c4=(8, 128, 128, 64), n5=1 << 20, m5=2048, window=64)


def reference_sample(steps: int, warmup: int):
    import numpy as np

    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(cores)
    from oracle import ref_binding

    sh = cpu_sample_shapes()
    nbytes = algorithmic_bytes(sh)
    n2, c3, c4, n5, m5, w = sh["n2"], sh["c3"], sh["c4"], sh["n5"], sh["m5"], sh["window"]
    rng = np.random.default_rng(0)
    rnd = lambda *s: rng.standard_normal(s, dtype=np.float32)  # noqa: E731
    if ref_binding.available(omp=True):
        R = ref_binding.RefLib(omp=True)
        kind = "reference"
        up = R.from_numpy
        a, b = up(rnd(n2)), up(rnd(n2))
        A, B, c = up(rnd(c3[0], 1, c3[2])), up(rnd(1, c3[1], c3[2])), up(rnd(c3[2]))
        X, base, X5, Y5 = up(rnd(*c4)), up(rnd(n5)), up(rnd(m5, m5)), up(rnd(m5, m5))
        keep = {}
        ops = [
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
    else:  # the oracle port (our C restatement; OpenMP on the elementwise loops only)
        from oracle import oracle as orc
        orc.build()
        kind = "port"
        a, b = rnd(n2), rnd(n2)
        A, B, c = rnd(c3[0], 1, c3[2]), rnd(1, c3[1], c3[2]), rnd(c3[2])
        X, base, X5, Y5 = rnd(*c4), rnd(n5), rnd(m5, m5), rnd(m5, m5)
        win = np.lib.stride_tricks.as_strided(base, shape=(n5 - w + 1, w), strides=(4, 4))
        keep = {}
        ops = [
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
    total = time.perf_counter() - t0
    s_per_step = total / steps
    step_bytes = sum(nbytes.values())
    value = step_bytes / s_per_step / 1e9
    ops_out = {n: {"ms": round(1e3 * statistics.median(v), 3), "gbs": round(nbytes[n] / statistics.median(v) / 1e9, 3)} for n, v in per_op.items()}
    base_obj = {"value": round(value, 3), "unit": "GB/s", "cores": cores, "kind": kind, "sample": CPU_SCALE_NOTE, "ops": ops_out}
    return {"value": value, "ms_per_step": 1e3 * s_per_step, "cpu_baseline": base_obj, "step_bytes": step_bytes}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    r = reference_sample(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(r["value"], 3), "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(r["ms_per_step"], 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (standard-normal float32)",
        "config": {"workload": "C2+C3+C4+C5 hot-path suite, " + CPU_SCALE_NOTE, "algorithmic_bytes_per_step": int(r["step_bytes"])},
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
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    with StdoutGuard() as out:
        line = run_reference(args) if args.impl == "reference" else run_gpu(args)
        if line is not None:
            out.emit(json.dumps(line))


if __name__ == "__main__":
    main()
