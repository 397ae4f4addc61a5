"""Development microbenchmarks (not part of the product or the bench contract): times single ops
with CUDA events on the library stream under different tuning knobs. Usage under gpurun:
    python tools/microbench.py [fill|ew|bcast|transpose|window|reduce ...]
"""
import ctypes as C
import statistics
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import arraylib as al  # noqa: E402
from arraylib_b200._lib import lib  # noqa: E402


def tune(**kw):
    for k, v in kw.items():
        assert lib.al_set_tuning(k.encode(), int(v)) == 0, lib.al_last_error()


def timed(fn, nbytes, label, reps=7, warm=2):
    for _ in range(warm):
        r = fn()
        del r
    ts = []
    for i in range(reps):
        lib.al_event_record(0)
        r = fn()
        lib.al_event_record(1)
        ms = C.c_float()
        assert lib.al_event_elapsed(0, 1, C.byref(ms)) == 0
        ts.append(ms.value)
        del r
    med = statistics.median(ts)
    print(f"{label:58s} {med:8.4f} ms  {nbytes / med / 1e6:8.1f} GB/s  (min {min(ts):.4f})  [{lib.al_last_kernel().decode()}]", flush=True)
    return med


def rnd(shape, seed=0):
    n = int(np.prod(shape))
    blk = np.random.default_rng(seed).standard_normal(min(n, 1 << 22), dtype=np.float32)
    return al.from_numpy(np.resize(blk, n).reshape(shape))


def bench_fill():
    for gib in (1, 4):
        n = gib << 28
        timed(lambda: al.full((n,), 1.5), 4 * n, f"fill_const {gib} GiB (write only)")


def bench_ew():
    n = 1 << 28
    a, b = rnd((n,), 1), rnd((n,), 2)
    for unroll in (1, 2, 4, 8):
        for pol in (0, 1, 2):
            for ss in (0, 1 << 62):
                tune(ew_unroll=unroll, load_policy=pol, stream_store_min_bytes=ss)
                timed(lambda: a + b, 12 * n, f"a+b 2^28 unroll={unroll} load_policy={pol} stream_store={'on' if ss == 0 else 'off'}", reps=5)
    tune(ew_unroll=4, load_policy=0, stream_store_min_bytes=64 << 20)
    tune(stream_tma=1)
    timed(lambda: a + b, 12 * n, "a+b 2^28 cp.async.bulk + mbarrier pipeline (experiment)")
    tune(stream_tma=0)
    timed(lambda: a + 1.0, 8 * n, "a+1.0 2^28")
    timed(lambda: al.copy(a, deep=True), 8 * n, "deep copy 2^28 (contiguous)")


def bench_bcast():
    A, B, c = rnd((1024, 1, 4096), 1), rnd((1, 1024, 4096), 2), rnd((4096,), 3)
    out = 1 << 32
    for unroll in (2, 4, 8):
        for ss in (0, 1 << 62):
            tune(ew_unroll=unroll, stream_store_min_bytes=ss)
            timed(lambda: A + B, 4 * out + 8 * (1 << 22), f"[1024,1,4096]+[1,1024,4096] unroll={unroll} stream_store={'on' if ss == 0 else 'off'}", reps=5)
    tune(ew_unroll=4, stream_store_min_bytes=64 << 20)
    t = A + B
    timed(lambda: t + c, 8 * out, "t[1024,1024,4096] + c[4096]", reps=5)


def bench_transpose():
    m = 16384
    X, Y = rnd((m, m), 1), rnd((m, m), 2)
    for tq in (32, 64, 128):
        tune(tile_tq=tq)
        timed(lambda: X.transpose((1, 0)) + Y, 12 * m * m, f"X.T + Y 16384^2 tile_tq={tq}")
        timed(lambda: X.transpose((1, 0)).reshape((m * m,)), 8 * m * m, f"X.T.reshape (strided copy) 16384^2 tile_tq={tq}")
    tune(tile_tq=64)
    m2 = 16383  # odd extents: no 128-bit accesses possible -> scalar tile
    X2, Y2 = rnd((m2, m2), 7), rnd((m2, m2), 8)
    timed(lambda: X2.transpose((1, 0)) + Y2, 12 * m2 * m2, "X.T + Y 16383^2 (scalar tile)")
    timed(lambda: X2.transpose((1, 0)).reshape((m2 * m2,)), 8 * m2 * m2, "X.T.reshape 16383^2 (scalar tile)")
    del X2, Y2
    n = 1 << 26
    for c in (3,):
        Z = rnd((n, c), 3)
        timed(lambda: Z.transpose((1, 0)).reshape((n * c,)), 8 * n * c, f"[2^26,{c}].T copy (interleaved -> planar)", reps=3)
        Zt = rnd((c, n), 4)
        timed(lambda: Zt.transpose((1, 0)).reshape((n * c,)), 8 * n * c, f"[{c},2^26].T copy (planar -> interleaved)", reps=3)
        del Z, Zt
    tune(force_general=1)
    Z = rnd((n, 4), 3)
    timed(lambda: Z.transpose((1, 0)).reshape((n * 4,)), 8 * n * 4, "[2^26,4].T copy (scalar gather fallback)", reps=3)
    del Z
    timed(lambda: X.transpose((1, 0)) + Y, 12 * m * m, "X.T + Y 16384^2 (scalar gather fallback)", reps=3)
    tune(force_general=0)


def bench_window():
    n, w = 1 << 28, 64
    base = rnd((n,), 1)
    win = lambda ww: base.as_strided(shape=(n - ww + 1, ww), stride=(1, 1))  # noqa: E731
    for variant, per_sms in ((1, (2, 3, 4)), (2, (2, 3)), (3, (2,))):
        for per_sm in per_sms:
            tune(window_shfl=variant, window_ctas_per_sm=per_sm)
            timed(lambda: win(w).reduce_sum(dims=[1]), 4 * n + 4 * (n - w + 1), f"window-64 reduce_sum 2^28 shfl ring variant={variant} ctas/SM={per_sm}", reps=5)
    tune(window_shfl=0, window_ctas_per_sm=0)
    timed(lambda: win(w).reduce_sum(dims=[1]), 4 * n + 4 * (n - w + 1), "window-64 reduce_sum 2^28 shared-memory half-block kernel")
    tune(window_shfl=1)
    timed(lambda: win(w).reduce_max(dims=[1]), 4 * n + 4 * (n - w + 1), "window-64 reduce_max 2^28")
    for w2 in (8, 16, 32, 128, 24, 100, 500):
        for variant in (1, 0):
            tune(window_shfl=variant)
            timed(lambda: win(w2).reduce_sum(dims=[1]), 4 * n + 4 * (n - w2 + 1), f"window-{w2} reduce_sum 2^28 shfl={variant}", reps=3)
    tune(window_shfl=1)
    n2 = 1 << 24
    b2 = rnd((n2,), 1)
    tune(disable_window=1)
    timed(lambda: b2.as_strided(shape=(n2 - w + 1, w), stride=(1, 1)).reduce_sum(dims=[1]), 8 * n2, "window-64 reduce_sum 2^24 generic path", reps=3)
    tune(disable_window=0)


def smi_clock():
    import subprocess
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-i", "0"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
    except Exception as exc:  # noqa: BLE001
        out = repr(exc)
    return out


def bench_c5a_attrib():
    """Why is C5a slower inside the suite than alone? Time it after an idle gap, after a read-heavy op, after a
    write-heavy op and inside a long busy stretch, logging SM clock / power around each measurement."""
    import threading
    import time
    n, w = 1 << 28, 64
    base = rnd((n,), 1)
    X = rnd((64, 512, 512, 64), 2)
    a, b = rnd((n,), 3), rnd((n,), 4)
    c5a = lambda: base.as_strided(shape=(n - w + 1, w), stride=(1, 1)).reduce_sum(dims=[1])  # noqa: E731
    nbytes = 4 * n + 4 * (n - w + 1)
    samples, stop = [], threading.Event()

    def sampler():
        while not stop.is_set():
            samples.append((time.perf_counter(), smi_clock()))
            time.sleep(0.05)

    def one(label, before=None, reps=5):
        ts = []
        for _ in range(reps):
            if before:
                r0 = before()
                del r0
            lib.al_event_record(0)
            r = c5a()
            lib.al_event_record(1)
            ms = C.c_float()
            assert lib.al_event_elapsed(0, 1, C.byref(ms)) == 0
            ts.append(ms.value)
            del r
        print(f"{label:64s} median {statistics.median(ts):.4f} ms = {nbytes / statistics.median(ts) / 1e6:7.1f} GB/s  all {[round(t, 4) for t in ts]}  [{lib.al_last_kernel().decode()}] smi: {smi_clock()}", flush=True)

    for variant, per_sm in ((0, 0), (1, 3), (2, 2), (3, 2)):
        tune(window_shfl=variant, window_ctas_per_sm=per_sm)
        print(f"--- window_shfl={variant} ctas/SM={per_sm}", flush=True)
        for _ in range(3):
            r = c5a()
            del r
        al.synchronize()
        time.sleep(1.0)
        one("after a 1 s idle gap")
        one("after reduce_sum(X, dims=(0,)) [64 MiB output left dirty in L2]", lambda: X.reduce_sum(dims=(0,)))
        th = threading.Thread(target=sampler, daemon=True)
        th.start()
        t_end = time.perf_counter() + 2.5
        while time.perf_counter() < t_end:   # long busy stretch: the power cap sets in here
            for _ in range(20):
                r = a + b
                del r
                r = X.reduce_sum(dims=(0,))
                del r
            al.synchronize()
        one("inside a busy stretch (after reduce dims=(0,))", lambda: X.reduce_sum(dims=(0,)), reps=9)
        stop.set()
        th.join()
        stop.clear()
        print("   smi during the busy stretch (sm MHz, W, sw_power_cap):", [s for _, s in samples[:: max(len(samples) // 6, 1)]], flush=True)
        samples.clear()
    tune(window_shfl=1, window_ctas_per_sm=0)


def bench_reduce():
    X = rnd((64, 512, 512, 64), 1)
    n = 1 << 30
    for dims in ((1, 2), (3,), (0,), (0, 1, 2, 3), (0, 1, 2), (2, 3), (1,)):
        timed(lambda: X.reduce_sum(dims=dims), 4 * n, f"reduce_sum [64,512,512,64] dims={dims}")
    timed(lambda: X.reduce_max(dims=(1, 2)), 4 * n, "reduce_max dims=(1,2)")
    for splits in (9, 18, 37, 74, 148, 296):
        tune(reduce_splits=splits)
        timed(lambda: X.reduce_sum(dims=(1, 2)), 4 * n, f"reduce_sum dims=(1,2) splits={splits}")
    tune(reduce_splits=0)


def bench_fused():
    n = 1 << 28
    a, b, c = rnd((n,), 1), rnd((n,), 2), rnd((n,), 3)
    timed(lambda: a * b + c, 12 * n + 12 * n, "eager a*b+c 2^28 (two passes, 24 B/elem algorithmic)")
    timed(lambda: (al.lazy(a) * b + c).eval(), 16 * n, "fused a*b+c 2^28 (16 B/elem)")
    timed(lambda: ((al.lazy(a) + b) * 0.5 - 1.0).eval(), 12 * n, "fused (a+b)*0.5-1 2^28 (12 B/elem)")
    timed(lambda: (al.lazy(a).abs().sqrt() * 2.0).eval(), 8 * n, "fused sqrt(|a|)*2 2^28 (8 B/elem)")
    timed(lambda: al.apply(lambda v: 2 * v * v + 1, a), 8 * n, "apply(lambda v: 2*v*v+1) FP64 tape 2^28")
    timed(lambda: al.apply(lambda v: 1 / (1 + v * v) - v / 3, a), 8 * n, "apply(rational lambda) FP64 tape 2^28")
    A, B, cc = rnd((1024, 1, 4096), 1), rnd((1, 1024, 4096), 2), rnd((4096,), 3)
    timed(lambda: (al.lazy(A) + B + cc + 1.0).eval(), 4 * (1 << 32), "fused C3 chain A+B+c+1.0 (README-style, 17.2 GB)")


def bench_pcie():
    import time
    n = 1 << 30
    buf = al.host_buffer((n,))
    buf[:] = 1.0
    for label, fn in (("H2D 4 GiB pinned (blocking)", lambda: al.from_numpy(buf)),):
        fn()
        t0 = time.perf_counter(); x = fn(); al.synchronize(); dt = time.perf_counter() - t0
        print(f"{label:40s} {dt*1e3:8.1f} ms  {4*n/dt/1e9:6.1f} GB/s", flush=True)
    X = al.from_numpy(buf)
    t0 = time.perf_counter(); al.to_numpy(X, out=buf); dt = time.perf_counter() - t0
    print(f"{'D2H 4 GiB pinned (blocking)':40s} {dt*1e3:8.1f} ms  {4*n/dt/1e9:6.1f} GB/s", flush=True)
    buf2 = al.host_buffer((n,))
    t0 = time.perf_counter()
    Y = al.from_numpy(buf2, blocking=False); al.to_numpy(X, out=buf, blocking=False); al.synchronize()
    dt = time.perf_counter() - t0
    print(f"{'H2D + D2H 4 GiB each, concurrent':40s} {dt*1e3:8.1f} ms  {8*n/dt/1e9:6.1f} GB/s total", flush=True)
    pageable = np.ones(n, np.float32)
    t0 = time.perf_counter(); Z = al.from_numpy(pageable); al.synchronize(); dt = time.perf_counter() - t0
    print(f"{'H2D 4 GiB pageable':40s} {dt*1e3:8.1f} ms  {4*n/dt/1e9:6.1f} GB/s", flush=True)
    t0 = time.perf_counter(); h = al.to_numpy(Z); dt = time.perf_counter() - t0
    print(f"{'D2H 4 GiB pageable (fresh np.empty)':40s} {dt*1e3:8.1f} ms  {4*n/dt/1e9:6.1f} GB/s", flush=True)


def bench_window_any():
    """Window widths outside 8/16/32/64/128: 32 elements per lane (round-2 kernel) against 16 (w <= 128)."""
    n = 1 << 28
    base = rnd((n,), 1)
    for w in (12, 100):
        for l16 in (0, 1):
            tune(window_any_l16=l16)
            timed(lambda: base.as_strided(shape=(n - w + 1, w), stride=(1, 1)).reduce_sum(dims=[1]), 8 * n, f"window-{w} reduce_sum 2^28 window_any_l16={l16}", reps=3, warm=1)
    tune(window_any_l16=1)
    timed(lambda: base.as_strided(shape=(n - 99, 100), stride=(1, 1)).reduce_max(dims=[1]), 8 * n, "window-100 reduce_max 2^28", reps=3, warm=1)
    timed(lambda: base.as_strided(shape=(n - 63, 64), stride=(1, 1)).reduce_max(dims=[1]), 8 * n, "window-64 reduce_max 2^28", reps=3, warm=1)
    X = base.reshape((16, 512, 512, 64))
    timed(lambda: X.reduce_max(dims=(3,)), 4 * n + 4 * (n // 64), "reduce_max [16,512,512,64] dims=(3,)", reps=3, warm=1)
    timed(lambda: X.reduce_min(dims=(1, 2)), 4 * n, "reduce_min [16,512,512,64] dims=(1,2)", reps=3, warm=1)


def bench_stock():
    """The stock signatures on pageable numpy arrays: upload alone, download alone, both at once (second host thread)."""
    import threading
    import time
    n = 1 << 28
    src = np.ones(n, np.float32)
    dst = np.empty(n, np.float32)
    X = al.from_numpy(src)
    al.synchronize()
    best = {"up": 0.0, "down": 0.0, "both": 0.0}
    for _ in range(3):
        t0 = time.perf_counter(); Y = al.from_numpy(src); al.synchronize(); dt = time.perf_counter() - t0
        best["up"] = max(best["up"], 4 * n / dt / 1e9)
        t0 = time.perf_counter(); al.to_numpy(X, out=dst); dt = time.perf_counter() - t0
        best["down"] = max(best["down"], 4 * n / dt / 1e9)
        th = threading.Thread(target=lambda: al.to_numpy(X, out=dst))
        t0 = time.perf_counter(); th.start(); Z = al.from_numpy(src); th.join(); al.synchronize(); dt = time.perf_counter() - t0
        best["both"] = max(best["both"], 8 * n / dt / 1e9)
        del Y, Z
    print(f"stock API 1 GiB pageable: slot {os.environ.get('ARRAYLIB_RING_SLOT_MB', '32')} MiB, tuning {os.environ.get('ARRAYLIB_TUNING', '-')}, "
          f"copy threads {os.environ.get('ARRAYLIB_COPY_THREADS', 'auto')}: upload {best['up']:.1f} GB/s, download {best['down']:.1f} GB/s, both {best['both']:.1f} GB/s total", flush=True)


def bench_latency():
    """Host-side cost per op on tiny arrays (the C1 / README regime)."""
    import time
    a, b = al.ones((3, 6)), al.ones((6,))
    for label, fn in (("a + b (tiny, broadcast)", lambda: a + b), ("a + 1.0 (tiny)", lambda: a + 1.0),
                      ("a.reshape", lambda: a.reshape((6, 3))), ("a.reduce_sum(dims=(1,))", lambda: a.reduce_sum(dims=(1,))),
                      ("a.shape property", lambda: a.shape)):
        for _ in range(200):
            fn()
        al.synchronize()
        t0 = time.perf_counter()
        for _ in range(5000):
            r = fn()
        al.synchronize()
        dt = (time.perf_counter() - t0) / 5000
        print(f"{label:34s} {dt*1e6:8.2f} us per call", flush=True)


def bench_reduce_sweep():
    """reduce_sum over many layouts: look for plans far below the roofline."""
    n = 1 << 28
    cases = [((n,), (0,)), ((1 << 14, 1 << 14), (0,)), ((1 << 14, 1 << 14), (1,)), ((1 << 26, 4), (0,)), ((1 << 26, 4), (1,)),
             ((4, 1 << 26), (0,)), ((4, 1 << 26), (1,)), ((1 << 25, 8), (0,)), ((3, 1 << 20, 85), (0, 2)), ((3, 1 << 20, 85), (1,)),
             ((1024, 1024, 256), (0, 2)), ((1024, 1024, 256), (1,)), ((1024, 1024, 256), (0, 1)), ((1024, 1024, 256), (1, 2)),
             ((256, 1024, 1024), (0,)), ((64, 64, 64, 1024), (1, 3)), ((64, 64, 64, 1024), (0, 2)), ((1 << 20, 255), (1,)),
             ((1 << 20, 255), (0,)), ((1 << 27, 2), (1,)), ((1 << 16, 1 << 12), (1,))]
    for shape, dims in cases:
        X = rnd(shape, 1)
        total = int(np.prod(shape))
        timed(lambda: X.reduce_sum(dims=dims), 4 * total, f"reduce_sum {shape} dims={dims}", reps=3, warm=1)
        del X
    for per_lane in (4, 8, 16, 32):
        tune(reduce_loads_per_lane=per_lane)
        for shape, dims in (((1 << 20, 255), (1,)), ((3, 1 << 20, 85), (0, 2)), ((1 << 24, 64), (1,))):
            X = rnd(shape, 1)
            timed(lambda: X.reduce_sum(dims=dims), 4 * int(np.prod(shape)), f"loads/lane={per_lane}: reduce_sum {shape} dims={dims}", reps=3, warm=1)
            del X
    tune(reduce_loads_per_lane=0)
    # transposed sources
    X = rnd((1 << 14, 1 << 14), 2)
    XT = X.transpose((1, 0))
    for dims in ((0,), (1,), (0, 1)):
        timed(lambda: XT.reduce_sum(dims=dims), 4 * (1 << 28), f"reduce_sum of X.T (16384^2) dims={dims}", reps=3, warm=1)


def bench_ew_sweep():
    """Broadcast patterns: look for elementwise plans far below the roofline (bytes = output + distinct inputs)."""
    cases = [((1 << 14, 1), (1, 1 << 14)), ((1 << 14, 1 << 14), (1 << 14, 1)), ((1 << 14, 1 << 14), (1, 1 << 14)),
             ((1 << 28,), (1,)), ((1,), (1 << 28,)), ((64, 1, 512, 512), (1, 16, 1, 1)), ((64, 16, 512, 512), (1, 16, 1, 1)),
             ((64, 16, 512, 512), (64, 1, 512, 512)), ((64, 16, 512, 512), (512, 1)), ((1 << 20, 1, 16), (1, 16, 16)),
             ((4096, 1, 4096, 1), (1, 4, 1, 4)), ((1 << 26, 3), (3,)), ((1 << 26, 3), (1 << 26, 1)), ((1 << 22, 1, 7), (1, 9, 7)),
             ((1 << 14, 1, 3), (1, 1 << 12, 3)), ((1 << 12, 1, 5, 3), (1, 1 << 11, 1, 3))]
    for sa, sb in cases:
        A, B = rnd(sa, 1), rnd(sb, 2)
        out_shape = np.broadcast_shapes(sa, sb)
        nbytes = 4 * (int(np.prod(out_shape)) + int(np.prod(sa)) + int(np.prod(sb)))
        timed(lambda: A + B, nbytes, f"{sa} + {sb}", reps=3, warm=1)
        del A, B


def bench_views():
    """Elementwise ops on views that defeat the 128-bit path."""
    n = 1 << 28
    a, b = rnd((n + 8,), 1), rnd((n + 8,), 2)
    timed(lambda: a[0:n] + b[0:n], 12 * n, "a[0:n] + b[0:n] (aligned slices)")
    timed(lambda: a[1:n + 1] + b[1:n + 1], 12 * n, "a[1:n+1] + b[1:n+1] (both misaligned by 4 B)")
    timed(lambda: a[1:n + 1] + b[0:n], 12 * n, "a[1:n+1] + b[0:n] (one misaligned)")
    timed(lambda: a[0:n:2] + b[0:n:2], 12 * (n // 2), "a[::2] + b[::2] (inner stride 2)")
    timed(lambda: al.copy(a[3:n + 3], deep=True), 8 * n, "deep copy of a[3:n+3]")
    m = 16384
    X = rnd((m, m + 4), 3)
    timed(lambda: X[0:m, 1:m + 1] + 1.0, 8 * m * m, "X[:, 1:m+1] + 1.0 (rows misaligned, pitch m+4)")
    timed(lambda: X[0:m, 0:m] + 1.0, 8 * m * m, "X[:, 0:m] + 1.0 (rows aligned, pitch m+4)")


def bench_ops():
    n = 1 << 28
    a, b = rnd((n,), 1), rnd((n,), 2)
    pos = al.abs(a) + 0.5
    # (max / min measure 6.3 TB/s wherever they stand in this list -- before or after pow -- so it is the kernels, not the clocks)
    for name in ("add", "sub", "mul", "div", "mod", "max", "min", "eq", "ne", "gt", "ge", "lt", "le", "pow"):
        x = pos if name == "pow" else a
        timed(lambda: getattr(al, name)(x, b), 12 * n, f"binary {name} 2^28", reps=3, warm=1)
    unit = al.sin(a)        # (-1, 1): the domain of asin / acos / atanh
    above1 = pos + 0.5      # >= 1: acosh
    for name in ("neg", "abs", "sqrt", "exp", "log", "sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh",
                 "asinh", "acosh", "atanh", "ceil", "floor"):
        x = unit if name in ("asin", "acos", "atanh") else above1 if name == "acosh" else pos if name in ("log", "sqrt") else a
        timed(lambda: getattr(al, name)(x), 8 * n, f"unary {name} 2^28", reps=3, warm=1)


def bench_heavy():
    """FP64 ufuncs / pow on a contiguous stream: ew_nd (load, compute, store, exit) against the persistent kernel that
    keeps the next batch's loads in flight while it computes."""
    n = 1 << 28
    a, b = rnd((n,), 1), rnd((n,), 2)
    pos = al.abs(a) + 0.5
    for name, x in (("exp", a), ("log", pos), ("sin", a), ("tanh", a), ("asinh", a), ("atan", a)):
        for hs, ctas in ((0, 0), (2, 0), (2, 6)):
            tune(heavy_stream=hs, heavy_ctas_per_sm=ctas)
            timed(lambda: getattr(al, name)(x), 8 * n, f"unary {name} 2^28 heavy_stream={hs} ctas/SM={ctas or 'auto'}", reps=3, warm=1)
    for hs, ctas in ((0, 0), (1, 0)):
        tune(heavy_stream=hs, heavy_ctas_per_sm=ctas)
        timed(lambda: al.pow(pos, b), 12 * n, f"binary pow 2^28 heavy_stream={hs} ctas/SM={ctas or 'auto'}", reps=3, warm=1)
        timed(lambda: pos ** 1.5, 8 * n, f"scalar pow 2^28 heavy_stream={hs} ctas/SM={ctas or 'auto'}", reps=3, warm=1)
    tune(heavy_stream=1, heavy_ctas_per_sm=0)


def bench_period():
    """Write-bound broadcasts with short inner axes: ew_run (period_plan=0) against the period-table kernel."""
    cases = [((1 << 22, 1, 7), (1, 9, 7)), ((1 << 14, 1, 3), (1, 1 << 12, 3)), ((1 << 12, 1, 5, 3), (1, 1 << 11, 1, 3)),
             ((1 << 20, 1, 6), (1, 40, 6)), ((1 << 16, 5, 1, 3), (1 << 16, 1, 100, 3)), ((4096, 1, 4096, 1), (1, 4, 1, 4))]
    for sa, sb in cases:
        A, B = rnd(sa, 1), rnd(sb, 2)
        out_shape = np.broadcast_shapes(sa, sb)
        nbytes = 4 * (int(np.prod(out_shape)) + int(np.prod(sa)) + int(np.prod(sb)))
        for pp, ot in ((0, 1), (1, 1), (1, 0)):
            tune(period_plan=pp, outer_tile=ot)
            timed(lambda: A + B, nbytes, f"{sa} + {sb} period_plan={pp} outer_tile={ot}", reps=3, warm=1)
        tune(period_plan=1, outer_tile=1)
        del A, B


if __name__ == "__main__":
    which = sys.argv[1:] or ["fill", "ew", "bcast", "transpose", "window", "reduce"]
    for w in which:
        globals()[f"bench_{w}"]()
