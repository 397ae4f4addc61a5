"""Generates tests/golden/reference_vectors.npz by running the UNMODIFIED Python reference.

Run in the build container only (needs /root/reference, cffi and gcc):
    python tests/golden/make_golden.py
The reference tree is copied to a temp dir (it must write its .so next to its sources), built
with its own Makefile recipe (`make shared`, /root/reference/Makefile:27-28) and imported from
there; nothing of it is copied into this repo -- only the input/output arrays below are saved.
Every output is copied out of the reference's buffers immediately (its to_numpy is non-owning).
"""
import math
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.npz")


def main():
    tmp = tempfile.mkdtemp(prefix="al_ref_")
    shutil.copytree("/root/reference", os.path.join(tmp, "ref"))
    root = os.path.join(tmp, "ref")
    subprocess.check_call(["make", "-s", "shared", "CC=/usr/bin/gcc"], cwd=root)
    sys.path.insert(0, root)
    import arraylib as ral  # the reference package

    def host(a):
        """Logical values of a reference array (its to_numpy ignores the view offset, SURVEY Q4,
        so materialise through a deep copy first)."""
        d = ral.copy(a, deep=True)
        return np.array(ral.to_numpy(d), dtype=np.float32, copy=True)

    def up(a):
        """numpy -> reference array (its from_numpy needs an even element count)."""
        return ral.array(a.tolist())

    g = {}
    # ---- config 1: the README pipeline (README.md:15-28) -------------------------------------
    a0 = ral.arange(1, 1 + 3 * 4 * 5 * 6).reshape((3, 4, 5, 6)).reduce_sum(dims=(1, 2))
    g["c1_reduce"] = host(a0)
    g["c1_reduce_shape"] = np.array(a0.shape)
    g["c1_reduce_stride"] = np.array(a0.stride)
    a1 = a0.apply(lambda x: math.sqrt(x))
    g["c1_sqrt"] = host(a1)
    a2 = a1.reshape((3, 6)).transpose((1, 0))
    g["c1_a"] = host(a2)
    g["c1_a_shape"] = np.array(a2.shape)
    g["c1_a_stride"] = np.array(a2.stride)
    g["c1_a_view"] = np.array(a2.view)
    b = ral.ones([3, 6]) + ral.ones([6]) + ral.ones([4, 3, 1]) + 1.0
    g["c1_b"] = host(b)
    g["c1_b_stride"] = np.array(b.stride)
    c = ral.arange(1, 11).as_strided(shape=(8, 3), stride=(1, 1)).reduce_sum(dims=[0])
    g["c1_c"] = host(c)

    # ---- arange across the 2^24 boundary ------------------------------------------
    lo = (1 << 24) - 6
    g["arange_sat_args"] = np.array([lo, lo + 40, 1])
    g["arange_sat"] = host(ral.arange(lo, lo + 40, 1))
    g["arange_step3_args"] = np.array([lo, lo + 200, 3])
    g["arange_step3"] = host(ral.arange(lo, lo + 200, 3))
    g["arange_step2_args"] = np.array([(1 << 25) - 9, (1 << 25) + 60, 2])
    g["arange_step2"] = host(ral.arange((1 << 25) - 9, (1 << 25) + 60, 2))
    g["arange_small"] = host(ral.arange(1, 10, 2))
    g["linspace_0_10_5"] = host(ral.linspace(0, 10, 5))
    g["linspace_1_100_10"] = host(ral.linspace(1, 100, 10))
    g["linspace_3_7_11"] = host(ral.linspace(3, 7, 11))

    # ---- broadcast binary ops on distinct values -----------------------------------------------
    rng = np.random.default_rng(2024)
    x = rng.standard_normal((5, 1, 7)).astype(np.float32)
    y = rng.standard_normal((3, 1)).astype(np.float32)
    x[0, 0, :4] = [0.0, -0.0, np.inf, np.nan]
    y[0, 0] = -0.0
    g["bin_x"], g["bin_y"] = x, y
    rx, ry = up(x), up(y)
    for op in ["sum", "sub", "mul", "div", "mod", "pow", "max", "min", "eq", "ne", "gt", "ge", "lt", "le"]:
        g[f"bin_{op}"] = host(ral.core.NDArray(getattr(ral.core.lib, f"array_array_{op}")(rx.buffer, ry.buffer)))
        g[f"sc_{op}"] = host(ral.core.NDArray(getattr(ral.core.lib, f"array_scalar_{op}")(rx.buffer, 0.75)))
    for op in ["neg", "abs", "sqrt", "exp", "log", "sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh",
               "asinh", "acosh", "atanh", "ceil", "floor"]:
        g[f"un_{op}"] = host(getattr(ral, op)(rx))

    # ---- apply() with arithmetic callbacks (the lambda runs per element in Python double precision) ----
    pos = np.abs(x) + np.float32(0.25)
    pos[~np.isfinite(pos)] = np.float32(3.5)
    g["apply_in"] = pos
    rpos = up(pos)
    g["apply_poly"] = host(rpos.apply(lambda v: 2 * v * v + 1))
    g["apply_rational"] = host(rpos.apply(lambda v: 1 / (1 + v * v) - v / 3))
    g["apply_mixed"] = host(rpos.apply(lambda v: abs(v - 2.5) * -1.0 + (v > 1.0) * v))

    # ---- reductions on distinct (small-integer) values, all dim subsets ---------------------------
    z = rng.integers(-8, 9, size=(4, 5, 6, 7)).astype(np.float32)
    g["red_z"] = z
    rz = up(z)
    for mask in range(16):
        dims = tuple(d for d in range(4) if mask >> d & 1)
        for op in ("sum", "max", "min"):
            g[f"red_{op}_{mask}"] = host(getattr(ral, f"reduce_{op}")(rz, dims=dims))

    # ---- views: transpose / deep copy / as_strided windows -----------------------------------------
    t = ral.arange(1, 1 + 3 * 4 * 5 * 6).reshape((3, 4, 5, 6)).transpose((1, 2, 0, 3))
    g["view_transpose_1203"] = host(t)
    g["view_transpose_1203_stride"] = np.array(t.stride)
    w = ral.arange(1, 41).as_strided(shape=(33, 8), stride=(1, 1))
    g["window_sum"] = host(w.reduce_sum(dims=[1]))
    g["window_plus"] = host(w + w)
    tt = ral.arange(0, 24).reshape((4, 6)).transpose((1, 0))
    g["t_plus"] = host(tt + ral.arange(0, 24).reshape((6, 4)))
    g["t_reshape"] = host(tt.reshape((24,)))  # forces array_deep_copy

    np.savez_compressed(OUT, **g)
    print("wrote", OUT, len(g), "arrays")
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
