"""CPU checks of the drop-in boundary: the shared library loads without a GPU, exports every
symbol include/arraylib_b200.h declares, keeps the reference's struct layout, and refuses to
compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "arraylib_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b((?:array|al)_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    from arraylib_b200._lib import LIB_PATH
    lib = C.CDLL(LIB_PATH)
    names = declared_symbols()
    assert len(names) > 100
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_reference_entry_points_are_all_present():
    """Every symbol the reference's core.py binds."""
    from arraylib_b200._lib import BINOPS, UFUNCS, lib
    wanted = ["array_fill", "array_ones", "array_zeros", "array_arange", "array_linspace", "array_free",
              "array_array_matmul", "array_array_dot", "array_op", "array_reshape", "array_transpose",
              "array_move_dim", "array_ravel", "array_get_view", "array_get_elem", "array_set_elem_from_scalar",
              "array_set_view_from_scalar", "array_set_view_from_array", "array_deep_copy", "array_shallow_copy",
              "array_reduce", "array_reduce_sum", "array_reduce_max", "array_reduce_min", "array_array_array_where",
              "array_array_scalar_where", "array_scalar_array_where", "array_as_strided"]
    wanted += [f"array_array_{o}" for o in BINOPS] + [f"array_scalar_{o}" for o in BINOPS] + [f"array_{u}" for u in UFUNCS]
    assert all(hasattr(lib, w) for w in wanted)


def test_struct_layout_matches_reference_abi():
    from arraylib_b200._lib import CArray, CLayout, Data
    assert C.sizeof(Data) == 24 and Data.size.offset == 8 and Data.refs.offset == 16
    assert C.sizeof(CLayout) == 24 and CLayout.shape.offset == 8 and CLayout.stride.offset == 16
    assert C.sizeof(CArray) == 32 and CArray.ptr.offset == 8 and CArray.lay.offset == 16 and CArray.view.offset == 24


def test_no_cpu_fallback_without_a_device():
    import pytest
    from arraylib_b200._lib import lib
    if lib.al_init(-1) == 0:
        pytest.skip("a GPU is visible here")
    import arraylib as al
    with pytest.raises(RuntimeError, match="no CUDA device"):
        al.ones((2, 2))
    with pytest.raises(RuntimeError):
        al.arange(10)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under arraylib_b200/ or arraylib/ may import, load or
    mention it (a product path routed through the CPU oracle would void every parity claim)."""
    bad = []
    for pkg in ("arraylib_b200", "arraylib"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, pkg)):
            if os.sep + "build" in dirpath:
                continue
            for f in files:
                if not f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                    continue
                text = open(os.path.join(dirpath, f), errors="replace").read()
                for needle in ("import oracle", "from oracle", "liboracle", "oracle/_ref", "arraylib_ref", "ref_binding"):
                    if needle in text:
                        bad.append((os.path.join(dirpath, f), needle))
    assert not bad, bad


def test_reference_cffi_cdef_binds_completely():
    """The strongest form of the drop-in claim that can be checked without a GPU: take the REFERENCE's own
    header text (the part its core.py feeds to ffi.cdef, core.py:19-22), dlopen OUR library with it and
    resolve every function it declares. Only runs where /root/reference exists (the build container)."""
    import pytest
    header = "/root/reference/arraylib/src/arraylib.h"
    if not os.path.exists(header):
        pytest.skip("reference sources not present on this machine")
    cffi = pytest.importorskip("cffi")
    from arraylib_b200._lib import LIB_PATH
    ffi = cffi.FFI()
    lines = open(header).read().splitlines()
    text = "\n".join(lines[lines.index("// START"):lines.index("// END")])
    ffi.cdef(text)
    lib = ffi.dlopen(LIB_PATH)
    names = sorted(set(re.findall(r"\b([a-z_0-9]+)\s*\(", re.sub(r"/\*.*?\*/", "", text, flags=re.S))))
    names = [n for n in names if n not in ("f32", "clamp")]  # clamp is declared but never defined by the reference
    assert len(names) >= 85
    missing = []
    for n in names:
        try:
            getattr(lib, n)
        except AttributeError:
            missing.append(n)
    assert not missing, missing
    assert (ffi.sizeof("NDArray"), ffi.sizeof("Data"), ffi.sizeof("Layout")) == (32, 24, 24)
