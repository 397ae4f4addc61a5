import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        from arraylib_b200._lib import lib
        return lib.al_init(-1) == 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def al():
    """The product package, bound to the GPU. GPU tests fail (not skip) when the CUDA library is
    missing or no device is visible -- there is no CPU fallback to fall back to."""
    import arraylib as al_mod
    assert _have_gpu(), "no CUDA device: -m gpu tests must run on the GPU box"
    return al_mod


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def ref():
    """The compiled, unmodified reference (oracle/_ref), or None when it has not been built."""
    from oracle import ref_binding
    return ref_binding.RefLib() if ref_binding.available() else None


@pytest.fixture
def rng():
    return np.random.default_rng(1234)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bit_equal(got, want, what="", nan_ok=False):
    """Bit-exact float32 comparison, NaN sign and payload included (the eager elementwise kernels follow the
    reference's x86/glibc NaN rules bit for bit). `nan_ok=True` lets any NaN match any NaN: only for the documented
    exceptions -- reductions (the tree order decides which NaN survives), fused chains / FP64 tapes (canonical GPU
    NaN, include/arraylib_b200.h:18-20) and comparisons against NumPy instead of the reference."""
    got = np.ascontiguousarray(got, dtype=np.float32)
    want = np.ascontiguousarray(want, dtype=np.float32)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    same = bits(got) == bits(want)
    if nan_ok:
        same |= np.isnan(got) & np.isnan(want)
    if not same.all():
        idx = np.argwhere(~same)[0]
        raise AssertionError(f"{what}: {int((~same).sum())} of {same.size} elements differ; first at {tuple(idx)}: "
                             f"got {got[tuple(idx)]!r} (0x{int(bits(got)[tuple(idx)]):08x}) want {want[tuple(idx)]!r} "
                             f"(0x{int(bits(want)[tuple(idx)]):08x})")
