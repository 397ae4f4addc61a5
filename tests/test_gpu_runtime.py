"""Runtime behaviour behind the C ABI (csrc/runtime.cu): asynchronous transfers, the block cache, error slots."""
import ctypes as C
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_async_downloads_are_reaped_without_synchronize(al):
    """a pipeline of non-blocking downloads that never calls synchronize() must not pin every
    downloaded block forever. Finished downloads are reaped when the next one is queued."""
    from arraylib_b200._lib import lib
    al.synchronize()
    base = lib.al_bytes_in_use()
    n = 1 << 20  # 4 MiB per result
    stage = al.host_buffer((n,))
    for i in range(64):
        r = al.full((n,), float(i))
        al.to_numpy(r, out=stage, blocking=False)
        del r
        if i % 8 == 7:
            time.sleep(0.01)  # lets the copies finish; NO synchronize
    time.sleep(0.05)
    r = al.full((n,), 64.0)
    al.to_numpy(r, out=stage, blocking=False)  # queues one more download: reaps the finished ones
    del r
    held = lib.al_bytes_in_use() - base
    assert held <= 8 * 4 * n, f"{held} bytes still pinned by finished downloads"
    al.synchronize()
    assert lib.al_bytes_in_use() == base
    assert stage[0] == 64.0 and stage[-1] == 64.0


def test_in_place_write_waits_for_a_pending_download(al):
    """X[...] = v right after to_numpy(X, blocking=False) must not race with the copy."""
    n = 1 << 26  # 256 MiB: the D2H takes ~5 ms, the fill kernel 0.1 ms
    X = al.ones((n,))
    stage = al.host_buffer((n,))
    for value in (2.0, 3.0):
        expect = float(al.to_numpy(X[(slice(0, 1),)])[0])
        al.to_numpy(X, out=stage, blocking=False)
        X[(slice(0, n),)] = value                # compute stream: must be ordered AFTER the download
        al.synchronize()
        assert stage.min() == expect and stage.max() == expect, "the host saw a mix of old and new values"
        assert float(al.to_numpy(X[(slice(n - 1, n),)])[0]) == value
    Y = al.ones((n,))
    al.to_numpy(Y, out=stage, blocking=False)
    Y[(5,)] = 9.0                                # set_elem path
    al.synchronize()
    assert stage[5] == 1.0 and float(Y[(5,)]) == 9.0


def test_block_cache_is_bounded(al):
    from arraylib_b200._lib import lib
    al.synchronize()
    lib.al_trim_pool()
    assert lib.al_bytes_cached() == 0
    try:
        assert lib.al_set_tuning(b"cache_limit_mb", 96) == 0
        blocks = [al.zeros(((32 << 20) // 4 * k,)) for k in (1, 2, 3, 4)]  # 32, 64, 96, 128 MiB
        del blocks
        assert 0 < lib.al_bytes_cached() <= 96 << 20   # the largest idle blocks went back to the driver
        a = al.zeros(((32 << 20) // 4,))               # and the small one is still served from the cache
        del a
    finally:
        lib.al_set_tuning(b"cache_limit_mb", 0)
        lib.al_trim_pool()


def test_error_message_survives_internal_cleanup(al):
    """A failing call that releases temporaries on its way out still reports its own message."""
    a, b = al.ones((4, 3)), al.ones((5, 2))
    with pytest.raises((ValueError, AssertionError, RuntimeError)) as info:
        al.matmul(a, b)
    assert "libarraylib_b200 call failed" not in str(info.value)
    with pytest.raises(Exception) as info:
        al.dot(al.ones((4,)), al.ones((5,)))
    assert "libarraylib_b200 call failed" not in str(info.value)
