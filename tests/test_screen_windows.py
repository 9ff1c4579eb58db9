"""
Host logic of the screening pass's window list (``gpso_debug_screen_windows``, no GPU needed): whatever the number of
candidates, the windows must tile [0, M) exactly once and in order, never exceed the window budget, start on multiples of 1024
(the digit-tile layout needs 128), ramp up through a quarter and a half window when the pass is pipelined over three or more
windows, and end without a short tail launch.
"""
import ctypes

import numpy as np
import pytest

from pygpso_b200 import backend


def windows(M, W, ramp=True, even=True):
    lib = backend.load_library()
    n = lib.gpso_debug_screen_windows(M, W, int(ramp), int(even), None, 0)
    assert n > 0
    buf = np.zeros(2 * n, dtype=np.int64)
    assert lib.gpso_debug_screen_windows(M, W, int(ramp), int(even), buf.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), buf.size) == n
    return buf.reshape(-1, 2)


@pytest.mark.parametrize("ramp,even", [(True, True), (False, True), (False, False)])
@pytest.mark.parametrize("W", [1024, 16384, 174080, 262144])
def test_windows_tile_the_candidates_exactly(W, ramp, even):
    rng = np.random.default_rng(W)
    sizes = [1, 127, 1024, W - 1, W, W + 1, 2 * W, 2 * W + 1, 3 * W - 5, 3 * W, 10_000_000, 1_250_000, 265_720]
    sizes += [int(v) for v in rng.integers(1, 40 * W, size=40)]
    for M in sizes:
        w = windows(M, W, ramp, even)
        assert w[0, 0] == 0 and w[-1, 0] + w[-1, 1] == M, (M, w)
        assert np.all(w[1:, 0] == w[:-1, 0] + w[:-1, 1]), (M, w)        # contiguous, in order, nothing twice
        assert np.all(w[:, 1] > 0) and np.all(w[:, 1] <= W), (M, w)
        assert np.all(w[:, 0] % 1024 == 0), (M, w)
        plain = -(-M // W)
        if ramp and plain >= 3:
            assert w[0, 1] == W // 4 // 1024 * 1024 or W < 4096, (M, w)
            assert len(w) <= plain + 2
        else:
            assert len(w) == plain, (M, w)
        if even and len(w) >= 2:
            body = w[2:] if (ramp and plain >= 3 and W >= 4096) else w
            if len(body) >= 2:  # no short tail: the last window is within 1024 per window of the others
                assert body[-1, 1] >= body[0, 1] - 1024 * len(body), (M, w)


def test_c3_shape_per_gpu_counts():
    """The windows of config C3 at 1 and 8 GPUs (262 144-candidate budget of the 2-digit rung)."""
    one = windows(10_000_000, 262_144)
    assert one[0, 1] == 65_536 and one[1, 1] == 131_072 and len(one) == 40
    eight = windows(1_250_000, 262_144)
    assert [int(v) for v in eight[:2, 1]] == [65_536, 131_072] and len(eight) == 7
    assert eight[2:, 1].min() > 200_000  # the 4 816-candidate tail of the plain split is gone


def test_bad_arguments():
    lib = backend.load_library()
    assert lib.gpso_debug_screen_windows(0, 1024, 1, 1, None, 0) == -1
    assert lib.gpso_debug_screen_windows(10, 1000, 1, 1, None, 0) == -1
