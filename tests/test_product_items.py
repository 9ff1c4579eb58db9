"""
Host-side check of the tile -> CTA tables of the int8 fit-path products (``gpso_debug_product_items``, no GPU needed).
The tables are replayed in numpy on a random lower-triangular matrix with 128 x 128 "tiles" shrunk to T x T: executing
exactly the listed tiles with exactly the listed k-ranges must reproduce K^-1 = L^-T L^-1 (kind 0) and, level by level,
the recursive-doubling inverse (kind 1); every tile must be listed once and the deal must be balanced.
"""
import ctypes

import numpy as np
import pytest

from pygpso_b200 import backend

T = 4           # rows per tile in the replay (128 on the device)
KS = T // 4     # columns per k-step (32 on the device): 4 k-steps per tile, as on the device


def product_items(kind, nb, nsm=148):
    lib = backend.load_library()
    ip = ctypes.POINTER(ctypes.c_int)
    levels = ctypes.c_int(0)
    info = np.zeros(64, dtype=np.int32)
    n = lib.gpso_debug_product_items(kind, nb, nsm, None, 0, info.ctypes.data_as(ip), info.size, ctypes.byref(levels))
    assert n > 0
    buf = np.zeros(n, dtype=np.int32)
    assert lib.gpso_debug_product_items(kind, nb, nsm, buf.ctypes.data_as(ip), n, info.ctypes.data_as(ip), info.size, ctypes.byref(levels)) == n
    return buf, info, levels.value


def random_factor(nb, seed):
    rng = np.random.default_rng(seed)
    n = nb * T
    L = np.tril(rng.normal(size=(n, n))) * 0.3
    L[np.arange(n), np.arange(n)] = 1.0 + rng.random(n)
    return L


@pytest.mark.parametrize("nb", [1, 2, 3, 5, 8, 13, 32, 64])
def test_kinv_table_covers_the_lower_triangle_once(nb):
    nsm = 148
    buf, info, _ = product_items(0, nb, nsm)
    rounds = int(info[0])
    table = buf.reshape(rounds, nsm)
    L = random_factor(nb, nb)
    Lit = np.linalg.inv(L).T                       # L^-T: row i holds column i of L^-1 (k >= i)
    out = np.full((nb * T, nb * T), np.nan)
    seen = set()
    load = np.zeros(nsm)
    for g in range(nsm):
        for code in table[:, g]:
            if code < 0:
                continue
            I, ct = int(code) >> 16, int(code) & 0xFFFF
            assert (I, ct) not in seen and 0 <= ct <= 2 * I + 1
            seen.add((I, ct))
            rows = slice(I * T, (I + 1) * T)
            cols = slice(ct * (T // 2), (ct + 1) * (T // 2))   # 64-wide column tile = half a row block
            k0 = 4 * I * KS                                     # contraction over k-steps [4 I, nks)
            out[rows, cols] = Lit[rows, k0:] @ Lit[cols, k0:].T
            load[g] += (4 * nb - 4 * I) + 6
    assert len(seen) == nb * (nb + 1)
    want = np.linalg.inv(L @ L.T)
    low = np.tril_indices(nb * T)
    np.testing.assert_allclose(out[low], want[low], rtol=0, atol=1e-9 * np.abs(want).max())
    if nb >= 32:
        assert load.max() <= 1.05 * load.mean() + 4 * nb  # longest-first deal: within one long tile of the mean


@pytest.mark.parametrize("nb", [2, 3, 5, 8, 9, 13, 32, 33, 64])
def test_inverse_tables_replay_the_recursive_doubling(nb):
    nsm = 148
    buf, info, levels = product_items(1, nb, nsm)
    L = random_factor(nb, 100 + nb)
    n = nb * T
    Linv = np.zeros((n, n))
    for p in range(nb):                                        # level 0: the diagonal blocks are inverted by the DIAG tasks
        blk = slice(p * T, (p + 1) * T)
        Linv[blk, blk] = np.linalg.inv(L[blk, blk])
    LinvT = Linv.T.copy()
    XT = np.zeros((n, n))
    s = 1
    for lv in range(levels):
        s_lv, xt_off, xt_rounds, y_off, y_rounds = (int(v) for v in info[5 * lv:5 * lv + 5])
        assert s_lv == s
        for off, rounds, step in ((xt_off, xt_rounds, "xt"), (y_off, y_rounds, "y")):
            table = buf[off:off + rounds * nsm * 4].reshape(rounds, nsm, 4)
            seen = set()
            for I, ct, ks0, nk in table.reshape(-1, 4):
                if I < 0:
                    continue
                assert (I, ct) not in seen
                seen.add((int(I), int(ct)))
                rows = slice(I * T, (I + 1) * T)
                cols = slice(ct * (T // 2), (ct + 1) * (T // 2))
                k = slice(ks0 * KS, (ks0 + nk) * KS)
                if step == "xt":    # X^T[u rows, v cols] = L11^-T[u][k] . L[v][k]
                    XT[rows, cols] = LinvT[rows, k] @ L[cols, k].T
                else:               # L21^-1[v rows, u cols] = -(L22^-1[v][k] . X^T[u][k]) and its transpose
                    y = -(Linv[rows, k] @ XT[cols, k].T)
                    Linv[rows, cols] = y
                    LinvT[cols, rows] = y.T
            pairs = [(q, min(s, nb - (2 * q * s + s))) for q in range((nb + 2 * s - 1) // (2 * s)) if nb - (2 * q * s + s) > 0]
            assert len(seen) == sum(2 * s * nv for _, nv in pairs)
        s *= 2
    assert s >= nb
    np.testing.assert_allclose(Linv @ L, np.eye(n), rtol=0, atol=1e-9)
    np.testing.assert_allclose(LinvT, Linv.T, rtol=0, atol=0)
