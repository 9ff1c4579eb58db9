"""
Host logic of the hybrid factorisation (no GPU needed): the step list the library executes for a matrix of nb tiles
(``gpso_debug_hybrid_plan``) is replayed here in numpy -- leaves by ``numpy.linalg.cholesky``, panel solve, Schur complement and
the merge of the two halves' inverses as plain matrix products -- and must reproduce the Cholesky factor and its inverse of
the whole matrix for every tile count, including ragged splits down to one-tile leaves.  The tile -> CTA tables of the two
int8 products (``gpso_debug_hybrid_items``) must cover every output tile exactly once with the right contraction range.
"""
import ctypes

import numpy as np
import pytest

from pygpso_b200 import backend

T = 3  # rows per tile in the replay (the plan is in tile units)
LEAF, PANEL, SCHUR, MERGE = range(4)


def plan(nb, leaf):
    lib = backend.load_library()
    count = lib.gpso_debug_hybrid_plan(nb, leaf, None, 0)
    assert count > 0
    buf = np.zeros(4 * count, dtype=np.int32)
    assert lib.gpso_debug_hybrid_plan(nb, leaf, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), buf.size) == count
    return buf.reshape(-1, 4)


@pytest.mark.parametrize("leaf", [1, 2, 5, 32])
@pytest.mark.parametrize("nb", [1, 2, 3, 5, 8, 9, 13, 17, 33, 34, 49, 64])
def test_plan_replayed_in_numpy_is_the_cholesky_factor_and_its_inverse(nb, leaf):
    ops = plan(nb, leaf)
    rng = np.random.default_rng(nb * 100 + leaf)
    n = nb * T
    B = rng.standard_normal((n, n))
    K = B @ B.T + n * np.eye(n)
    A = np.tril(K).copy()      # the library factorises in place, lower triangle
    Linv = np.zeros((n, n))
    covered = np.zeros(nb, dtype=int)
    for op, t0, nt, s in ops:
        lo, mid, hi = t0 * T, (t0 + s) * T, (t0 + nt) * T
        if op == LEAF:
            blk = np.tril(A[lo:hi, lo:hi])
            blk = blk + np.tril(blk, -1).T
            L = np.linalg.cholesky(blk)
            A[lo:hi, lo:hi] = L
            Linv[lo:hi, lo:hi] = np.linalg.inv(L)
            covered[t0:t0 + nt] += 1
            assert nt <= leaf
        elif op == PANEL:
            assert 0 < s < nt and s & (s - 1) == 0 and 2 * s >= nt  # largest power of two below the node size
            A[mid:hi, lo:mid] = A[mid:hi, lo:mid] @ Linv[lo:mid, lo:mid].T
        elif op == SCHUR:
            A[mid:hi, mid:hi] -= np.tril(A[mid:hi, lo:mid] @ A[mid:hi, lo:mid].T)
        else:
            assert op == MERGE
            Linv[mid:hi, lo:mid] = -Linv[mid:hi, mid:hi] @ A[mid:hi, lo:mid] @ Linv[lo:mid, lo:mid]
    assert np.all(covered == 1)  # the leaves tile the diagonal exactly
    L_ref = np.linalg.cholesky(K)
    assert np.allclose(np.tril(A), L_ref, rtol=0, atol=1e-10 * np.abs(L_ref).max())
    assert np.allclose(Linv, np.linalg.inv(L_ref), rtol=0, atol=1e-10)
    if nb <= leaf:
        assert len(ops) == 1 and ops[0][0] == LEAF
    else:
        kinds = [int(o[0]) for o in ops]
        assert kinds.count(PANEL) == kinds.count(SCHUR) == kinds.count(MERGE) == kinds.count(LEAF) - 1


def items(kind, s, n, nsm=148):
    lib = backend.load_library()
    rounds = ctypes.c_int(0)
    count = lib.gpso_debug_hybrid_items(kind, s, n, nsm, None, 0, ctypes.byref(rounds))
    assert count == rounds.value * nsm * 4
    buf = np.zeros(count, dtype=np.int32)
    assert lib.gpso_debug_hybrid_items(kind, s, n, nsm, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), buf.size, ctypes.byref(rounds)) == count
    return buf.reshape(rounds.value, nsm, 4)


@pytest.mark.parametrize("s,n", [(1, 2), (2, 3), (4, 5), (8, 13), (16, 17), (32, 34), (32, 64)])
def test_product_tables_cover_every_tile_once(s, n):
    for kind in (0, 1):
        table = items(kind, s, n)
        live = table[table[:, :, 0] >= 0]
        seen = {(int(i), int(jt)): (int(k0), int(nk)) for i, jt, k0, nk in live}
        assert len(seen) == len(live)  # no tile twice
        if kind == 0:   # L21 = A21 L11^-T: rows I >= s, 64-wide column tiles of the first s tiles, k over the tiles [0, J]
            want = {(i, 2 * j + h): (0, 4 * (j + 1)) for i in range(s, n) for j in range(s) for h in (0, 1)}
        else:           # A22 -= L21 L21^T: lower tiles (I, J), s <= J <= I, k over the s tiles of L21
            want = {(i, 2 * j + h): (0, 4 * s) for i in range(s, n) for j in range(s, i + 1) for h in (0, 1)}
        assert seen == want
        # dealt longest-first to the least-loaded CTA: no CTA carries more than the lightest one plus one longest tile
        load = np.where(table[:, :, 0] >= 0, table[:, :, 3] + 6, 0).sum(axis=0)
        assert load.max() - load.min() <= live[:, 3].max() + 6


def test_bad_arguments_are_rejected():
    lib = backend.load_library()
    assert lib.gpso_debug_hybrid_plan(0, 4, None, 0) == -1
    rounds = ctypes.c_int(0)
    assert lib.gpso_debug_hybrid_items(2, 1, 2, 148, None, 0, ctypes.byref(rounds)) == -1
    assert lib.gpso_debug_hybrid_items(0, 4, 4, 148, None, 0, ctypes.byref(rounds)) == -1
