"""
Hybrid factorisation (``gpso_set_factor_mode``: FP64 leaves, int8 tensor-core panels L21 = A21 L11^-T and Schur complements
A22 -= L21 L21^T, the inverse factor's merge level right behind them) on the GPU: LML, gradient and predictions against the
oracle at the same tolerances as the one-kernel schedule, on ragged tile counts (leaves of one, two and three tiles), at the
BASELINE shapes, and the LAPACK-style pivot report of a matrix that stops being positive definite inside a later leaf.
"""
import numpy as np
import pytest

from oracle import gpr_oracle as go
from pygpso_b200 import backend
from tests.test_gpu_parity import open_session, synthetic, theta_of

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    return backend.default_backend()


@pytest.mark.parametrize("kernel", ["Matern52", "SquaredExponential"])
@pytest.mark.parametrize("N,d", [(520, 2), (640, 3), (1100, 4), (1664, 5), (2100, 6)])
def test_forced_hybrid_lml_grad_vs_oracle(cuda, kernel, N, d):
    """Leaves of two tiles: 5 tiles = 4 + 1, 9 = 8 + 1, 13 = 8 + (4 + 1), 17 = 16 + 1 -- every branch of the recursion."""
    X, y = synthetic(N, d, seed=N)
    h = go.Hyper(0.25 * np.sqrt(d) * (1.0 + 0.2 * np.cos(np.arange(d))), 1.3, 2e-3, 0.1)
    u = h.pack()
    f_ref, g_ref = go.neg_lml_and_grad(kernel, X, y, u, d, True)
    s = open_session(cuda, kernel, X, y, n_ls=d)
    s.set_factor_mode(True, hybrid=True)
    f, g = s.neg_lml_and_grad(u)
    info = s.factor_info()
    s.set_factor_mode(True, hybrid=False)
    f1, g1 = s.neg_lml_and_grad(u)
    info1 = s.factor_info()
    s.close()
    assert info["schedule"] == "hybrid" and info["nodes"] >= 2, info
    assert info1["schedule"] == "persistent" and info1["nodes"] == 0, info1
    assert abs(f - f_ref) <= 1e-9 * max(abs(f_ref), N), (f, f_ref)
    err = np.abs(g - g_ref) / np.maximum(np.abs(g_ref), 1.0)
    assert np.all(err <= 1e-6), (float(err.max()), g, g_ref)
    # the two schedules agree far below the parity tolerance
    assert abs(f - f1) <= 1e-11 * max(abs(f1), N), (f, f1)


def test_forced_hybrid_predict_vs_oracle(cuda):
    N, d, M = 1664, 4, 3000
    X, y = synthetic(N, d, seed=31)
    h = go.Hyper(0.4, 1.0, 1e-3, 0.0)
    rng = np.random.default_rng(2)
    Xc = np.vstack([rng.random((M, d)), X[:100] + 1e-5 * rng.standard_normal((100, d))])
    mean_ref, var_ref = go.predict_y("Matern52", X, y, h, Xc)
    s = open_session(cuda, "Matern52", X, y)
    s.set_factor_mode(True, hybrid=True)
    s.factorize(theta_of(h))
    assert s.factor_info()["schedule"] == "hybrid"
    mean, var = s.predict_y(Xc)
    s.close()
    assert np.all(np.abs(mean - mean_ref[:, 0]) <= 1e-8 * np.maximum(np.abs(mean_ref[:, 0]), np.abs(y).max()))
    assert np.all(np.abs(var - var_ref[:, 0]) <= 1e-8 * np.maximum(np.abs(var_ref[:, 0]), h.variance))


def test_automatic_mode_uses_the_hybrid_above_4096_rows(cuda):
    N, d = 4300, 8  # 34 tiles = 32 + 2
    X, y = synthetic(N, d, seed=5)
    h = go.Hyper(0.25 * np.sqrt(d), 1.0, 1e-3, 0.05)
    u = h.pack()
    f_ref, g_ref = go.neg_lml_and_grad("Matern52", X, y, u, 1, True)
    s = open_session(cuda, "Matern52", X, y)
    f, g = s.neg_lml_and_grad(u)
    info = s.factor_info()
    s.close()
    assert info == {"schedule": "hybrid", "nodes": 1}, info
    assert abs(f - f_ref) <= 1e-9 * max(abs(f_ref), N), (f, f_ref)
    err = np.abs(g - g_ref) / np.maximum(np.abs(g_ref), 1.0)
    assert np.all(err <= 1e-6), (float(err.max()), g, g_ref)


@pytest.mark.parametrize("N", [1100, 4096])
def test_matrices_up_to_4096_rows_keep_the_single_kernel(cuda, N):
    X, y = synthetic(N, 3, seed=1)
    s = open_session(cuda, "Matern52", X, y)
    s.factorize(np.array([0.4, 1.0, 1e-3, 0.0]))
    assert s.factor_info() == {"schedule": "persistent", "nodes": 0}
    s.close()


@pytest.mark.parametrize("hybrid", [True, False])
def test_not_positive_definite_pivot_is_the_same_in_both_schedules(cuda, hybrid):
    """Duplicated rows with a vanishing noise variance: the factorisation breaks down at a pivot inside a later leaf; both
    schedules must finish and report a LAPACK-style pivot (which one is rounding-dependent)."""
    N, d = 1100, 2
    X, y = synthetic(N, d, seed=9)
    X[900:1050] = X[:150]
    s = open_session(cuda, "SquaredExponential", X, y)
    s.set_factor_mode(True, hybrid=hybrid)
    with pytest.raises(np.linalg.LinAlgError) as err:
        s.factorize(np.array([0.5, 1.0, 1.0e-18, 0.0]))
    s.close()
    text = str(err.value)
    assert "pivot" in text
    import re

    pivot = int(re.search(r"pivot (-?\d+)", text).group(1))
    assert 1 <= pivot <= N, text
