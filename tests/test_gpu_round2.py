"""
Round-2 GPU tests (``-m gpu``, through the C ABI): handle life cycle, the BASELINE shapes that had no oracle parity
(config C4: N=8192, d=20), factor-cache invalidation of the model object, config C1 as BASELINE states it (n_workers=4).
"""
import numpy as np
import pytest

from oracle import gpr_oracle as go
from pygpso_b200 import GPRSurrogate, GPSOptimiser, ParameterSpace, backend, gpmodel
from tests.conftest import paper_objective
from tests.test_gpu_parity import VARSIGMA, open_session, synthetic, theta_of

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    return backend.default_backend()


# ---------------------------------------------------------------------------------------------------------------------
# life cycle: gpso_destroy releases every device buffer of the handle (round 1 leaked the int8 digit-tile buffers)
# ---------------------------------------------------------------------------------------------------------------------
def test_create_destroy_does_not_leak(cuda):
    import torch

    N, d = 1100, 4  # int8 engines on every stage (inverse factor, K_y^-1, variance product): all digit buffers allocated
    X, y = synthetic(N, d, seed=3)
    h = go.Hyper(0.5, 1.0, 1e-3, 0.0)
    Xc = np.random.default_rng(1).random((3000, d))

    def cycle():
        s = open_session(cuda, "Matern52", X, y)
        s.neg_lml_and_grad(h.pack())
        s.factorize(theta_of(h))
        s.ucb_argmax(Xc, VARSIGMA)
        s.ucb_topk(Xc, VARSIGMA, 4)
        s.close()

    for _ in range(3):
        cycle()
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info(cuda.device)
    for _ in range(200):
        cycle()
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info(cuda.device)
    # one handle at this size owns ~120 MB; 200 leaked handles would be 24 GB.  Allow the allocator a few MB of slack.
    assert free0 - free1 < 32 * 2 ** 20, (free0, free1)


def test_closed_sessions_feed_the_next_one_and_trim_returns_the_memory(cuda):
    """Blocks of a closed session stay in the library's pool (the optimiser opens one session per fit) and can be handed back."""
    import torch

    N, d = 700, 3
    X, y = synthetic(N, d, seed=6)
    theta = np.array([0.5, 1.0, 1e-3, 0.0])
    cuda.trim_pool()
    torch.cuda.synchronize()
    free_empty, _ = torch.cuda.mem_get_info(cuda.device)
    s = open_session(cuda, "Matern52", X, y)
    s.factorize(theta)
    first = s.ucb_argmax(np.random.default_rng(0).random((2000, d)), VARSIGMA)
    s.close()
    free_cached, _ = torch.cuda.mem_get_info(cuda.device)
    assert free_cached < free_empty  # the closed session's blocks are still held
    s = open_session(cuda, "Matern52", X, y)
    s.factorize(theta)
    again = s.ucb_argmax(np.random.default_rng(0).random((2000, d)), VARSIGMA)
    s.close()
    free_reused, _ = torch.cuda.mem_get_info(cuda.device)
    assert again == first  # recycled (not zeroed) blocks do not change the result
    assert free_cached - free_reused < 4 * 2 ** 20, (free_cached, free_reused)  # the second session ran on the first one's blocks
    released = cuda.trim_pool()
    free_after, _ = torch.cuda.mem_get_info(cuda.device)
    assert released > 0 and free_after >= free_cached + released - 4 * 2 ** 20, (released, free_after, free_cached)


def test_predict_after_loss_evaluation_refactorises(cuda):
    """predict -> training_loss (overwrites the device factor) -> predict at unchanged hyper-parameters."""
    X, y = synthetic(300, 3, seed=5)
    model = gpmodel.GPR(data=(X, y), kernel=gpmodel.Matern52(lengthscales=0.4), mean_function=gpmodel.Constant(0.0),
                        noise_variance=1e-3, backend=cuda)
    Xc = np.random.default_rng(2).random((500, 3))
    m0, v0 = model.predict_y(Xc)
    lml = model.log_marginal_likelihood()
    assert np.isfinite(lml)
    m1, v1 = model.predict_y(Xc)
    assert np.array_equal(m0, m1) and np.array_equal(v0, v1)
    best0 = model.ucb_argmax(Xc, VARSIGMA)
    model._session.set_predict_mode(1, 0)  # engine switch drops the factor too
    best1 = model.ucb_argmax(Xc, VARSIGMA)
    assert best0[0] == best1[0]
    model._session.set_kinv_mode(1)
    model.training_loss()
    assert model.grow_ucb_argmax(np.array([[0.0, 1.0]] * 3), 4, VARSIGMA)[0] >= 0
    model.close()


# ---------------------------------------------------------------------------------------------------------------------
# config C4 shape: N = 8192, d = 20 (recursion level s = 32, W = 4 blocking) against the oracle
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def c4_reference():
    N, d = 8192, 20
    X, y = synthetic(N, d)
    out = {"X": X, "y": y}
    h = go.Hyper(0.25 * np.sqrt(d), 1.0, 1e-3, 0.0)
    out["scalar"] = (h.pack() + 0.01, ) + go.neg_lml_and_grad("Matern52", X, y, h.pack() + 0.01, 1, True)
    ls = 0.25 * np.sqrt(d) * (1.0 + 0.3 * np.sin(np.arange(d)))
    ha = go.Hyper(ls, 1.1, 2e-3, 0.05)
    out["ard"] = (ha.pack(), ) + go.neg_lml_and_grad("Matern52", X, y, ha.pack(), d, True)
    return out


@pytest.mark.parametrize("ard", [False, True])
@pytest.mark.parametrize("engines", ["int8", "int8-persistent", "dmma", "int8-stepwise"])
def test_c4_shape_lml_grad_vs_oracle(cuda, c4_reference, ard, engines):
    """LML within 1e-9 * max(|LML|, N), gradient within 1e-6 * max(|g|, 1) at the shape the evals/s figure is quoted on."""
    X, y = c4_reference["X"], c4_reference["y"]
    N, d = X.shape
    u, f_ref, g_ref = c4_reference["ard" if ard else "scalar"]
    s = open_session(cuda, "Matern52", X, y, n_ls=d if ard else 1)
    if engines == "dmma":
        s.set_kinv_mode(1)
        s.set_inverse_mode(1)
    elif engines == "int8-stepwise":
        s.set_factor_mode(False)
    elif engines == "int8-persistent":
        s.set_factor_mode(True, hybrid=False)
    f, g = s.neg_lml_and_grad(u)
    schedule = s.factor_info()["schedule"]
    s.close()
    assert schedule == {"int8": "hybrid", "int8-persistent": "persistent", "dmma": "persistent", "int8-stepwise": "stepwise"}[engines]
    assert abs(f - f_ref) <= 1e-9 * max(abs(f_ref), N), (f, f_ref, abs(f - f_ref) / max(abs(f_ref), N))
    err = np.abs(g - g_ref) / np.maximum(np.abs(g_ref), 1.0)
    assert np.all(err <= 1e-6), (float(err.max()), g, g_ref)


def test_c4_shape_predict_vs_oracle(cuda, c4_reference):
    """predict_y at N = 8192, d = 20 on a sample of candidates, both engines, with the pure-relative error reported."""
    X, y = c4_reference["X"], c4_reference["y"]
    N, d = X.shape
    h = go.Hyper(0.25 * np.sqrt(d), 1.0, 1e-3, 0.0)
    rng = np.random.default_rng(9)
    Xc = np.vstack([rng.random((1500, d)), X[:200] + 1e-4 * rng.standard_normal((200, d))])  # incl. near-training points
    mean_ref, var_ref = go.predict_y("Matern52", X, y, h, Xc)
    for engine in ("int8", "dmma"):
        s = open_session(cuda, "Matern52", X, y, engine=engine)
        s.factorize(theta_of(h))
        mean, var = s.predict_y(Xc)
        s.close()
        mtol = 1e-8 * np.maximum(np.abs(mean_ref[:, 0]), np.abs(y).max())
        vtol = 1e-8 * np.maximum(np.abs(var_ref[:, 0]), h.variance)
        assert np.all(np.abs(mean - mean_ref[:, 0]) <= mtol)
        assert np.all(np.abs(var - var_ref[:, 0]) <= vtol)
        # pure-relative error of the variance (no floor): near training points var ~ 1e-3 and the subtraction
        # variance - |L^-1 k*|^2 cancels three digits, so the bound is 1e-8 * variance / var ~ 1e-5 there
        rel = np.abs(var - var_ref[:, 0]) / np.abs(var_ref[:, 0])
        assert rel.max() <= 1e-8 * h.variance / var_ref[:, 0].min() * 1.01, float(rel.max())


# ---------------------------------------------------------------------------------------------------------------------
# config C1 exactly as BASELINE.json states it: README 2-D problem, GPSOptimiser(n_workers=4), default budget
# (reference tests/test_optimisation.py:183-221 run the same with eval_repeats and a saver)
# ---------------------------------------------------------------------------------------------------------------------
from tests.test_optimiser_pool import StubSaver, objective_with_result, space_2d  # noqa: E402


@pytest.mark.timeout(600)
def test_c1_readme_example_with_worker_pool(cuda):
    """Objective evaluations in a 4-process pool (forked after CUDA initialisation; the children never touch CUDA) give
    the same run as the serial loop."""
    kw = dict(exploration_method="tree", exploration_depth=5, budget=40, stopping_condition="evaluations")
    serial = GPSOptimiser(parameter_space=space_2d(), n_workers=1, **kw)
    best_serial = serial.run(paper_objective)
    pooled = GPSOptimiser(parameter_space=space_2d(), n_workers=4, **kw)
    best_pooled = pooled.run(paper_objective)
    assert np.array_equal(best_serial.normed_coord, best_pooled.normed_coord)
    assert best_serial.score_mu == best_pooled.score_mu
    assert serial.gp_surr.num_evaluated == pooled.gp_surr.num_evaluated


@pytest.mark.timeout(900)
def test_c1_default_budget_n_workers_4(cuda):
    """BASELINE configs[0] verbatim: ParameterSpace x in [-3,5], y in [-3,3], GPSOptimiser(n_workers=4).run, default budget."""
    opt = GPSOptimiser(parameter_space=space_2d(), n_workers=4)
    best = opt.run(paper_objective)
    assert opt.n_eval_counter >= 100
    assert best.score_mu > 8.0  # global maximum of the README function is 8.1062
    # the same run driven by the oracle backend on the CPU takes the same decisions
    from tests.oracle_backend import OracleBackend

    ref = GPSOptimiser(parameter_space=space_2d(), n_workers=1, gp_surrogate=GPRSurrogate.default(backend=OracleBackend()))
    best_ref = ref.run(paper_objective)
    assert opt.iterations == ref.iterations and opt.n_eval_counter == ref.n_eval_counter
    assert np.allclose(best.normed_coord, best_ref.normed_coord, rtol=0, atol=1e-12)
    assert abs(best.score_mu - best_ref.score_mu) <= 1e-9


@pytest.mark.timeout(600)
def test_reference_v2_sample_method_pool_repeats(cuda):
    """Reference tests/test_optimisation.py:154-181 on the GPU: sample method, 12 iterations, 4 workers, 4 repeats."""
    opt = GPSOptimiser(parameter_space=space_2d(), exploration_method="sample", exploration_depth=3, budget=12,
                       stopping_condition="iterations", update_cycle=1, n_workers=4)
    best = opt.run(paper_objective, init_samples=np.array([[-1.0, 0.0], [1.0, 0.0], [-1.5, 1], [1.5, 1]]), eval_repeats=4, seed=42)
    assert best.score_mu >= 6.5


@pytest.mark.timeout(600)
def test_reference_v3_saver_and_repeats(cuda):
    """Reference tests/test_optimisation.py:183-221 on the GPU (TableSaver replaced by a stub with the same call)."""
    saver = StubSaver()
    opt = GPSOptimiser(parameter_space=space_2d(), exploration_method="tree", exploration_depth=3, budget=50,
                       stopping_condition="evaluations", update_cycle=1, n_workers=4, saver=saver)
    best = opt.run(objective_with_result, eval_repeats=4)
    np.testing.assert_almost_equal(np.array([0.23525377, 0.68518519]), best.normed_coord)
    assert np.around(best.score_mu, decimals=8) == 8.10560594
    assert len(saver.calls) == opt.n_eval_counter and all(len(c[0]) == 4 for c in saver.calls)


# ---------------------------------------------------------------------------------------------------------------------
# the sharding axes behind the kept API: two ranks (gloo for the collectives, both sessions on this
# GPU) replay GPSOptimiser.run with GPRSurrogate(group=True); candidates / leaf batches are sharded over the ranks, the
# fitted state is broadcast once per fit.  CUDA results do not depend on the shard a candidate arrives in, so the SPMD run
# must equal the single-process run BIT FOR BIT.
# ---------------------------------------------------------------------------------------------------------------------
def _spmd_gpu_run(group, n_restarts):
    def rastrigin(point):
        x = np.asarray(point)
        return -float(10 * x.size + np.sum(x * x - 10 * np.cos(2 * np.pi * x)))

    space = ParameterSpace(parameter_names=[f"p{i}" for i in range(4)], parameter_bounds=[[-5.12, 5.12]] * 4)
    surr = GPRSurrogate.default(group=group, n_restarts=n_restarts, restart_maxiter=15)
    opt = GPSOptimiser(parameter_space=space, gp_surrogate=surr, exploration_method="tree", exploration_depth=8, budget=40,
                       stopping_condition="evaluations", update_cycle=1, n_workers=1)
    best = opt.run(rastrigin)
    return {"coord": np.asarray(best.normed_coord), "score": best.score_mu, "iterations": opt.iterations, "evals": opt.n_eval_counter,
            "theta": surr.gpflow_model._theta(), "ucbs": np.array([p.score_ucb for p in surr.points]),
            "mus": np.array([p.score_mu for p in surr.points])}


def _spmd_gpu_worker(rank, world, port, out_dir, n_restarts):
    import os
    import pickle
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["GPSO_DEVICE"] = "0"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = _spmd_gpu_run(True, n_restarts)
        with open(os.path.join(out_dir, f"spmd{rank}.pkl"), "wb") as fh:
            pickle.dump(out, fh)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(900)
@pytest.mark.parametrize("n_restarts", [1, 3])
def test_spmd_two_ranks_equal_single_process_bit_for_bit(cuda, tmp_path, n_restarts):
    import pickle

    import torch.multiprocessing as mp

    from tests.test_spmd_gloo import _free_port

    mp.spawn(_spmd_gpu_worker, args=(2, _free_port(), str(tmp_path), n_restarts), nprocs=2, join=True)
    r0, r1 = (pickle.load(open(tmp_path / f"spmd{r}.pkl", "rb")) for r in range(2))
    single = _spmd_gpu_run(None, n_restarts)
    for other in (r1, single):
        assert r0["iterations"] == other["iterations"] and r0["evals"] == other["evals"]
        for key in ("coord", "theta", "ucbs", "mus"):
            assert np.array_equal(r0[key], other[key]), key
        assert r0["score"] == other["score"]


def test_post_iteration_plotting_on_gpu(cuda, tmp_path):
    """N4: the model-reading part of ``PostIterationPlotting`` (conditional surrogate slices, one batched predict_y) on the GPU
    against the oracle at the final hyper-parameters."""
    import os

    from pygpso_b200.callbacks import PostIterationPlotting

    space = ParameterSpace(parameter_names=["a", "b", "c"], parameter_bounds=[[-1, 1], [0, 2], [-3, 3]])
    pattern = str(tmp_path / "plots" / "run")
    opt = GPSOptimiser(parameter_space=space, exploration_depth=4, budget=20,
                       callbacks=[PostIterationPlotting(pattern, from_iteration=1, granularity=9)])
    opt.run(lambda p: -float(np.sum((np.asarray(p) - 0.3) ** 2)))
    files = sorted(f for f in os.listdir(tmp_path / "plots") if f.endswith(".npz"))
    data = np.load(tmp_path / "plots" / files[-1])
    model = opt.gp_surr.gpflow_model
    X, y = model.data
    theta = model._theta()
    h = go.Hyper(theta[0], theta[1], theta[2], theta[3])
    g = 9
    best = np.vstack([opt.gp_surr.highest_score.normed_coord] * g ** 2)
    gx, gy = np.meshgrid(np.linspace(0, 1, g), np.linspace(0, 1, g))
    for (i, j) in ((0, 1), (0, 2), (1, 2)):
        at = best.copy()
        at[:, i], at[:, j] = gx.flatten(), gy.flatten()
        # the reference's call shape: one predict_y per pair, .numpy().reshape(grid) -- bit-identical to the batched call
        # (a candidate's result does not depend on the batch it arrives in)
        m_pair, v_pair = model.predict_y(at)
        assert np.array_equal(data[f"mean_{i}_{j}"], m_pair.numpy().reshape(gx.shape))
        assert np.array_equal(data[f"var_{i}_{j}"], v_pair.numpy().reshape(gx.shape))
        # ... and the oracle at the fitted hyper-parameters (a fitted noise variance near the 1e-6 floor makes K_y
        # ill-conditioned: the comparison carries the condition number, as in test_gpu_parity's fitted-theta cases)
        mean, var = go.predict_y("Matern52", X, y, h, at)
        np.testing.assert_allclose(data[f"mean_{i}_{j}"].reshape(-1), mean[:, 0], rtol=1e-5, atol=1e-6 * np.abs(y).max())
        np.testing.assert_allclose(data[f"var_{i}_{j}"].reshape(-1), var[:, 0], rtol=1e-4, atol=1e-6 * h.variance)
