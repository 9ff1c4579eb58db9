"""
Pin the oracle (oracle/gpr_oracle.py) and the host-side loop against every golden number the reference ships
(SURVEY.md section 8c).  CPU only: the optimiser runs with the checker backend injected explicitly.
"""
import json
import os

import numpy as np
import pytest

from oracle import gpr_oracle as go
from pygpso_b200 import GPRSurrogate, GPSOptimiser, ParameterSpace, PointLabels
from pygpso_b200 import gpmodel
from pygpso_b200.gp_surrogate import GPPoint
from pygpso_b200.optimisation import CallbackTypes, GPSOCallback
from tests.conftest import paper_objective

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
KATS = json.load(open(os.path.join(GOLDEN, "reference_kats.json")))
TRACE = json.load(open(os.path.join(GOLDEN, "notebook_trace_depth5.json")))


def seeded_points(n_points=10, seed=42):
    """The reference's seeded fixture (tests/test_gp_surrogate.py:221-237)."""
    points = []
    for i in range(n_points):
        np.random.seed(seed + i)
        points.append(
            GPPoint(
                normed_coord=np.random.rand(2),
                score_mu=np.random.rand(),
                score_sigma=np.random.rand(),
                score_ucb=np.random.rand(),
                label=PointLabels(np.random.choice([1, 2], p=[0.8, 0.2])),
            )
        )
    return points


def test_seeded_fixture_properties(oracle_backend):
    fx = KATS["surrogate_fixture"]
    surr = GPRSurrogate(gp_kernel=gpmodel.Matern52(), gp_meanf=gpmodel.Constant(), points=seeded_points(), backend=oracle_backend)
    assert len(surr.points) == fx["num_points"]
    assert surr.num_evaluated == fx["num_evaluated"]
    assert surr.num_gp_based == fx["num_points"] - fx["num_evaluated"]
    best = surr.highest_score
    assert best.label == PointLabels.evaluated
    assert best.score_mu == fx["highest_score"]
    np.testing.assert_almost_equal(best.normed_coord, fx["highest_coords"])
    x, y = surr.current_training_data
    assert x.shape == (6, 2) and y.shape == (6,)
    assert surr.gp_based_coords.shape == (4, 2)


def test_oracle_fit_predict_kat():
    """tests/test_gp_surrogate.py:259-268 straight on the oracle model."""
    kat = KATS["fit_predict"]
    pts = [p for p in seeded_points() if p.label == PointLabels.evaluated]
    X = np.array([p.normed_coord for p in pts])
    y = np.array([p.score_mu for p in pts])[:, None]
    model = go.OracleGPR(X, y, "Matern52", lengthscales=1.0, variance=1.0, noise_variance=1e-3, mean_c=0.0)
    res = model.fit()
    assert res.success
    mean, var = model.predict_y(np.array(kat["predict_at"]))
    assert float(np.around(mean, 8)[0, 0]) == kat["mean_8dp"]
    assert float(np.around(var, 8)[0, 0]) == kat["var_8dp"]
    # the floor of the Gaussian likelihood variance and the softplus parameterisation
    assert model.h.noise_variance > go.NOISE_FLOOR
    assert model.log_marginal_likelihood() == pytest.approx(-res.fun, rel=1e-12)


def test_surrogate_fit_predict_ucb_kat(oracle_backend):
    """The same KATs through GPRSurrogate (train, gp_predict, gp_eval_best_ucb): test_gp_surrogate.py:259-309."""
    kat = KATS["fit_predict"]
    surr = GPRSurrogate(gp_kernel=gpmodel.Matern52(), gp_meanf=gpmodel.Constant(), points=seeded_points(), backend=oracle_backend)
    x, y = surr.current_training_data
    surr._gp_train(x=x, y=y[:, np.newaxis])
    mean, var = surr.gpflow_model.predict_y(np.array(kat["predict_at"]))
    assert mean.shape == (1, 1) and var.shape == (1, 1)
    assert float(np.around(mean.numpy(), 8)[0, 0]) == kat["mean_8dp"]
    assert float(np.around(var.numpy(), 8)[0, 0]) == kat["var_8dp"]
    # gp_predict appends exactly one GP-based point
    surr.gp_predict(np.array(kat["predict_at"]))
    assert len(surr.points) == 11
    new = surr.points[-1]
    np.testing.assert_equal(new.normed_coord, np.array(kat["predict_at"])[0])
    assert float(np.around(new.score_mu, 8)) == kat["mean_8dp"]
    assert float(np.around(new.score_sigma, 8)) == kat["var_8dp"]
    assert new.label == PointLabels.gp_based
    # UCB uses the variance (not the standard deviation)
    best = surr.gp_eval_best_ucb(np.array(kat["ucb_candidates"]))
    exp_ucb = np.around(kat["mean_8dp"] + surr.gp_varsigma * kat["var_8dp"], 8)
    assert float(np.around(best[0], 8)) == kat["mean_8dp"]
    assert float(np.around(best[1], 8)) == kat["var_8dp"]
    assert float(np.around(best[2], 8)) == exp_ucb
    assert len(surr.points) == 11
    assert surr.gp_varsigma == pytest.approx(1.8213863677184496, rel=1e-15)


def make_optimiser(backend, depth, budget, callbacks=None):
    space = ParameterSpace(parameter_names=["x", "y"], parameter_bounds=[[-3, 5], [-3, 3]])
    return GPSOptimiser(
        parameter_space=space,
        gp_surrogate=GPRSurrogate.default(backend=backend),
        exploration_method="tree",
        exploration_depth=depth,
        budget=budget,
        stopping_condition="evaluations",
        update_cycle=1,
        n_workers=1,
        callbacks=callbacks,
    )


def test_end_to_end_depth3(oracle_backend):
    """tests/test_optimisation.py:66-89."""
    kat = KATS["end_to_end_depth3"]
    opt = make_optimiser(oracle_backend, depth=3, budget=50)
    best = opt.run(paper_objective)
    np.testing.assert_almost_equal(np.array(kat["best_coords_7dp"]), best.normed_coord)
    assert np.around(best.score_mu, decimals=8) == kat["best_score_8dp"]
    assert opt.iterations == 13 and opt.n_eval_counter == 55


def test_end_to_end_resume(oracle_backend):
    """tests/test_optimisation.py:91-117: 25 + 25 evaluations reach the same optimum."""
    kat = KATS["end_to_end_depth3"]
    opt = make_optimiser(oracle_backend, depth=3, budget=25)
    opt.run(paper_objective)
    best = opt.resume_run(additional_budget=25)
    np.testing.assert_almost_equal(np.array(kat["best_coords_7dp"]), best.normed_coord)
    assert np.around(best.score_mu, decimals=8) == kat["best_score_8dp"]


def test_end_to_end_save_load_resume(oracle_backend, tmp_path):
    """tests/test_optimisation.py:119-152: save -> load -> resume gives the same optimum."""
    kat = KATS["end_to_end_depth3"]
    opt = make_optimiser(oracle_backend, depth=3, budget=25)
    opt.run(paper_objective)
    folder = str(tmp_path / "state")
    opt.save_state(folder)
    for name in ("parameter_space.pkl", "points.json", "GPRinfo.json", "GPRmodel.pkl", "opt_attributes.json"):
        assert os.path.exists(os.path.join(folder, name))
    best, resumed = GPSOptimiser.resume_from_saved(folder, additional_budget=25, objective_function=paper_objective,
                                                   backend=oracle_backend)
    np.testing.assert_almost_equal(np.array(kat["best_coords_7dp"]), best.normed_coord)
    assert np.around(best.score_mu, decimals=8) == kat["best_score_8dp"]
    assert isinstance(resumed, GPSOptimiser)


class _Recorder(GPSOCallback):
    def __init__(self, kind, sink):
        self.callback_type = kind
        super().__init__()
        self.sink = sink

    def run(self, optimiser):
        surr = optimiser.gp_surr
        if self.callback_type == CallbackTypes.post_iteration:
            self.sink.append((optimiser.n_eval_counter, surr.highest_score.score_mu, surr.highest_ucb.score_ucb))
        else:
            m = surr.gpflow_model
            self.sink.append({
                "mean_function.c": float(m.mean_function.c), "kernel.variance": float(m.kernel.variance),
                "kernel.lengthscales": float(m.kernel.lengthscales), "likelihood.variance": float(m.likelihood.variance),
            })


def test_notebook_trace_depth5(oracle_backend):
    """examples/0-basic-optimisation.ipynb + examples/1-callbacks.ipynb: 13 iterations and 14 fits of the depth-5 run."""
    iters, fits = [], []
    opt = make_optimiser(oracle_backend, depth=5, budget=50,
                         callbacks=[_Recorder(CallbackTypes.post_iteration, iters), _Recorder(CallbackTypes.post_update, fits)])
    best = opt.run(paper_objective)
    assert len(iters) == len(TRACE["iterations"]) == 13
    for got, want in zip(iters, TRACE["iterations"]):
        assert got[0] == want["evaluations"]
        assert got[1] == want["highest_score"]  # objective values: exact
        assert got[2] == pytest.approx(want["highest_ucb"], abs=5e-10)
    np.testing.assert_almost_equal(best.normed_coord, TRACE["best_point"]["normed_coord"])
    assert best.score_mu == TRACE["best_point"]["score_mu"]
    assert len(fits) == len(TRACE["hyperparameters_per_update"]) == 14
    for got, want in zip(fits, TRACE["hyperparameters_per_update"]):
        for key, value in want.items():
            # the notebook prints 6 significant digits: agree to within 0.6 units of the last printed digit
            last_digit = 10.0 ** (np.floor(np.log10(abs(value))) - 5)
            assert abs(got[key] - value) <= 0.6 * last_digit, (key, got[key], value)


def test_gradient_matches_finite_differences():
    rng = np.random.default_rng(3)
    X = rng.random((25, 3))
    y = np.sin(3 * X.sum(1))[:, None]
    for kernel, ls in (("Matern52", 0.4), ("Matern32", [0.3, 0.5, 0.7]), ("SquaredExponential", 0.6), ("Matern12", [0.5, 0.5, 0.9])):
        model = go.OracleGPR(X, y, kernel, lengthscales=ls, variance=1.3, noise_variance=1e-2, mean_c=0.1)
        u = model.h.pack()
        f, g = model.training_loss(u)
        # Matern12 is not differentiable at r = 0 and GPflow's |x|^2+|x'|^2-2x.x' distance leaves rounding noise on the
        # diagonal, so its finite differences are only meaningful with a coarse step
        step, rel = (1e-3, 2e-3) if kernel == "Matern12" else (1e-4, 2e-6)
        for k in range(u.size):
            e = np.zeros_like(u)
            e[k] = step
            fd = (model.training_loss(u + e)[0] - model.training_loss(u - e)[0]) / (2 * step)
            assert g[k] == pytest.approx(fd, rel=rel, abs=1e-8), (kernel, k)


def test_predict_matches_extended_precision():
    rng = np.random.default_rng(5)
    X = rng.random((40, 2))
    y = np.sin(3 * X.sum(1))[:, None]
    h = go.Hyper(0.3, 1.5, 1e-4, 0.2)
    Xn = rng.random((30, 2))
    mean, var = go.predict_y("Matern52", X, y, h, Xn)
    mean_ld, var_ld = go.predict_y_longdouble("Matern52", X, y, h, Xn)
    np.testing.assert_allclose(mean, mean_ld.astype(float), rtol=0, atol=1e-9)
    np.testing.assert_allclose(var, var_ld.astype(float), rtol=0, atol=1e-9)
