"""
Worker pool, evaluation repeats and saver of ``GPSOptimiser.evaluate_objective_function`` (reference
``gpso/optimisation.py:464-535``, tests ``tests/test_optimisation.py:154-221``), on the CPU with the oracle backend.
"""
import pickle

import numpy as np
import pytest

from pygpso_b200 import GPRSurrogate, GPSOptimiser, ParameterSpace
from tests.conftest import paper_objective


def space_2d():
    return ParameterSpace(parameter_names=["x", "y"], parameter_bounds=[[-3, 5], [-3, 3]])


def objective_with_result(point):
    """(result, score) pair as the saver protocol expects (reference ``_obj_func_w_return``)."""
    score = paper_objective(point)
    return np.array([score, 2.0 * score]), score


class StubSaver:
    """Duck type of ``gpso.saving_helper.TableSaver`` as the optimiser uses it (``save_runs(results, scores, params)``)."""

    def __init__(self):
        self.calls = []

    def save_runs(self, results, scores, parameters):
        self.calls.append((list(results), list(scores), dict(parameters)))

    def close(self):
        pass


def test_v2_sample_method_four_workers_four_repeats(oracle_backend):
    opt = GPSOptimiser(parameter_space=space_2d(), gp_surrogate=GPRSurrogate.default(backend=oracle_backend),
                       exploration_method="sample", exploration_depth=3, budget=12, stopping_condition="iterations",
                       update_cycle=1, n_workers=4)
    best = opt.run(paper_objective, init_samples=np.array([[-1.0, 0.0], [1.0, 0.0], [-1.5, 1], [1.5, 1]]), eval_repeats=4, seed=42)
    assert best.score_mu >= 6.5
    assert opt.iterations == 12


@pytest.mark.parametrize("n_workers", [1, 4])
def test_v3_saver_and_repeats(oracle_backend, n_workers):
    saver = StubSaver()
    opt = GPSOptimiser(parameter_space=space_2d(), gp_surrogate=GPRSurrogate.default(backend=oracle_backend),
                       exploration_method="tree", exploration_depth=3, budget=50, stopping_condition="evaluations",
                       update_cycle=1, n_workers=n_workers, saver=saver)
    best = opt.run(objective_with_result, eval_repeats=4)
    np.testing.assert_almost_equal(np.array([0.23525377, 0.68518519]), best.normed_coord)
    assert np.around(best.score_mu, decimals=8) == 8.10560594
    assert len(saver.calls) == opt.n_eval_counter == 55
    assert all(len(results) == 4 and len(scores) == 4 for results, scores, _ in saver.calls)
    assert set(saver.calls[0][2]) == {"x", "y"}


def test_pool_matches_serial(oracle_backend):
    from tests.oracle_backend import OracleBackend

    kw = dict(exploration_method="tree", exploration_depth=4, budget=30, stopping_condition="evaluations")
    serial = GPSOptimiser(parameter_space=space_2d(), gp_surrogate=GPRSurrogate.default(backend=oracle_backend), n_workers=1, **kw)
    pooled = GPSOptimiser(parameter_space=space_2d(), gp_surrogate=GPRSurrogate.default(backend=OracleBackend()), n_workers=4, **kw)
    a, b = serial.run(paper_objective), pooled.run(paper_objective)
    assert np.array_equal(a.normed_coord, b.normed_coord) and a.score_mu == b.score_mu


def test_point_list_and_tree_pickle(oracle_backend):
    """The reference's list subclass pickles; the tree must not drag the point list along."""
    opt = GPSOptimiser(parameter_space=space_2d(), gp_surrogate=GPRSurrogate.default(backend=oracle_backend),
                       exploration_method="tree", exploration_depth=3, budget=12, stopping_condition="evaluations")
    opt.run(paper_objective)
    points = opt.gp_surr.points
    clone = pickle.loads(pickle.dumps(points))
    assert type(clone) is type(points) and len(clone) == len(points)
    assert all(a == b for a, b in zip(clone, points))
    assert clone.index_by_coords(points[3].normed_coord) == 3
    clone.append(points[0])  # duplicate of an evaluated point: ignored
    assert len(clone) == len(points)
    tree_bytes = pickle.dumps(opt.param_space)
    assert len(tree_bytes) < 20 * len(pickle.dumps(points)) and b"GPListOfPoints" not in tree_bytes
