"""Model-reading callbacks (reference gpso/callbacks.py:90-155) on the checker backend -- CPU only."""
import logging
import os

import numpy as np

from pygpso_b200 import GPRSurrogate, GPSOptimiser, ParameterSpace
from pygpso_b200.callbacks import GPFlowCheckpoints, PostUpdateLogging, PreFinaliseSave
from tests.conftest import paper_objective
from tests.test_oracle_goldens import make_optimiser


def test_callbacks_run_and_persist(oracle_backend, tmp_path, caplog):
    ckpt_dir, save_dir = str(tmp_path / "ckpt"), str(tmp_path / "final")
    callbacks = [PostUpdateLogging(), GPFlowCheckpoints(ckpt_dir, max_to_keep=3), PreFinaliseSave(save_dir)]
    opt = make_optimiser(oracle_backend, depth=3, budget=20, callbacks=callbacks)
    with caplog.at_level(logging.INFO):
        best = opt.run(paper_objective)
    assert "GPR summary:" in caplog.text and ".kernel.lengthscales" in caplog.text
    # checkpoints: only the newest three are kept, the newest carries the final hyper-parameters and evaluation count
    files = sorted(os.listdir(ckpt_dir))
    assert len(files) == 3
    state, newest = GPFlowCheckpoints.latest(ckpt_dir)
    assert newest.endswith(f"ckpt-{callbacks[1].saved}.pkl") and state["evaluations"] == opt.n_eval_counter
    model = opt.gp_surr.gpflow_model
    assert np.array_equal(state["parameters"][".kernel.lengthscales"], np.asarray(model.kernel.lengthscales))
    # the pre-finalise save is loadable and predicts the same
    assert os.path.exists(os.path.join(save_dir, "parameter_space.pkl"))
    loaded = GPRSurrogate.from_saved(save_dir, backend=oracle_backend)
    x = np.array([[0.3, 0.7], [0.5, 0.5]])
    for a, b in zip(loaded.predict_y(x), opt.gp_surr.predict_y(x)):
        np.testing.assert_array_equal(np.asarray(a), np.asarray(b))
    assert loaded.highest_score.score_mu == best.score_mu


def test_post_iteration_plotting_conditional_surrogate(oracle_backend, tmp_path):
    """``PostIterationPlotting`` (reference callbacks.py:19-87) and the call shape of plotting.py:346-356: per parameter pair
    ``gpflow_model.predict_y(grid)`` -> ``.numpy().reshape(meshgrid shape)``; here one batched call for all pairs."""
    from pygpso_b200 import plotting
    from pygpso_b200.callbacks import PostIterationPlotting

    space = ParameterSpace(parameter_names=["a", "b", "c"], parameter_bounds=[[-1, 1], [0, 2], [-3, 3]])
    pattern = str(tmp_path / "plots" / "run")
    opt = GPSOptimiser(parameter_space=space, gp_surrogate=GPRSurrogate.default(backend=oracle_backend), exploration_depth=3, budget=14,
                       callbacks=[PostIterationPlotting(pattern, from_iteration=2, granularity=7)])
    opt.run(lambda p: -float(np.sum((np.asarray(p) - 0.3) ** 2)))
    files = sorted(os.listdir(tmp_path / "plots"))
    assert files and all("_iter" in f for f in files) and not any("_iter1_" in f for f in files)
    data = np.load(tmp_path / "plots" / [f for f in files if f.endswith(".npz")][-1])
    assert {"mean_0_1", "var_0_1", "mean_0_2", "mean_1_2", "evaluated_scores"} <= set(data.files)
    # the last file was written at the final model state: compare with the reference's per-pair evaluation
    g = 7
    best = np.vstack([opt.gp_surr.highest_score.normed_coord] * g ** 2)
    x, y = np.meshgrid(np.linspace(0, 1, g), np.linspace(0, 1, g))
    for (i, j) in ((0, 1), (0, 2), (1, 2)):
        predict_at = best.copy()
        predict_at[:, i] = x.flatten()
        predict_at[:, j] = y.flatten()
        mean, var = opt.gp_surr.gpflow_model.predict_y(predict_at)
        np.testing.assert_allclose(data[f"mean_{i}_{j}"], mean.numpy().reshape(x.shape), rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(data[f"var_{i}_{j}"], var.numpy().reshape(x.shape), rtol=1e-10, atol=1e-12)
    slices = plotting.conditional_surrogate_slices(opt, granularity=5)
    assert set(slices) == {(0, 1), (0, 2), (1, 2)} and slices[(0, 1)][0].shape == (5, 5)
