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
