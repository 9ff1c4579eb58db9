"""
Callbacks that read the surrogate model (reference ``gpso/callbacks.py``): the logging, save-before-finalise and
checkpoint callbacks, and ``PostIterationPlotting`` (``callbacks.py:19-87``) with its model-reading part -- the conditional
surrogate distributions, one batched ``predict_y`` -- computed and saved as arrays; figures are drawn only when matplotlib is
importable (it is not a dependency).

Same class names, constructor arguments and callback types as the reference, so a callback list written for pyGPSO runs
unchanged.  ``GPFlowCheckpoints`` kept its name; instead of a TensorFlow checkpoint manager it writes the model's parameter
dictionary (the same dictionary ``GPRSurrogate.save`` pickles) plus the evaluation counter, rotating ``max_to_keep`` files.
"""
import glob
import logging
import os

import dill
import numpy as np

from . import gpmodel
from .optimisation import CallbackTypes, GPSOCallback
from .utils import PKL_EXT, make_dirs


class PostIterationPlotting(GPSOCallback):
    """
    After every iteration: conditional surrogate distributions (posterior mean / variance on 2-D slices through the best
    point, reference ``plotting.py:257-494``) and the evaluated scores per parameter.  Constructor arguments as in the
    reference (``callbacks.py:27-64``).  Writes ``<pattern>_iter<k>_surrogate_dist.npz`` with one ``mean_i_j`` / ``var_i_j``
    array per parameter pair, and ``<pattern>_iter<k>_surrogate_dist<plot_ext>`` when matplotlib is available.
    """

    callback_type = CallbackTypes.post_iteration

    def __init__(self, filename_pattern, plot_ext=".png", gp_mean_limits=[-10, 10], gp_var_limits=[0, 5], marginal_plot_type="kde",
                 marginal_percentile=0.9, from_iteration=1, granularity=None):
        super().__init__()
        self.filename_pattern = filename_pattern
        self.plot_ext = plot_ext
        self.gp_mean_limits = gp_mean_limits
        self.gp_var_limits = gp_var_limits
        self.marginal_plot_type = marginal_plot_type
        self.marginal_percentile = marginal_percentile
        self.from_iteration = from_iteration
        self.granularity = granularity

    def run(self, optimiser):
        super().run(optimiser)
        if optimiser.iterations < self.from_iteration:
            return
        from . import plotting

        stem = self.filename_pattern + f"_iter{optimiser.iterations}"
        make_dirs(os.path.dirname(os.path.abspath(stem)))
        kwargs = {} if self.granularity is None else {"granularity": self.granularity}
        slices = plotting.conditional_surrogate_slices(optimiser, **kwargs)
        coords, scores = plotting.evaluated_scores_by_parameter(optimiser)
        arrays = {"evaluated_coords": coords, "evaluated_scores": scores}
        for (i, j), (mean, var) in slices.items():
            arrays[f"mean_{i}_{j}"] = mean
            arrays[f"var_{i}_{j}"] = var
        np.savez(stem + "_surrogate_dist.npz", **arrays)
        plotting.render_conditional_surrogate(optimiser, slices, self.gp_mean_limits, self.gp_var_limits,
                                              fname=stem + f"_surrogate_dist{self.plot_ext}")


class PostUpdateLogging(GPSOCallback):
    """Log the GPR summary after every update (reference ``callbacks.py:90-102``)."""

    callback_type = CallbackTypes.post_update

    def run(self, optimiser):
        super().run(optimiser)
        logging.info("GPR summary:\n" + gpmodel.tabulate_module_summary(optimiser.gp_surr.gpflow_model))


class PreFinaliseSave(GPSOCallback):
    """Save parameter space and surrogate right before the run finishes (reference ``callbacks.py:105-121``)."""

    callback_type = CallbackTypes.pre_finalise

    def __init__(self, path):
        super().__init__()
        self.path = path

    def run(self, optimiser):
        super().run(optimiser)
        make_dirs(self.path)
        optimiser.param_space.save(os.path.join(self.path, f"parameter_space{PKL_EXT}"))
        optimiser.gp_surr.save(self.path)


class GPFlowCheckpoints(GPSOCallback):
    """
    Checkpoint the GP hyper-parameters after every update, keeping the newest ``max_to_keep`` (reference
    ``callbacks.py:124-155``).  Files: ``<path>/ckpt-<k>.pkl`` holding ``{"parameters": {...}, "evaluations": n}``.
    """

    callback_type = CallbackTypes.post_update
    PATTERN = "ckpt-{:d}" + PKL_EXT

    def __init__(self, path, max_to_keep=10):
        super().__init__()
        self.path = path
        self.max_to_keep = max_to_keep
        self.saved = 0

    def run(self, optimiser):
        super().run(optimiser)
        make_dirs(self.path)
        model = optimiser.gp_surr.gpflow_model
        state = {
            "parameters": {key: np.asarray(p).copy() for key, p in gpmodel.parameter_dict(model).items()},
            "evaluations": int(optimiser.n_eval_counter),
        }
        self.saved += 1
        filename = os.path.join(self.path, self.PATTERN.format(self.saved))
        with open(filename, "wb") as handle:
            dill.dump(state, handle)
        logging.info(f"Saved checkpoint for step {self.saved}: {filename}")
        if self.max_to_keep:
            for old in range(self.saved - self.max_to_keep, 0, -1):
                stale = os.path.join(self.path, self.PATTERN.format(old))
                if not os.path.exists(stale):
                    break
                os.remove(stale)

    @classmethod
    def latest(cls, path):
        """(state dict, filename) of the newest checkpoint in ``path``, or (None, None)."""
        files = glob.glob(os.path.join(path, "ckpt-*" + PKL_EXT))
        if not files:
            return None, None
        newest = max(files, key=lambda f: int(os.path.basename(f)[len("ckpt-"):-len(PKL_EXT)]))
        with open(newest, "rb") as handle:
            return dill.load(handle), newest
