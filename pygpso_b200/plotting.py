"""
Data side of the reference's surrogate plots (``gpso/plotting.py:257-494``, used by the ``PostIterationPlotting`` callback,
``gpso/callbacks.py:19-87``): the conditional surrogate distributions -- posterior mean and variance on orthogonal 2-D slices
of the normalised search space through the best evaluated point, one slice per pair of parameters.

The reference calls ``gpflow_model.predict_y`` once per pair inside its matplotlib loop and converts with ``.numpy()``.  Here
all pairs are predicted in ONE batched ``predict_y`` call (ndim (ndim-1)/2 * granularity^2 rows through the windowed device
pipeline) and returned as arrays; drawing them is left to the caller.  matplotlib is not a dependency of this package:
``render_conditional_surrogate`` draws the same N x N grid of panels when it is importable and is a no-op otherwise.
"""
import itertools
import logging

import numpy as np

from .utils import PointLabels

N_BINS = 10  # the reference's default granularity is N_BINS ** 2 points per axis


def conditional_surrogate_slices(gpso_optimiser, granularity=N_BINS ** 2):
    """
    ``{(i, j): (mean[g, g], var[g, g])}`` for every parameter pair i < j: the posterior over the grid
    ``linspace(0, 1, g) x linspace(0, 1, g)`` in dimensions (i, j), all other coordinates fixed at the best evaluated point.
    ``mean[a, b]`` belongs to coordinate i = grid[b], coordinate j = grid[a] (``np.meshgrid`` order, what the reference
    reshapes its flat predictions to).
    """
    surr = gpso_optimiser.gp_surr
    best = np.asarray(surr.highest_score.normed_coord, dtype=np.float64)
    ndim = best.size
    axis = np.linspace(0.0, 1.0, granularity)
    gx, gy = np.meshgrid(axis, axis)
    gx, gy = gx.reshape(-1), gy.reshape(-1)
    pairs = list(itertools.combinations(range(ndim), 2))
    if not pairs:
        return {}
    batch = np.tile(best, (len(pairs) * gx.size, 1))
    for k, (i, j) in enumerate(pairs):
        rows = slice(k * gx.size, (k + 1) * gx.size)
        batch[rows, i] = gx
        batch[rows, j] = gy
    mean, var = surr.gpflow_model.predict_y(batch)
    mean = mean.numpy().reshape(len(pairs), granularity, granularity)
    var = var.numpy().reshape(len(pairs), granularity, granularity)
    return {pair: (mean[k], var[k]) for k, pair in enumerate(pairs)}


def evaluated_scores_by_parameter(gpso_optimiser):
    """(coords[P, ndim], scores[P]) of the evaluated points: what the diagonal panels and the marginal plots are drawn from."""
    points = [p for p in gpso_optimiser.gp_surr.points if p.label == PointLabels.evaluated]
    return np.array([p.normed_coord for p in points]), np.array([p.score_mu for p in points])


def render_conditional_surrogate(gpso_optimiser, slices, mean_limits=(-10, 10), var_limits=(0, 5), fname=None):
    """Draw the N x N panel grid (mean below, variance above the diagonal) when matplotlib is available; returns True if drawn."""
    try:
        import matplotlib

        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except ImportError:
        logging.info("matplotlib is not installed: conditional-surrogate arrays computed, figure not rendered")
        return False
    ndim = gpso_optimiser.param_space.ndim
    names = gpso_optimiser.param_space.parameter_names
    fig, axes = plt.subplots(nrows=ndim, ncols=ndim, squeeze=False)
    for (i, j), (mean, var) in slices.items():
        axes[j, i].imshow(mean, vmin=mean_limits[0], vmax=mean_limits[1], cmap="Spectral", origin="lower")
        axes[i, j].imshow(var.T, vmin=var_limits[0], vmax=var_limits[1], cmap="plasma", origin="lower")
    coords, scores = evaluated_scores_by_parameter(gpso_optimiser)
    for k in range(ndim):
        axes[k, k].scatter(coords[:, k], scores, s=4)
        axes[0, k].set_title(names[k])
        axes[k, 0].set_ylabel(names[k])
    if fname is not None:
        fig.savefig(fname)
    plt.close(fig)
    return True
