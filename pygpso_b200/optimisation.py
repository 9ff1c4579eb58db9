"""
Gaussian-process surrogate optimisation loop (explore -> select -> update) -- the reference's
``gpso/optimisation.py`` with an unchanged public API: ``GPSOptimiser`` (``run`` :538, ``resume_run`` :634,
``resume_from_saved`` :76, ``save_state`` :697, ``evaluate_objective_function`` :464), ``GPSOCallback`` and
``CallbackTypes`` (:19-51).

The control flow is host Python, as in the reference.  What moved to the GPU:
  * the exploitation step of ``_tree_explore`` (:342-403): for each new child leaf the batch of ``grow(depth)`` leaf
    centres is generated on the device, scored with the fused predict_y + UCB kernel and reduced to its arg-max in a
    single call (``GPSurrogate.gp_eval_best_ucb_in_leaf``); the "sample" exploration method scores its random batch
    with the same fused kernel (``gp_eval_best_ucb``);
  * the update step (``_gp_update`` :314-340): hyper-parameter fit and re-prediction of all GP-based centres.

Objective evaluations run in a ``multiprocess`` pool (dill-based, stands in for pathos) when ``n_workers > 1``.
"""
import json
import logging
import os
from enum import Enum, auto, unique
from functools import partial

import numpy as np

from .gp_surrogate import GPPoint, GPRSurrogate, GPSurrogate, PointLabels
from .param_space import NORM_PARAMS_BOUNDS, ParameterSpace, PreOrderIter
from .utils import JSON_EXT, PKL_EXT, load_json, make_dirs


@unique
class CallbackTypes(Enum):
    post_initialise = auto()
    pre_iteration = auto()
    post_iteration = auto()
    post_update = auto()
    pre_finalise = auto()


class GPSOCallback:
    """Base class of user callbacks; subclasses set ``callback_type`` and implement ``run(optimiser)``."""

    callback_type = None

    def __init__(self):
        assert self.callback_type in CallbackTypes, "Callback type must be one of `CallbackTypes`"

    def run(self, optimiser):
        assert isinstance(optimiser, GPSOptimiser)
        logging.info(f"Running {self.__class__.__name__} callback...")


def _pool_start_method():
    """``fork`` while this process has not touched CUDA (the usual case: the pool of a run is created by the initial design,
    before the first fit opens the device session), else ``forkserver`` -- a forked child of a CUDA-initialised,
    multi-threaded parent can inherit locked driver state.  ``GPSO_POOL_START`` overrides."""
    forced = os.environ.get("GPSO_POOL_START")
    if forced:
        return forced
    from . import backend

    return "forkserver" if backend.cuda_initialised() else "fork"


def _make_pool(n_workers):
    """Process pool for objective evaluations (children never touch CUDA)."""
    try:
        import multiprocess as mp  # dill-based: handles lambdas / local functions like pathos does
    except ImportError:  # pragma: no cover
        import multiprocessing as mp
    return mp.get_context(_pool_start_method()).Pool(n_workers)


class GPSOptimiser:
    """Bayesian optimisation over a ternary partition tree with a GP surrogate."""

    SAVE_ATTRS = [
        "iterations",
        "budget",
        "eval_repeats",
        "last_explored_levels",
        "last_update_idx",
        "method",
        "max_depth",
        "stop_cond",
        "update_cycle",
        "n_eval_counter",
        "n_workers",
        "expl_seed",
    ]
    PARAM_SPACE_FILE = f"parameter_space{PKL_EXT}"
    OPT_ATTRS_FILE = f"opt_attributes{JSON_EXT}"

    @classmethod
    def resume_from_saved(
        cls,
        folder,
        additional_budget,
        objective_function,
        gp_surrogate=GPRSurrogate,
        eval_repeats_function=np.mean,
        callbacks=None,
        saver=None,
        **surrogate_kwargs,
    ):
        """
        Load a saved optimiser state from ``folder`` and continue for ``additional_budget``.  Returns
        ``(best_point, optimiser)``.  Callbacks and saver are not persisted and have to be passed again.
        """
        param_space = ParameterSpace.from_file(os.path.join(folder, cls.PARAM_SPACE_FILE))
        gp_surr = gp_surrogate.from_saved(folder, **surrogate_kwargs)
        opt_attrs = load_json(os.path.join(folder, cls.OPT_ATTRS_FILE))
        optimiser = cls(parameter_space=param_space, callbacks=callbacks, saver=saver)
        for attr, value in opt_attrs.items():
            setattr(optimiser, attr, value)
        optimiser.gp_surr = gp_surr
        optimiser.group = getattr(gp_surr, "group", None)
        assert callable(objective_function)
        optimiser.obj_func = objective_function
        assert callable(eval_repeats_function)
        optimiser.eval_repeats_function = partial(eval_repeats_function, axis=0)
        return optimiser.resume_run(additional_budget=additional_budget), optimiser

    def __init__(
        self,
        parameter_space,
        gp_surrogate=None,
        exploration_method="tree",
        exploration_depth=5,
        budget=100,
        stopping_condition="evaluations",
        update_cycle=1,
        n_workers=1,
        callbacks=None,
        saver=None,
        group=None,
    ):
        """
        :param parameter_space: `ParameterSpace` to optimise over
        :param gp_surrogate: `GPSurrogate`; None = ``GPRSurrogate.default()``
        :param exploration_method: "tree" (ternary subtree of ``exploration_depth`` levels per child) or "sample"
            (``exploration_depth * ndim**2`` uniform samples per child)
        :param budget: number of evaluations / iterations / tree depth, see ``stopping_condition``
        :param stopping_condition: "evaluations", "iterations" or "depth"
        :param update_cycle: re-train the GP after this many new evaluations
        :param n_workers: processes used to evaluate the objective
        :param callbacks: list of `GPSOCallback`
        :param saver: object with ``save_runs(results, scores, params)``; then the objective returns (result, score)
        :param group: ``True`` / a ``torch.distributed`` group for a multi-GPU run (one process per GPU, every rank runs this
            same loop): rank 0 evaluates the objective and broadcasts the scores, fits and candidate scoring are sharded by
            the surrogate (``GPRSurrogate(group=...)``, which is created with this group when none is passed)
        """
        assert isinstance(parameter_space, ParameterSpace)
        self.param_space = parameter_space
        self.method = exploration_method
        if self.method == "tree":
            self.max_depth = exploration_depth
        elif self.method == "sample":
            self.max_depth = exploration_depth * self.param_space.ndim ** 2
        else:
            raise ValueError(f"Unknown exploration method: {self.method}")
        self.budget = budget
        assert stopping_condition in ["evaluations", "iterations", "depth"]
        self.stop_cond = stopping_condition
        self.update_cycle = update_cycle
        self.n_eval_counter = 0
        self.iterations = 0
        self.n_workers = n_workers
        callbacks = [] if callbacks is None else callbacks
        assert all(isinstance(callback, GPSOCallback) for callback in callbacks)
        self.callbacks = callbacks
        self.group = group if group is not None else getattr(gp_surrogate, "group", None)
        self.gp_surr = gp_surrogate or GPRSurrogate.default(group=self.group)
        assert isinstance(self.gp_surr, GPSurrogate)
        self.saver = saver
        if saver is not None:
            assert callable(getattr(self.saver, "save_runs", None))
        self._pool = None

    def _close_pool(self):
        """The worker pool lives for one ``run`` / ``resume_run`` (the reference builds one per evaluation call)."""
        pool, self._pool = getattr(self, "_pool", None), None
        if pool is not None:
            pool.close()
            pool.join()

    def __del__(self):
        try:
            self._close_pool()
        except Exception:
            pass

    # -----------------------------------------------------------------------------------------------------------------
    def _run_callbacks(self, callback_type):
        assert callback_type in CallbackTypes
        for callback in self.callbacks:
            if callback.callback_type == callback_type:
                callback.run(self)

    def _initialise(self, init_samples):
        """
        Evaluate the initial design -- ``init_samples`` (original coordinates) or, by default, two points per
        dimension at distance 0.25 from the centre along each axis -- plus the centre of the domain.
        """
        ndim = self.param_space.ndim
        mid = np.mean(NORM_PARAMS_BOUNDS)
        if init_samples is None:
            logging.info(
                "Sampling 2 vertices per dimension within L1 ball of 0.25 of the domain size radius in normalised "
                f"coordinates using {self.n_workers} worker(s)..."
            )
            radius = np.sum(NORM_PARAMS_BOUNDS) * 0.25
            normed = np.vstack([mid - radius * np.eye(ndim), mid + radius * np.eye(ndim)])
            orig_coords = self.param_space.denormalise_coords(normed)
        elif isinstance(init_samples, np.ndarray):
            assert init_samples.ndim == 2
            assert init_samples.shape[1] == ndim
            if init_samples.shape[0] <= 2:
                logging.warning(f"Only {init_samples.shape[0]} points selected for sampling, you might want to add more...")
            elif init_samples.shape[0] > (2 * ndim):
                logging.warning("Too many initial points obtained, you will run out of budget of objective function evaluations!")
            logging.info(
                f"Got {init_samples.shape[0]} points for initial sampling. Note that these are interpreted in the "
                "original parameter space coordinates!"
            )
            orig_coords = init_samples.copy()
        else:
            raise TypeError("init_samples must be None or a numpy array of original coordinates")

        centre = self.param_space.denormalise_coords(np.array([[mid] * ndim]))
        all_coords = np.vstack([orig_coords, centre])
        all_scores = self.evaluate_objective_function(all_coords)
        self.param_space.score = float(all_scores[-1])
        self.param_space.label = PointLabels.evaluated
        self.gp_surr.append(self.param_space.normalise_coords(all_coords), all_scores)
        logging.debug(
            f"Initialised with {all_coords.shape[0]} points:"
            + "".join(f"\n\t{coord}: {score}" for coord, score in zip(all_coords, all_scores))
        )

    def _gp_update(self, update_idx):
        """Re-train the GP when ``update_cycle`` new evaluations have arrived; refresh the UCB of GP-based leaves."""
        if (self.gp_surr.num_evaluated - update_idx) >= self.update_cycle:
            logging.info("Update step: retraining GP model and updating scores...")
            self.gp_surr.gp_update()
            points = self.gp_surr.points
            epoch, uid = getattr(points, "epoch", None), getattr(points, "uid", None)
            # every node once, order irrelevant here: the root's per-depth index instead of a generator walk of the tree
            by_depth = self.param_space._depth_index() if hasattr(self.param_space, "_depth_index") else None
            nodes = (leaf for level in by_depth.values() for leaf in level) if by_depth is not None else PreOrderIter(self.param_space)
            for leaf in nodes:
                # the position of a node's point does not change while points are only appended / replaced in place; the
                # cache names the list by its uid (not by reference: a pickled tree must not drag the point list along)
                cached = getattr(leaf, "_point_ref", None)
                if cached is not None and epoch is not None and cached[0] == uid and cached[1] == epoch:
                    leaf_point = points[cached[2]]
                else:
                    position = points.index_by_coords(leaf.center_array())
                    assert position is not None
                    leaf._point_ref = (uid, epoch, position)
                    leaf_point = points[position]
                if leaf_point.label == PointLabels.gp_based:
                    leaf.score = leaf_point.score_ucb
            self._run_callbacks(callback_type=CallbackTypes.post_update)
        return self.gp_surr.num_evaluated

    def _score_child(self, child, **kwargs):
        """(mean, var, ucb) of the best candidate inside ``child`` according to the exploration method."""
        if self.method == "tree":
            return self.gp_surr.gp_eval_best_ucb_in_leaf(child, depth=self.max_depth)
        samples = child.sample_uniformly(n_points=self.max_depth, seed=kwargs.pop("seed", None))
        rank, world = self._rank_world()
        if world > 1:
            # SPMD: unseeded draws differ between the ranks; every rank scores rank 0's samples
            import torch
            import torch.distributed as dist

            from .distributed import _comm_device

            group = None if self.group is True else self.group
            box = torch.tensor(np.ascontiguousarray(samples, dtype=np.float64), dtype=torch.float64, device=_comm_device(group))
            dist.broadcast(box, src=0, group=group)
            samples = box.cpu().numpy()
        return self.gp_surr.gp_eval_best_ucb(samples)

    def _tree_explore(self, levels_to_explore, **kwargs):
        """Exploration: per flagged level split the best leaf and score the children the GP has not seen yet."""
        logging.info("Exploration step: sampling children in the ternary tree...")
        assert len(levels_to_explore) == self.param_space.max_depth + 1
        points = self.gp_surr.points
        for level in range(self.param_space.max_depth + 1):
            if not levels_to_explore[level]:
                continue
            logging.debug(f"Exploring {level} level...")
            parent = self.param_space.get_best_score_leaf(depth=level)
            for child in parent.ternary_split():
                centre = child.center_array()
                if points.find_by_coords(centre) is None:
                    mean, var, ucb = self._score_child(child, **kwargs)
                    kwargs.pop("seed", None)  # the seed is consumed by the first sampled child, as in the reference
                    child.score = ucb
                    child.label = PointLabels.gp_based
                    points.append(
                        GPPoint(
                            normed_coord=np.array(centre),
                            score_mu=mean,
                            score_sigma=var,
                            score_ucb=ucb,
                            label=PointLabels.gp_based,
                        )
                    )
                else:
                    # the centre already carries a point (the middle child shares its parent's centre)
                    child.score = parent.score
                    child.label = parent.label
                logging.debug(f"{child.name} best score: {child.score}")
            parent.sampled = True

    def _tree_select(self):
        """Selection: per level take the best not-yet-sampled leaf; evaluate the objective there if it beats the levels above."""
        logging.info("Selecting step: evaluating best leaves...")
        max_score = -np.inf
        depth = self.param_space.max_depth
        levels_to_explore = [False] * (depth + 1)
        points = self.gp_surr.points
        for level in range(depth + 1):
            logging.debug(f"Selecting within {level} level...")
            max_leaf = self.param_space.get_best_score_leaf(depth=level, only_not_sampled=True)
            if not (max_leaf and max_leaf.score > max_score):
                continue
            levels_to_explore[level] = True
            max_score = float(max_leaf.score)
            leaf_point = points.find_by_coords(max_leaf.center_array())
            if leaf_point.label == PointLabels.gp_based:
                new_score = float(
                    np.ravel(
                        self.evaluate_objective_function(
                            self.param_space.denormalise_coords(leaf_point.normed_coord[np.newaxis, :])
                        )
                    )[0]
                )
                points.append(
                    GPPoint(
                        normed_coord=leaf_point.normed_coord,
                        score_mu=new_score,
                        score_sigma=0.0,
                        score_ucb=0.0,
                        label=PointLabels.evaluated,
                    )
                )
                max_leaf.score = new_score
                max_leaf.label = PointLabels.evaluated
                logging.debug(f"Leaf {max_leaf.name} updated to new evaluated score: {max_leaf.score}")
        logging.debug(f"Level to explore in the next iteration: {levels_to_explore}")
        return levels_to_explore

    def evaluate_objective_function(self, orig_coords):
        """Scores of the objective at ``orig_coords[n, ndim]`` (original coordinates), aggregated over the repeats."""
        assert orig_coords.ndim == 2
        assert orig_coords.shape[1] == self.param_space.ndim
        rank, world = self._rank_world()
        if world > 1:
            # SPMD: the objective runs once, on rank 0 (it may be expensive or stochastic); every rank gets the same scores
            import torch
            import torch.distributed as dist

            from .distributed import _comm_device

            group = None if self.group is True else self.group
            scores = self._evaluate_local(orig_coords) if rank == 0 else np.zeros(orig_coords.shape[0])
            box = torch.tensor(np.asarray(scores, dtype=np.float64).reshape(-1), dtype=torch.float64, device=_comm_device(group))
            dist.broadcast(box, src=0, group=group)
            if rank != 0:
                self.n_eval_counter += orig_coords.shape[0]
            return box.cpu().numpy()
        return self._evaluate_local(orig_coords)

    def _rank_world(self):
        if self.group is None:
            return 0, 1
        from .distributed import _rank_world

        return _rank_world(None if self.group is True else self.group)

    def _evaluate_local(self, orig_coords):
        repeated = np.vstack(self.eval_repeats * [orig_coords])
        if self.n_workers > 1 and (self.eval_repeats * orig_coords.shape[0]) > 1:
            if getattr(self, "_pool", None) is None:
                self._pool = _make_pool(self.n_workers)
            try:
                scores = list(self._pool.map(self.obj_func, repeated))
            except BaseException:
                self._close_pool()
                raise
        else:
            scores = [self.obj_func(coords) for coords in repeated]
        self.n_eval_counter += orig_coords.shape[0]  # repeats are not charged to the budget

        if self.saver is not None:
            results = [score[0] for score in scores]
            scores = [score[1] for score in scores]
            n_points = orig_coords.shape[0]
            for coord_idx, coords in enumerate(orig_coords):
                run_results = results[coord_idx::n_points]
                run_scores = scores[coord_idx::n_points]
                assert len(run_results) == len(run_scores) == self.eval_repeats
                self.saver.save_runs(run_results, run_scores, dict(zip(self.param_space.parameter_names, coords)))
        return self.eval_repeats_function(np.array(scores).astype(float).reshape((self.eval_repeats, -1)))

    def _stopping_condition(self):
        if self.stop_cond == "evaluations":
            return self.n_eval_counter < self.budget
        elif self.stop_cond == "iterations":
            return self.iterations < self.budget
        elif self.stop_cond == "depth":
            return self.param_space.max_depth <= self.budget

    # -----------------------------------------------------------------------------------------------------------------
    def _iterate(self, explore_levels, update_idx):
        """The explore / select / update loop shared by ``run`` and ``resume_run``."""
        try:
            return self._iterate_loop(explore_levels, update_idx)
        finally:
            self._close_pool()

    def _iterate_loop(self, explore_levels, update_idx):
        keep_going = True
        while keep_going:
            self._run_callbacks(callback_type=CallbackTypes.pre_iteration)
            self._tree_explore(levels_to_explore=explore_levels, seed=self.expl_seed)
            explore_levels = self._tree_select()
            update_idx = self._gp_update(update_idx)
            self.iterations += 1
            highest_ucb = self.gp_surr.highest_ucb
            logging.info(
                f"After {self.iterations}th iteration: \n\t number of obj. func. evaluations: {self.n_eval_counter} \n\t"
                f" highest score: {self.gp_surr.highest_score.score_mu} \n\t highest UCB: "
                f"{highest_ucb.score_ucb if highest_ucb is not None else None}"
            )
            logging.debug(
                f"\n\t Total number of points: {len(self.gp_surr.points)} \n\t evaluated points: "
                f"{self.gp_surr.num_evaluated} \n\t GP-based estimates: {self.gp_surr.num_gp_based} \n\t depth of the "
                f"tree: {self.param_space.max_depth}"
            )
            self._run_callbacks(callback_type=CallbackTypes.post_iteration)
            keep_going = self._stopping_condition()
        logging.info(f"Done. Highest evaluated score: {self.gp_surr.highest_score.score_mu}")
        self._run_callbacks(callback_type=CallbackTypes.pre_finalise)
        self.last_explored_levels = explore_levels
        self.last_update_idx = update_idx
        return self.gp_surr.highest_score

    def run(self, objective_function, init_samples=None, eval_repeats=1, eval_repeats_function=np.mean, **kwargs):
        """
        Run the optimisation and return the evaluated point with the highest score.

        :param objective_function: callable taking one point in original coordinates, returning a scalar score (or
            ``(result, score)`` when a saver is used)
        :param init_samples: initial design in original coordinates, or None for the default diamond design
        :param eval_repeats: evaluations per point for stochastic objectives (not charged to the budget)
        :param eval_repeats_function: aggregator over the repeats, must accept ``axis``
        :kwargs: ``seed`` for the "sample" exploration method
        """
        assert callable(objective_function)
        self.obj_func = objective_function
        self.eval_repeats = eval_repeats
        assert callable(eval_repeats_function)
        self.eval_repeats_function = partial(eval_repeats_function, axis=0)
        self.expl_seed = kwargs.pop("seed", None)
        logging.info(
            f"Starting {self.param_space.ndim}-dimensional optimisation with budget of {self.budget} objective function "
            "evaluations..."
        )
        self._initialise(init_samples)
        self._run_callbacks(callback_type=CallbackTypes.post_initialise)
        update_idx = self._gp_update(0)
        return self._iterate(explore_levels=[True], update_idx=update_idx)

    def resume_run(self, additional_budget):
        """Continue a finished run for ``additional_budget`` more budget units."""
        assert callable(self.obj_func)
        assert callable(self.eval_repeats_function)
        assert self.iterations > 0
        self.budget += additional_budget
        logging.info(f"Resuming optimisation for with additional budget of {additional_budget}")
        return self._iterate(explore_levels=self.last_explored_levels, update_idx=self.last_update_idx)

    def save_state(self, folder):
        """Persist tree, surrogate and loop counters to ``folder`` (callbacks and saver are not saved)."""
        if self._rank_world()[0] != 0:
            return  # multi-GPU run: every rank holds the same state, rank 0 writes it
        make_dirs(folder)
        logging.warning("When saving, all callbacks and saver will be lost!")
        self.param_space.save(os.path.join(folder, self.PARAM_SPACE_FILE))
        self.gp_surr.save(folder)
        opt_attrs = {attr: getattr(self, attr) for attr in self.SAVE_ATTRS}
        with open(os.path.join(folder, self.OPT_ATTRS_FILE), "w") as handle:
            handle.write(json.dumps(opt_attrs))
        logging.info(f"Saved optimiser to {folder}")
