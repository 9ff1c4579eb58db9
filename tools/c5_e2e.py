#!/usr/bin/env python
"""
Config C5 at full size through the public API: GPSOptimiser on the 10-D Rastrigin function (maximised), budget 500
evaluations, ternary exploration tree of depth 12 (265 720 leaf candidates per scored child), surrogate on the GPU.
Reports wall time and where it went (fit closure, fused grow+score calls, host-side bookkeeping).

    python tools/c5_e2e.py [budget] [depth] [out.json]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pygpso_b200 import GPRSurrogate, GPSOptimiser, ParameterSpace, gpmodel  # noqa: E402


def rastrigin_max(point):
    x = np.asarray(point, dtype=np.float64)
    return -(10.0 * x.size + np.sum(x * x - 10.0 * np.cos(2.0 * np.pi * x)))


class Stopwatch:
    """Wraps a bound method and accumulates calls and seconds."""

    def __init__(self, owner, name):
        self.calls, self.seconds = 0, 0.0
        self._fn = getattr(owner, name)
        setattr(owner, name, self)

    def __call__(self, *args, **kwargs):
        t0 = time.perf_counter()
        try:
            return self._fn(*args, **kwargs)
        finally:
            self.calls += 1
            self.seconds += time.perf_counter() - t0


def main():
    budget = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    depth = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    out_path = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "c5_e2e.json")
    d = 10
    space = ParameterSpace(parameter_names=[f"p{i}" for i in range(d)], parameter_bounds=[[-4.1, 5.12]] * d)
    surr = GPRSurrogate.default()
    opt = GPSOptimiser(parameter_space=space, gp_surrogate=surr, exploration_method="tree", exploration_depth=depth, budget=budget,
                       stopping_condition="evaluations", update_cycle=1, n_workers=1)
    watches = {"gp_update (fit + centre re-prediction)": Stopwatch(surr, "gp_update"),
               "_gp_train (L-BFGS-B, closure on the GPU)": Stopwatch(surr, "_gp_train"),
               "gp_eval_best_ucb_in_leaf (fused grow + score + arg-max)": Stopwatch(surr, "gp_eval_best_ucb_in_leaf"),
               "_tree_explore": Stopwatch(opt, "_tree_explore"),
               "_tree_select": Stopwatch(opt, "_tree_select")}
    t0 = time.perf_counter()
    best = opt.run(rastrigin_max)
    wall = time.perf_counter() - t0
    model = surr.gpflow_model
    closure = Stopwatch(model, "neg_log_marginal_likelihood_and_grad")  # (installed after the run: count only)
    leaves = (3 ** depth - 1) // 2
    scored = watches["gp_eval_best_ucb_in_leaf (fused grow + score + arg-max)"]
    report = {
        "workload": f"C5: GPSOptimiser, 10-D Rastrigin (maximised), budget {budget} evaluations, ternary tree depth {depth} "
                    f"({leaves} leaf candidates per scored child)",
        "wall_s": wall, "evaluations": opt.n_eval_counter, "iterations": opt.iterations,
        "training_points_at_end": surr.num_evaluated, "gp_based_points": surr.num_gp_based,
        "best_score": float(best.score_mu), "best_normed_coord": [float(v) for v in best.normed_coord],
        "stages": {k: {"calls": w.calls, "seconds": w.seconds} for k, w in watches.items()},
        "candidates_scored": scored.calls * leaves,
        "candidates_per_s_inside_scoring_calls": scored.calls * leaves / max(scored.seconds, 1e-9),
        "kernel_launches": model._session.launch_count() if getattr(model, "_session", None) else None,
    }
    del closure
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as fh:
        json.dump(report, fh, indent=1)
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
