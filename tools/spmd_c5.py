#!/usr/bin/env python
"""
Config C5 (10-D Rastrigin, 500-evaluation budget, depth-12 ternary exploration) through the kept API on 1..8 GPUs:

    python tools/spmd_c5.py [out.json]                                         # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/spmd_c5.py [out.json]

Every rank builds ``GPSOptimiser(..., gp_surrogate=GPRSurrogate.default(group=True, n_restarts=R))`` and runs the same loop
(SPMD): restarts of every fit are dealt over the ranks, the fitted state is broadcast once per fit over NCCL, every leaf batch
(265 720 candidates per explored child) is sharded over the ranks.  Prints the wall time, the decisions' fingerprint and
whether all ranks ended in the same state.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rastrigin(point):
    x = np.asarray(point)
    return -float(10 * x.size + np.sum(x * x - 10 * np.cos(2 * np.pi * x)))


def main():
    import torch
    import torch.distributed as dist

    from pygpso_b200 import GPRSurrogate, GPSOptimiser, ParameterSpace

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ["GPSO_DEVICE"] = str(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    restarts = int(os.environ.get("GPSO_RESTARTS", "1"))
    budget = int(os.environ.get("GPSO_BUDGET", "500"))
    group = True if world > 1 else None
    space = ParameterSpace(parameter_names=[f"p{i}" for i in range(10)], parameter_bounds=[[-5.12, 5.12]] * 10)
    surr = GPRSurrogate.default(group=group, n_restarts=restarts)
    opt = GPSOptimiser(parameter_space=space, gp_surrogate=surr, exploration_method="tree", exploration_depth=12, budget=budget,
                       stopping_condition="evaluations", update_cycle=1, n_workers=1)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    best = opt.run(rastrigin)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ucbs = np.array([p.score_ucb for p in surr.points])
    digest = hashlib.sha256(ucbs.tobytes() + np.asarray(best.normed_coord).tobytes()).hexdigest()[:16]
    same = True
    if world > 1:
        box = [None] * world
        dist.all_gather_object(box, digest)
        same = len(set(box)) == 1
        t = torch.tensor([wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall = float(t.item())
    if rank == 0:
        out = {"config": "C5: 10-D Rastrigin, budget %d, exploration depth 12 (265 720 leaf candidates per scored child)" % budget,
               "n_gpus": world, "restarts_per_fit": restarts, "wall_s": wall, "evaluations": opt.n_eval_counter, "iterations": opt.iterations,
               "fits": surr.gpflow_model.n_loss_evaluations, "best_score": best.score_mu, "points": len(surr.points),
               "state_digest": digest, "all_ranks_same_state": bool(same)}
        print(json.dumps(out), flush=True)
        if len(sys.argv) > 1:
            with open(sys.argv[1], "w") as fh:
                json.dump(out, fh, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
