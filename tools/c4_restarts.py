#!/usr/bin/env python
"""
Config C4 (hyper-parameter fit by multi-start L-BFGS-B, N=8192, d=20) on the GPUs of one box.  Under torchrun the
restarts are dealt round-robin over the ranks (pygpso_b200.distributed.sharded_multistart_fit); alone it runs them in
sequence on cuda:0.  The full config is 64 restarts x maxiter 50; the defaults here are a bounded sample of it.

    python tools/c4_restarts.py [restarts] [maxiter] [N] [d] [out.json]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def main():
    import torch
    import torch.distributed as dist

    from pygpso_b200 import backend, gpmodel
    from pygpso_b200.distributed import sharded_multistart_fit

    restarts = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    maxiter = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    N = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
    d = int(sys.argv[4]) if len(sys.argv) > 4 else 20
    out_path = sys.argv[5] if len(sys.argv) > 5 else os.path.join(ROOT, "gpurun_out", "c4_restarts.json")
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("gloo", rank=0, world_size=1)
    X, y = bench.synthetic_training(N, d)
    model = gpmodel.GPR(data=(X, y), kernel=gpmodel.Matern52(lengthscales=0.25 * np.sqrt(d)), mean_function=gpmodel.Constant(0.0),
                        noise_variance=1e-3, backend=backend.CudaBackend(device=local))
    evals = [0]
    dev_ms = [0.0]

    def objective(u):
        f, g = model.neg_log_marginal_likelihood_and_grad(u)
        evals[0] += 1
        dev_ms[0] += model._session.last_timing_ms()[0]
        return f, g

    objective(model._pack())  # allocations, task queue
    evals[0], dev_ms[0] = 0, 0.0
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    u_best, f_best, which, table = sharded_multistart_fit(objective, model._pack(), restarts, maxiter=maxiter)
    torch.cuda.synchronize()
    dist.barrier()
    wall = time.perf_counter() - t0
    counts = torch.tensor([float(evals[0]), dev_ms[0]], dtype=torch.float64, device="cuda" if world > 1 else "cpu")
    dist.all_reduce(counts)
    if rank == 0:
        total_evals = int(counts[0].item())
        report = {"workload": f"C4 sample: {restarts} L-BFGS-B restarts (maxiter {maxiter}) of the LML fit, N={N}, d={d}, Matern-5/2",
                  "n_gpus": world, "wall_s": wall, "lml_grad_evaluations": total_evals,
                  "evals_per_s_whole_job": total_evals / wall, "device_ms_per_eval": counts[1].item() / max(total_evals, 1),
                  "best_restart": which, "best_neg_lml": f_best, "neg_lml_per_restart": [float(v) for v in table[:, 0]]}
        os.makedirs(os.path.dirname(out_path), exist_ok=True)
        with open(out_path, "w") as fh:
            json.dump(report, fh, indent=1)
        print(json.dumps(report))
    model.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
