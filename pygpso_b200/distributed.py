"""
Multi-GPU sharding of the two axes of the path that shard naturally -- one process per GPU,
``torch.distributed`` (NCCL over NVLink/NVSwitch on the B200 box, gloo in the CPU tests) for the plumbing:

  1. candidates: rows of the candidate matrix are independent, so rank r scores ``shard_bounds(M, W, r)`` with the fused
     predict_y + UCB + arg-max kernel locally.  One exchange per *fit*: the fitting rank broadcasts its state (scaled
     inputs, alpha, L^-1, hyper-parameters) as ONE contiguous buffer (``gpso_export_state_dev`` -> NCCL broadcast ->
     ``gpso_import_state_dev``).  One exchange per *scoring call*: an all-gather of a 32-byte record
     (ucb, global index, mean, var) per rank; every rank then picks the winner with numpy's arg-max rule (first NaN,
     else the largest UCB, lowest global index on ties), which is exactly what a single GPU / the reference returns.
  2. multi-start restarts of the hyper-parameter fit: restart i runs on rank i mod W (independent L-BFGS-B runs),
     followed by an all-gather of (-LML*, u*); the winner is the smallest -LML, lowest restart id on ties.

Nothing here does arithmetic on candidates or Gram matrices; it only partitions, exchanges and selects.
"""
import numpy as np
import scipy.optimize

RECORD_LEN = 4  # ucb, global index, mean, var


def shard_bounds(total, world_size, rank):
    """Contiguous shard [start, stop) of ``total`` items for ``rank``; the first ``total % world_size`` ranks get one more."""
    base, extra = divmod(int(total), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def pick_best(records):
    """
    Winner among per-shard records ``[[ucb, global_idx, mean, var], ...]`` with np.argmax semantics over the
    concatenated candidate list: the first NaN if any, else the largest UCB, the lowest global index on ties.
    Records of empty shards carry ``global_idx < 0`` and are skipped.  Returns (global_idx, mean, var, ucb).
    """
    best = None
    for ucb, gidx, mean, var in np.asarray(records, dtype=np.float64).reshape(-1, RECORD_LEN):
        if gidx < 0:
            continue
        cand = (float(ucb), int(gidx), float(mean), float(var))
        if best is None:
            best = cand
            continue
        cn, bn = np.isnan(cand[0]), np.isnan(best[0])
        if cn or bn:
            better = (cn and not bn) or (cn and bn and cand[1] < best[1])
        else:
            better = cand[0] > best[0] or (cand[0] == best[0] and cand[1] < best[1])
        if better:
            best = cand
    if best is None:
        raise ValueError("no candidates in any shard")
    return best[1], best[2], best[3], best[0]


def merge_topk(records, k):
    """
    The k best of per-shard top-k lists (rows ``[ucb, global_idx, mean, var]``; ``global_idx < 0`` = padding) in arg-max
    order: NaNs first (by index), then descending UCB, the lowest global index on ties.  Returns rows
    ``[global_idx, mean, var, ucb]`` like ``session.ucb_topk``.
    """
    rows = np.asarray(records, dtype=np.float64).reshape(-1, RECORD_LEN)
    rows = rows[rows[:, 1] >= 0]
    nan = np.isnan(rows[:, 0])
    order = np.lexsort((rows[:, 1], -np.where(nan, np.inf, rows[:, 0]), ~nan))
    best = rows[order[:k]]
    return best[:, [1, 2, 3, 0]]


def _dist():
    import torch.distributed as dist

    return dist


def _comm_device(group=None):
    """Device on which collective buffers must live: CUDA for NCCL, CPU for gloo."""
    import torch

    dist = _dist()
    if dist.get_backend(group) == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def gather_records(record, group=None):
    """All-gather the same number of RECORD_LEN-double records from every rank (one, or k for the top-k path); returns
    an array [world * records_per_rank, RECORD_LEN] on every rank, rank-major."""
    import torch

    dist = _dist()
    world = dist.get_world_size(group)
    dev = _comm_device(group)
    mine = torch.tensor(np.asarray(record, dtype=np.float64).reshape(-1), dtype=torch.float64, device=dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    return torch.stack(parts).cpu().numpy().reshape(-1, RECORD_LEN)


class ShardedScorer:
    """
    Candidate-sharded ``gp_eval_best_ucb``.  ``session`` is this rank's device session (``model._session``); any object
    with ``ucb_argmax(X, varsigma) -> (idx, mean, var, ucb)`` works, which is how the gloo test drives it on CPU.
    """

    def __init__(self, session, group=None):
        self.session = session
        self.group = group
        dist = _dist()
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    # -- one exchange per fit -----------------------------------------------------------------------------------------
    def broadcast_fit(self, n, d, src=0):
        """Broadcast the fitted state of rank ``src`` to every rank's session: one contiguous device buffer
        (``gpso_export_state_dev`` -> broadcast -> ``gpso_import_state_dev``) over NCCL; with gloo the same buffer is staged
        through host memory.  Sessions without device state (the checker backend of the CPU tests) exchange
        ``get_state()`` / ``set_state()`` objects."""
        import torch

        dist = _dist()
        if not hasattr(self.session, "export_state_dev"):
            box = [self.session.get_state() if self.rank == src else None]
            dist.broadcast_object_list(box, src=src, group=self.group)
            if self.rank != src:
                self.session.set_state(box[0])
            return 0
        nbytes = self.session.state_bytes(n, d)
        dev = torch.device("cuda", torch.cuda.current_device())
        buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream().cuda_stream
        if self.rank == src:
            self.session.export_state_dev(buf.data_ptr(), nbytes, stream)
        if _comm_device(self.group).type == "cuda":
            dist.broadcast(buf, src=src, group=self.group)
        else:
            staged = buf.cpu()
            dist.broadcast(staged, src=src, group=self.group)
            if self.rank != src:
                buf.copy_(staged)
        if self.rank != src:
            torch.cuda.current_stream().synchronize()
            self.session.import_state_dev(buf.data_ptr(), nbytes, n, d, stream)
        return nbytes

    # -- one exchange per scoring call --------------------------------------------------------------------------------
    def local_record(self, x_local, global_offset, varsigma):
        if len(x_local) == 0:
            return [-np.inf, -1.0, 0.0, 0.0]
        idx, mean, var, ucb = self.session.ucb_argmax(x_local, varsigma)
        return [ucb, float(global_offset + idx), mean, var]

    def ucb_argmax(self, x_local, global_offset, varsigma):
        """Score this rank's shard (rows ``global_offset ...`` of the full candidate list) and agree on the winner."""
        records = gather_records(self.local_record(x_local, global_offset, varsigma), self.group)
        return pick_best(records)

    def ucb_topk(self, x_local, global_offset, varsigma, k):
        """Top-k over all shards: k records per rank are gathered (32 k bytes each) and merged identically on every rank."""
        mine = np.full((k, RECORD_LEN), [-np.inf, -1.0, 0.0, 0.0])
        if len(x_local):
            local = np.asarray(self.session.ucb_topk(x_local, varsigma, k), dtype=np.float64).reshape(-1, 4)
            mine[: len(local)] = np.column_stack([local[:, 3], local[:, 0] + global_offset, local[:, 1], local[:, 2]])
        return merge_topk(gather_records(mine, self.group), k)

    def grow_ucb_argmax(self, bounds, depth, varsigma):
        """``gp_eval_best_ucb(leaf.grow(depth))`` with the rows of the leaf batch sharded over the ranks: every rank generates
        and scores its own contiguous slice on its GPU (no candidate ever crosses a link), one record per rank is gathered."""
        total = (3 ** int(depth) - 1) // 2
        start, stop = shard_bounds(total, self.world, self.rank)
        record = [-np.inf, -1.0, 0.0, 0.0]
        if stop > start:
            idx, mean, var, ucb = self.session.grow_ucb_argmax(bounds, depth, varsigma, rows=(start, stop))
            record = [ucb, float(idx), mean, var]
        return pick_best(gather_records(record, self.group))

    def ucb_argmax_full(self, x_all, varsigma):
        """Convenience: every rank holds the full candidate matrix and scores only its own contiguous shard."""
        start, stop = shard_bounds(len(x_all), self.world, self.rank)
        return self.ucb_argmax(x_all[start:stop], start, varsigma)


def restart_points(u0, n_restarts, seed=20240517):
    """Restart 0 is the warm start ``u0``; restart i > 0 is ``u0 + N(0, 1)`` in unconstrained space."""
    rng = np.random.default_rng(seed)
    starts = [np.array(u0, dtype=np.float64)]
    for _ in range(1, n_restarts):
        starts.append(u0 + rng.normal(0.0, 1.0, size=len(u0)))
    return starts


def _rank_world(group=None):
    """(rank, world) of this process; (0, 1) when torch.distributed is not initialised (single-process use)."""
    try:
        dist = _dist()
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(group), dist.get_world_size(group)
    except ImportError:
        pass
    return 0, 1


def sharded_multistart_fit(objective, u0, n_restarts, group=None, seed=20240517, maxiter=50, on_error=np.inf):
    """
    Multi-start L-BFGS-B with the restarts dealt round-robin over the ranks.  ``objective(u) -> (f, grad)`` is this
    rank's device closure (``model.neg_log_marginal_likelihood_and_grad``).  Returns (u_best, f_best, restart_id, table)
    identically on every rank; ``table[i] = (f_i, restart i's optimum)``.  Restart 0 is the warm start ``u0`` itself, so
    ``n_restarts=1`` with ``maxiter=None`` (SciPy's default budget) is the reference's single fit.  Works without an
    initialised process group (all restarts run here).
    """
    rank, world = _rank_world(group)
    starts = restart_points(u0, n_restarts, seed)
    p = len(u0)
    options = {} if maxiter is None else {"maxiter": maxiter}
    mine = np.full((n_restarts, p + 1), np.nan)
    for i in range(rank, n_restarts, world):
        try:
            res = scipy.optimize.minimize(objective, starts[i], jac=True, method="L-BFGS-B", options=options)
            mine[i, 0], mine[i, 1:] = res.fun, res.x
        except np.linalg.LinAlgError:
            mine[i, 0], mine[i, 1:] = on_error, starts[i]
    if world > 1:
        import torch

        dist = _dist()
        local = torch.tensor(mine, dtype=torch.float64, device=_comm_device(group))
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local, group=group)
        # row i is taken verbatim from its owner (rank i mod W): a restart whose own optimum is NaN/inf stays non-finite here
        # and is ranked last below, instead of turning into a spurious f = 0 under a SUM reduction
        table = np.stack([parts[i % world][i].cpu().numpy() for i in range(n_restarts)])
    else:
        table = mine
    f = np.where(np.isfinite(table[:, 0]), table[:, 0], np.inf)
    best = int(np.argmin(f))  # first minimum = lowest restart id on ties
    return table[best, 1:].copy(), float(f[best]), best, table


def multistart_fit(surrogate, model, n_restarts, group=None, seed=20240517, maxiter=50):
    """
    The hyper-parameter fit of ``GPRSurrogate._gp_train`` with ``n_restarts`` restarts dealt round-robin over the ranks of
    ``group`` (SPMD: every rank calls it with a model holding the same data and hyper-parameters).

    Restart 0 is the reference's own fit -- ``surrogate.optimiser.minimize(model.training_loss, ...)`` from the warm start,
    whatever optimiser the surrogate was given, with its full budget (reference gp_surrogate.py:500-503) -- so
    ``n_restarts=1`` is the reference trajectory.  Restart i > 0 runs SciPy L-BFGS-B (``maxiter`` iterations) from
    ``u0 + N(0,1)`` in unconstrained space.  All ranks end with the winner's hyper-parameters installed.  Returns a dict with
    the table of (-LML*, u*) per restart.
    """
    rank, world = _rank_world(group)
    u0 = model._pack()
    starts = restart_points(u0, n_restarts, seed)
    p = len(u0)
    mine = np.full((n_restarts, p + 1), np.nan)
    for i in range(rank, n_restarts, world):
        try:
            if i == 0:
                model._unpack(u0)
                surrogate.optimiser.minimize(model.training_loss, model.trainable_variables)
                x = model._pack()
                mine[i, 0], mine[i, 1:] = model.neg_log_marginal_likelihood_and_grad(x)[0], x
            else:
                res = scipy.optimize.minimize(model.neg_log_marginal_likelihood_and_grad, starts[i], jac=True, method="L-BFGS-B",
                                              options={} if maxiter is None else {"maxiter": maxiter})
                mine[i, 0], mine[i, 1:] = res.fun, res.x
        except (np.linalg.LinAlgError, ValueError):
            mine[i, 0], mine[i, 1:] = np.inf, starts[i]
    if world > 1:
        import torch

        dist = _dist()
        local = torch.tensor(mine, dtype=torch.float64, device=_comm_device(group))
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local, group=group)
        table = np.stack([parts[i % world][i].cpu().numpy() for i in range(n_restarts)])
    else:
        table = mine
    f = np.where(np.isfinite(table[:, 0]), table[:, 0], np.inf)
    best = int(np.argmin(f))  # restart 0 wins ties: the reference's optimum is kept unless a restart is strictly better
    model._unpack(table[best, 1:])
    return {"restart": best, "fun": float(f[best]), "x": table[best, 1:].copy(), "table": table}
