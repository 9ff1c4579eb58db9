"""
Checker backend for CPU tests: implements the session interface of ``pygpso_b200.backend`` on top of the numpy oracle.
It exists so that the *host* logic (optimiser loop, point store, tree, persistence) can be exercised and pinned against
the reference's golden numbers on a machine without a GPU.  It lives under ``tests/`` and is only ever injected
explicitly (``GPRSurrogate(..., backend=OracleBackend())``); the package never selects it by itself.
"""
import numpy as np

from oracle import gpr_oracle as go
from oracle import grow_oracle


class OracleSession:
    def __init__(self, kernel, n_lengthscales, has_mean, rowwise=False):
        # rowwise: predict one candidate at a time, so that a candidate's result does not depend on which batch (shard) it
        # arrives in -- the property the CUDA kernels have by construction (BLAS blocking makes numpy's last bits vary)
        self.rowwise = rowwise
        self.kernel = kernel
        self.n_ls = n_lengthscales
        self.has_mean = has_mean
        self.X = self.y = self.h = None
        self.calls = {"neg_lml_and_grad": 0, "factorize": 0, "predict_y": 0, "ucb_argmax": 0, "candidates": 0}

    def set_data(self, x, y):
        self.X = np.ascontiguousarray(x, dtype=np.float64)
        self.y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1, 1)
        self.h = None

    def neg_lml_and_grad(self, u):
        self.calls["neg_lml_and_grad"] += 1
        return go.neg_lml_and_grad(self.kernel, self.X, self.y, u, self.n_ls, self.has_mean)

    def factorize(self, theta):
        self.calls["factorize"] += 1
        n = self.n_ls
        ls = theta[:n] if n > 1 else theta[0]
        self.h = go.Hyper(ls, theta[n], theta[n + 1], theta[n + 2] if self.has_mean else None)

    def log_marginal_likelihood(self):
        return go.lml(self.kernel, self.X, self.y, self.h)

    def _predict(self, xnew):
        if not self.rowwise:
            return go.predict_y(self.kernel, self.X, self.y, self.h, xnew)
        parts = [go.predict_y(self.kernel, self.X, self.y, self.h, xnew[i:i + 1]) for i in range(xnew.shape[0])]
        return np.vstack([p[0] for p in parts]), np.vstack([p[1] for p in parts])

    def predict_y(self, xnew):
        self.calls["predict_y"] += 1
        self.calls["candidates"] += xnew.shape[0]
        mean, var = self._predict(xnew)
        return mean[:, 0], var[:, 0]

    def ucb_argmax(self, xnew, varsigma):
        self.calls["ucb_argmax"] += 1
        self.calls["candidates"] += xnew.shape[0]
        mean, var = self._predict(xnew)
        return go.ucb_argmax(mean, var, varsigma)

    def ucb_topk(self, xnew, varsigma, k):
        mean, var = go.predict_y(self.kernel, self.X, self.y, self.h, xnew)
        mean, var = mean[:, 0], var[:, 0]
        ucb = mean + varsigma * var
        nan = np.isnan(ucb)
        order = np.lexsort((np.arange(len(ucb)), -np.where(nan, np.inf, ucb), ~nan))[:k]
        return np.column_stack([order.astype(float), mean[order], var[order], ucb[order]])

    def grow_ucb_argmax(self, bounds, depth, varsigma, rows=None):
        leaves = grow_oracle.grow_by_level(bounds, depth)
        if rows is None:
            return self.ucb_argmax(leaves, varsigma)
        idx, mean, var, ucb = self.ucb_argmax(leaves[rows[0]:rows[1]], varsigma)
        return idx + rows[0], mean, var, ucb

    # fitted state as one picklable object (what ShardedScorer.broadcast_fit exchanges when there is no device buffer)
    def get_state(self):
        return (self.X, self.y, self.h)

    def set_state(self, state):
        self.X, self.y, self.h = state

    def close(self):
        pass


class OracleBackend:
    name = "numpy-oracle (tests only)"

    def __init__(self, rowwise=False):
        self.sessions = []
        self.rowwise = rowwise

    def open_session(self, kernel, n_lengthscales, has_mean):
        session = OracleSession(kernel, n_lengthscales, has_mean, rowwise=self.rowwise)
        self.sessions.append(session)
        return session

    def grow_count(self, depth):
        return (3 ** depth - 1) // 2

    def grow_leaves(self, bounds, depth):
        return grow_oracle.grow_by_level(bounds, depth)
