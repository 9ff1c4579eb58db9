"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU (numpy / scipy, fp64) restatement of the Gaussian-process-regression arithmetic that pyGPSO delegates to
GPflow + TensorFlow + SciPy.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module; ``pygpso_b200`` never does.

Why a restatement: the arithmetic of the hot path is *not* in /root/reference.  It lives in the un-vendored,
un-pinned dependency ``gpflow>=2.0.0`` (reference ``requirements.txt:5``; pygpso 0.6.1 era => GPflow 2.0-2.2 on
TensorFlow 2.1-2.5) plus ``scipy.optimize.minimize(method="L-BFGS-B")``.  Neither GPflow nor TensorFlow is
installable here (no wheels, no network), so the published GPflow-2 algorithm is restated below and *pinned* against
every golden number the reference itself ships (see ``tests/test_oracle_goldens.py``):

  * ``tests/test_gp_surrogate.py:259-309``   fit -> predict KAT  (mean 0.61633117, var 0.06010023, UCB formula)
  * ``tests/test_optimisation.py:22-23``     end-to-end best point [0.23525377, 0.68518519] / 8.10560594
  * ``examples/0-basic-optimisation.ipynb:208-295``  13-iteration trace (evals, highest score, highest UCB)
  * ``examples/1-callbacks.ipynb:294-526``   14 fitted hyper-parameter rows

Reference call sites this file stands in for (all under /root/reference/gpso/):
  gp_surrogate.py:490-495  gpflow.models.GPR(data, kernel, mean_function, noise_variance)
  gp_surrogate.py:500-503  optimiser.minimize(model.training_loss, model.trainable_variables)
  gp_surrogate.py:298,325  model.predict_y(Xnew) -> (mean[M,1], var[M,1])   (noise variance INCLUDED)
  gp_surrogate.py:326-328  ucb = mean + varsigma * var ; np.argmax (first max)

GPflow-2 semantics restated (function names are GPflow's):
  gpflow.utilities.positive()            -> softplus transform;  Gaussian likelihood variance: 1e-6 + softplus(u)
  gpflow.utilities.ops.square_distance   -> |x|^2 + |x'|^2 - 2 x.x'  on inputs pre-divided by the lengthscale
  gpflow.kernels.Matern52/32/12.K_r, SquaredExponential.K_r2  (Matern: r = sqrt(max(r2, 1e-36)))
  gpflow.models.GPR.log_marginal_likelihood   -> multivariate_normal(y, m, L) with L = chol(K + s2 I)
  gpflow.conditionals.base_conditional   -> A = L^-1 Kmn ; fvar = kdiag - colsum(A^2) ; A = L^-T A ; fmean = A^T (y-m)
  gpflow.likelihoods.Gaussian.predict_mean_and_var -> (fmean + m(x*), fvar + s2)
  gpflow.optimizers.Scipy.minimize       -> scipy L-BFGS-B on the packed *unconstrained* vector, jac=True, defaults
  trainable_variables order              -> kernel.lengthscales, kernel.variance, likelihood.variance, mean_function.c
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla
import scipy.optimize as sopt
from scipy.special import erfcinv

KERNEL_IDS = {"Matern12": 0, "Matern32": 1, "Matern52": 2, "SquaredExponential": 3}
NOISE_FLOOR = 1.0e-6  # gpflow.likelihoods.Gaussian: variance = Parameter(v, transform=positive(lower=1e-6))
R2_CLIP = 1.0e-36  # gpflow.kernels.stationaries.IsotropicStationary.scaled_squared_euclid_dist -> sqrt(max(r2,1e-36))
VARSIGMA_DEFAULT = float(erfcinv(0.01))  # gp_surrogate.py:139


# ---------------------------------------------------------------------------------------------------------------------
# transforms (gpflow.utilities.positive -> tfp.bijectors.Softplus [+ Shift])
# ---------------------------------------------------------------------------------------------------------------------
def softplus(u):
    return np.logaddexp(0.0, u)


def softplus_inv(v):
    v = np.asarray(v, dtype=np.float64)
    # log(expm1(v)); beyond v = 30 the overflow-free form v + log1p(-exp(-v)) (tfp.math.softplus_inverse is stable too)
    small = np.minimum(v, 30.0)
    return np.where(v > 30.0, v + np.log1p(-np.exp(-np.maximum(v, 30.0))), np.log(np.expm1(small)))


def sigmoid(u):
    return 1.0 / (1.0 + np.exp(-u))


class Hyper:
    """Constrained hyper-parameters <-> packed unconstrained vector, in GPflow's trainable_variables order."""

    def __init__(self, lengthscales, variance, noise_variance, mean_c=None):
        self.lengthscales = np.atleast_1d(np.asarray(lengthscales, dtype=np.float64)).copy()
        self.ard = np.ndim(lengthscales) > 0 and np.size(lengthscales) > 1
        self.variance = float(variance)
        self.noise_variance = float(noise_variance)
        self.mean_c = None if mean_c is None else float(np.ravel(mean_c)[0])

    @property
    def has_mean(self):
        return self.mean_c is not None

    def pack(self):
        u = list(softplus_inv(self.lengthscales))
        u.append(float(softplus_inv(self.variance)))
        u.append(float(softplus_inv(self.noise_variance - NOISE_FLOOR)))
        if self.has_mean:
            u.append(self.mean_c)
        return np.array(u, dtype=np.float64)

    @classmethod
    def unpack(cls, u, n_ls, has_mean):
        u = np.asarray(u, dtype=np.float64)
        ls = softplus(u[:n_ls])
        var = float(softplus(u[n_ls]))
        noise = float(NOISE_FLOOR + softplus(u[n_ls + 1]))
        c = float(u[n_ls + 2]) if has_mean else None
        h = cls(ls if n_ls > 1 else ls[0], var, noise, c)
        return h


# ---------------------------------------------------------------------------------------------------------------------
# kernels
# ---------------------------------------------------------------------------------------------------------------------
def square_distance(Xs, X2s=None):
    """gpflow.utilities.ops.square_distance on already-scaled inputs."""
    if X2s is None:
        n = np.sum(np.square(Xs), axis=-1, keepdims=True)
        d = -2.0 * (Xs @ Xs.T)
        d += n + n.T
        return d
    n1 = np.sum(np.square(Xs), axis=-1)
    n2 = np.sum(np.square(X2s), axis=-1)
    d = -2.0 * (Xs @ X2s.T)
    d += n1[:, None] + n2[None, :]
    return d


def k_of_r2(kernel, r2, variance):
    """K from scaled squared distance.  Returns (K, r, clipped_mask)."""
    if kernel == "SquaredExponential":
        return variance * np.exp(-0.5 * r2), None
    r = np.sqrt(np.maximum(r2, R2_CLIP))
    if kernel == "Matern52":
        s5 = np.sqrt(5.0)
        return variance * (1.0 + s5 * r + 5.0 / 3.0 * np.square(r)) * np.exp(-s5 * r), r
    if kernel == "Matern32":
        s3 = np.sqrt(3.0)
        return variance * (1.0 + s3 * r) * np.exp(-s3 * r), r
    if kernel == "Matern12":
        return variance * np.exp(-r), r
    raise ValueError(kernel)


def kern(kernel, X, X2, h: Hyper):
    Xs = X / h.lengthscales
    X2s = None if X2 is None else X2 / h.lengthscales
    r2 = square_distance(Xs, X2s)
    return k_of_r2(kernel, r2, h.variance)[0]


# ---------------------------------------------------------------------------------------------------------------------
# log marginal likelihood and analytic gradient (what TF autodiff produces, written in closed form)
# ---------------------------------------------------------------------------------------------------------------------
def lml(kernel, X, y, h: Hyper):
    """GPR.log_marginal_likelihood: y is [N,1]."""
    N = X.shape[0]
    K = kern(kernel, X, None, h)
    Ky = K + h.noise_variance * np.eye(N)
    L = np.linalg.cholesky(Ky)
    resid = y[:, 0] - (h.mean_c if h.has_mean else 0.0)
    a = sla.solve_triangular(L, resid, lower=True)
    return float(-0.5 * a @ a - 0.5 * N * np.log(2.0 * np.pi) - np.sum(np.log(np.diag(L))))


def neg_lml_and_grad(kernel, X, y, u, n_ls, has_mean):
    """f(u) = -LML and d f / d u over the packed unconstrained vector (the closure handed to L-BFGS-B)."""
    h = Hyper.unpack(u, n_ls, has_mean)
    N, d = X.shape
    ls = h.lengthscales
    Xs = X / ls
    r2 = square_distance(Xs, None)
    K, r = k_of_r2(kernel, r2, h.variance)
    Ky = K + h.noise_variance * np.eye(N)
    L = np.linalg.cholesky(Ky)
    resid = y[:, 0] - (h.mean_c if has_mean else 0.0)
    a = sla.solve_triangular(L, resid, lower=True)
    alpha = sla.solve_triangular(L, a, lower=True, trans="T")
    val = -0.5 * a @ a - 0.5 * N * np.log(2.0 * np.pi) - np.sum(np.log(np.diag(L)))
    Linv = sla.solve_triangular(L, np.eye(N), lower=True)
    Kinv = Linv.T @ Linv
    W = np.outer(alpha, alpha) - Kinv  # dLML/dK = 0.5 * W

    # radial factor g such that dK/d(ls_j) = g * Delta_j^2 / ls_j^3   (zero where r2 was clipped)
    if kernel == "SquaredExponential":
        g = K
    else:
        live = r2 > R2_CLIP
        if kernel == "Matern52":
            s5 = np.sqrt(5.0)
            g = (5.0 / 3.0) * h.variance * (1.0 + s5 * r) * np.exp(-s5 * r)
        elif kernel == "Matern32":
            s3 = np.sqrt(3.0)
            g = 3.0 * h.variance * np.exp(-s3 * r)
        else:  # Matern12
            g = K / r
        g = np.where(live, g, 0.0)
    WG = W * g
    if n_ls > 1:
        d_ls = np.empty(n_ls)
        for j in range(n_ls):
            dj = X[:, j][:, None] - X[:, j][None, :]
            d_ls[j] = 0.5 * np.sum(WG * dj * dj) / ls[j] ** 3
    else:
        # scalar lengthscale: sum_j Delta_j^2 / ls^3 = r2 * ls^2 / ls^3 = r2 / ls  (r2 un-clipped, >= 0 enforced)
        d_ls = np.array([0.5 * np.sum(WG * np.maximum(r2, 0.0)) / ls[0]])
    # TF autodiff gives 0.5*sum(W*k_unit) * sigmoid(u); when softplus(u) underflows to 0 the chain factor is 0 as well, so the
    # reference's gradient entry is exactly 0 there (not 0/0) -- L-BFGS-B line searches do visit such points (config C5).
    d_var = 0.5 * np.sum(W * K) / h.variance if h.variance > 0.0 else 0.0
    d_noise = 0.5 * np.trace(W)
    grad_theta = list(d_ls) + [d_var, d_noise]
    chain = list(sigmoid(u[:n_ls])) + [sigmoid(u[n_ls]), sigmoid(u[n_ls + 1])]
    if has_mean:
        grad_theta.append(np.sum(alpha))
        chain.append(1.0)
    grad_u = -np.array(grad_theta) * np.array(chain)
    return -float(val), grad_u


# ---------------------------------------------------------------------------------------------------------------------
# posterior prediction
# ---------------------------------------------------------------------------------------------------------------------
def predict_y(kernel, X, y, h: Hyper, Xnew, chunk=65536):
    """GPR.predict_f + Gaussian.predict_mean_and_var in GPflow's op order.  Returns (mean[M,1], var[M,1])."""
    N = X.shape[0]
    M = Xnew.shape[0]
    c = h.mean_c if h.has_mean else 0.0
    Kmm = kern(kernel, X, None, h) + h.noise_variance * np.eye(N)
    Lm = np.linalg.cholesky(Kmm)  # GPflow factorises on every predict call
    err = y[:, 0] - c
    mean = np.empty((M, 1))
    var = np.empty((M, 1))
    for s in range(0, M, chunk):
        e = min(M, s + chunk)
        Kmn = kern(kernel, X, Xnew[s:e], h)
        A = sla.solve_triangular(Lm, Kmn, lower=True)
        fvar = h.variance - np.sum(np.square(A), axis=0)
        A = sla.solve_triangular(Lm, A, lower=True, trans="T")
        fmean = A.T @ err
        mean[s:e, 0] = fmean + c
        var[s:e, 0] = fvar + h.noise_variance
    return mean, var


def predict_y_longdouble(kernel, X, y, h: Hyper, Xnew):
    """Extended-precision (x87 80-bit) posterior for error attribution; small problems only (O(N^3) python loops)."""
    ld = np.longdouble
    X = X.astype(ld)
    Xn = Xnew.astype(ld)
    ls = h.lengthscales.astype(ld)
    N = X.shape[0]

    def kk(A, B):
        D = (A[:, None, :] - B[None, :, :]) / ls
        r2 = np.sum(D * D, axis=-1)
        v = ld(h.variance)
        if kernel == "SquaredExponential":
            return v * np.exp(-r2 / 2)
        r = np.sqrt(np.maximum(r2, ld(R2_CLIP)))
        if kernel == "Matern52":
            s5 = np.sqrt(ld(5))
            return v * (1 + s5 * r + ld(5) / 3 * r * r) * np.exp(-s5 * r)
        if kernel == "Matern32":
            s3 = np.sqrt(ld(3))
            return v * (1 + s3 * r) * np.exp(-s3 * r)
        return v * np.exp(-r)

    K = kk(X, X) + ld(h.noise_variance) * np.eye(N, dtype=ld)
    L = np.zeros((N, N), dtype=ld)
    for j in range(N):
        s = K[j, j] - np.dot(L[j, :j], L[j, :j])
        L[j, j] = np.sqrt(s)
        for i in range(j + 1, N):
            L[i, j] = (K[i, j] - np.dot(L[i, :j], L[j, :j])) / L[j, j]
    Ks = kk(X, Xn)
    c = ld(h.mean_c if h.has_mean else 0.0)
    err = y[:, 0].astype(ld) - c

    def fsolve(B):
        B = B.copy()
        for i in range(N):
            B[i] = (B[i] - L[i, :i] @ B[:i]) / L[i, i]
        return B

    def bsolve(B):
        B = B.copy()
        for i in range(N - 1, -1, -1):
            B[i] = (B[i] - L[i + 1:, i] @ B[i + 1:]) / L[i, i]
        return B

    A = fsolve(Ks)
    fvar = ld(h.variance) - np.sum(A * A, axis=0)
    alpha = bsolve(fsolve(err))
    fmean = Ks.T @ alpha + c
    return fmean[:, None], (fvar + ld(h.noise_variance))[:, None]


def ucb_argmax(mean, var, varsigma=VARSIGMA_DEFAULT):
    """gp_surrogate.py:326-328.  Returns (index, mean*, var*, ucb*); np.argmax = first maximum (first NaN if any)."""
    ucb = mean + varsigma * var
    best = int(np.argmax(ucb))
    return best, float(np.ravel(mean)[best]), float(np.ravel(var)[best]), float(np.ravel(ucb)[best])


# ---------------------------------------------------------------------------------------------------------------------
# model object: the role gpflow.models.GPR plays behind GPSurrogate.gpflow_model
# ---------------------------------------------------------------------------------------------------------------------
class OracleGPR:
    """
    Stand-in for gpflow.models.GPR (gp_surrogate.py:490-495): holds data + hyper-parameters, warm-starts across fits.
    kernel: one of KERNEL_IDS; lengthscales scalar or [d]; mean_c None (zero mean) or float (Constant).
    """

    def __init__(self, X, y, kernel="Matern52", lengthscales=1.0, variance=1.0, noise_variance=1.0, mean_c=None):
        assert kernel in KERNEL_IDS
        self.kernel = kernel
        self.h = Hyper(lengthscales, variance, noise_variance, mean_c)
        self.n_ls = self.h.lengthscales.size
        self.set_data(X, y)
        self.n_evals = 0
        self.last_result = None

    def set_data(self, X, y):
        X = np.ascontiguousarray(X, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        assert X.ndim == 2 and y.ndim == 2 and y.shape == (X.shape[0], 1)
        self.X, self.y = X, y

    data = property(lambda self: (self.X, self.y), lambda self, v: self.set_data(*v))

    def training_loss(self, u=None):
        u = self.h.pack() if u is None else u
        return neg_lml_and_grad(self.kernel, self.X, self.y, u, self.n_ls, self.h.has_mean)

    def log_marginal_likelihood(self):
        return lml(self.kernel, self.X, self.y, self.h)

    def fit(self, **scipy_kwargs):
        """gpflow.optimizers.Scipy().minimize(model.training_loss, model.trainable_variables) with SciPy defaults."""
        u0 = self.h.pack()

        def fun(u):
            self.n_evals += 1
            return neg_lml_and_grad(self.kernel, self.X, self.y, u, self.n_ls, self.h.has_mean)

        res = sopt.minimize(fun, u0, jac=True, method="L-BFGS-B", **scipy_kwargs)
        self.h = Hyper.unpack(res.x, self.n_ls, self.h.has_mean)
        self.last_result = res
        return res

    def predict_y(self, Xnew):
        Xnew = np.ascontiguousarray(Xnew, dtype=np.float64)
        return predict_y(self.kernel, self.X, self.y, self.h, Xnew)

    def parameter_dict(self):
        d = {
            ".kernel.lengthscales": self.h.lengthscales.copy() if self.n_ls > 1 else float(self.h.lengthscales[0]),
            ".kernel.variance": self.h.variance,
            ".likelihood.variance": self.h.noise_variance,
        }
        if self.h.has_mean:
            d[".mean_function.c"] = self.h.mean_c
        return d
