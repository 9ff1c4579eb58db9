"""LML+grad closure timing at given shapes (run on the GPU box; under `ncu --metrics gpu__time_duration.sum` it yields
the per-kernel launch list of one evaluation).  usage: python tools/lml_profile.py N d [evals] [--factorize] [--no-hybrid] [--stepwise]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from pygpso_b200 import backend


def synthetic(N, d, seed=20240517):
    rng = np.random.default_rng(seed)
    X = rng.random((N, d))
    y = np.sin(3 * X.sum(1)) + 0.01 * rng.standard_normal(N)
    return X, y[:, None]


def softplus_inv(v):
    return np.log(np.expm1(v))


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    N, d = int(args[0]), int(args[1])
    evals = int(args[2]) if len(args) > 2 else 3
    X, y = synthetic(N, d)
    cuda = backend.default_backend()
    s = cuda.open_session("Matern52", 1, True)
    s.set_data(X, y)
    if "--stepwise" in sys.argv:
        s.set_factor_mode(False)
    if "--no-hybrid" in sys.argv:
        s.set_factor_mode(True, hybrid=False)
    if "--kinv-dmma" in sys.argv:
        s.set_kinv_mode(1)
    if "--kinv-int8" in sys.argv:
        s.set_kinv_mode(2)
    if "--inv-dmma" in sys.argv:
        s.set_inverse_mode(1)
    if "--inv-int8" in sys.argv:
        s.set_inverse_mode(2)
    theta = np.array([0.25 * np.sqrt(d), 1.0, 1e-3, 0.0])
    u = np.array([softplus_inv(theta[0]), softplus_inv(theta[1]), softplus_inv(theta[2] - 1e-6), 0.0])
    if "--factorize" in sys.argv:
        s.factorize(theta)
        t = time.perf_counter()
        for _ in range(evals):
            s.factorize(theta)
        print(f"N={N} d={d}: factorize {(time.perf_counter() - t) / evals * 1e3:.3f} ms/call")
    else:
        f, g = s.neg_lml_and_grad(u)
        dev = 0.0
        t = time.perf_counter()
        for i in range(evals):
            f, g = s.neg_lml_and_grad(u + 1e-3 * (i + 1))
            dev += s.last_timing_ms()[0]
        wall = (time.perf_counter() - t) / evals * 1e3
        dev /= evals
        print(f"N={N} d={d}: neg_lml_grad wall {wall:.3f} ms device {dev:.3f} ms -> {N ** 3 / dev * 1e-9:.2f} TFLOP/s (N^3), "
              f"launches/eval {s.launch_count() // (evals + 1)} {s.factor_info()}  f={f:.12g} g={g}")
    s.close()


if __name__ == "__main__":
    main()
