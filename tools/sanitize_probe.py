#!/usr/bin/env python
"""
Small end-to-end pass over the kernels added in round 2, meant to run under compute-sanitizer (memcheck / racecheck):
forced hybrid factorisation with one- and two-tile leaves, the strip-split panel and the in-kernel inverse merges, the
one-round diagonal transposes, the leaf generator at d = 64, the screening ladder (2-digit all-pairs rung, 3-digit rung,
refine pass) and the pooled handle memory (a second session on recycled blocks).

    compute-sanitizer --tool memcheck python tools/sanitize_probe.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygpso_b200 import backend  # noqa: E402


def main():
    cuda = backend.default_backend()
    rng = np.random.default_rng(0)
    for N, d in ((300, 3), (640, 4)):
        X = rng.random((N, d))
        y = np.sin(3 * X.sum(1))[:, None]
        for hybrid in (None, True):
            s = cuda.open_session("Matern52", 1, True)
            s.set_data(X, y)
            s.set_factor_mode(True, hybrid=hybrid)
            f, g = s.neg_lml_and_grad(np.array([0.0, 0.5, -6.0, 0.1]))
            s.factorize(np.array([0.5, 1.0, 1e-3, 0.0]))
            mean, var = s.predict_y(rng.random((500, d)))
            print(N, hybrid, s.factor_info(), f, float(mean[0]), float(var[0]))
            s.close()
    out = cuda.grow_leaves(np.array([[0.0, 1.0]] * 64), 5)
    print("leaves", out.shape, float(out.sum()))
    N, d, M = 1024, 4, 66_000
    X = rng.random((N, d))
    y = np.sin(3 * X.sum(1))[:, None]
    Xc = rng.random((M, d))
    for mode in (6, 3, 1):
        s = cuda.open_session("Matern52", 1, True)
        s.set_data(X, y)
        s.set_screen_mode(mode)
        s.factorize(np.array([0.4, 1.2, 1e-3, 0.1]))
        rec = s.ucb_argmax(Xc, 1.8)
        print("screen", mode, rec[0], s.screen_info()["path"], s.screen_info()["survivors"])
        s.close()
    # windows with side-stream overlap, the early verdict, the mean-bound level, top-k, both full-precision engines
    for mode, window in ((1, 8192), (5, 16384)):
        s = cuda.open_session("Matern52", 1, True)
        s.set_data(X, y)
        s.set_screen_mode(mode)
        s.factorize(np.array([0.4, 1.2, 1e-3, 0.1]))
        s.set_window(window)
        rec = s.ucb_argmax(Xc, 1.8)
        top = s.ucb_topk(Xc[:9000], 1.8, 5)
        print("windows", mode, rec[0], s.screen_info()["path"], s.screen_info()["screen_windows"], int(top[0, 0]))
        s.set_predict_mode(1, 0)  # FP64 DMMA engine
        s.factorize(np.array([0.4, 1.2, 1e-3, 0.1]))
        mean, var = s.predict_y(Xc[:3000])
        print("dmma", float(mean[0]), float(var[0]))
        s.close()
    s = cuda.open_session("SquaredExponential", d, True)  # ARD gradient, int8 inverse and K_y^-1 at nine tiles
    X9 = rng.random((1100, d))
    s.set_data(X9, np.sin(3 * X9.sum(1))[:, None])
    f, g = s.neg_lml_and_grad(np.array([0.0] * d + [0.5, -6.0, 0.1]))
    print("ard", f, g[:2])
    s.close()
    print("released", cuda.trim_pool())


if __name__ == "__main__":
    main()
