#!/usr/bin/env python
"""Deviation of the int8 tcgen05 engine from the FP64 DMMA engine (and from the oracle on a sample) at config C3 shape.

    python tools/engine_error.py [--N 4096 --d 10 --M 200000]

Prints, per digit count S, max |var_int8 - var_dmma| / (1e-8 * variance) over M random candidates plus the training
points themselves (the cancellation case), and the library's own estimate."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygpso_b200 import backend  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=4096)
    ap.add_argument("--d", type=int, default=10)
    ap.add_argument("--M", type=int, default=200000)
    ap.add_argument("--noise", type=float, default=1e-3)
    ap.add_argument("--ls", type=float, default=0.0)
    ap.add_argument("--variance", type=float, default=1.0)
    args = ap.parse_args()
    rng = np.random.default_rng(20240517)
    X = rng.random((args.N, args.d))
    y = np.sin(3 * X.sum(1)) + 0.01 * rng.standard_normal(args.N)
    theta = np.array([args.ls or 0.25 * np.sqrt(args.d), args.variance, args.noise, 0.0])
    Xc = np.concatenate([np.random.default_rng(1).random((args.M, args.d)), X])
    cuda = backend.default_backend()
    ref = cuda.open_session("Matern52", 1, True)
    ref.set_predict_mode(1, 0)
    ref.set_data(X, y[:, None])
    ref.factorize(theta)
    mean_r, var_r = ref.predict_y(Xc)
    tol = 1e-8 * theta[1]
    for S in (5, 6, 7, 8):
        s = cuda.open_session("Matern52", 1, True)
        s.set_predict_mode(2, S)
        s.set_data(X, y[:, None])
        s.factorize(theta)
        info = s.predict_info()
        mean, var = s.predict_y(Xc)
        dv = np.abs(var - var_r)
        print(f"S={S}: max|dvar|/tol = {dv.max() / tol:.3e} (random cands {dv[:args.M].max() / tol:.3e}, at training points "
              f"{dv[args.M:].max() / tol:.3e}), rms {np.sqrt((dv ** 2).mean()) / tol:.3e}; library estimate {info['error_estimate_over_tol']:.3e}; "
              f"same argmax: {int(np.argmax(mean + 1.82 * var)) == int(np.argmax(mean_r + 1.82 * var_r))}")
        s.close()
    ref.close()


if __name__ == "__main__":
    main()
