"""A/B of the 3-digit screening product as CTA pairs (cta_group::2) vs single CTAs at config C3.

    python tools/screen_pair_ab.py <pair 0|1> <overlap 0|1> <candidates> <host 0|1>
"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import torch
from pygpso_b200 import backend

pair, overlap, M, host = int(sys.argv[1]), int(sys.argv[2]), int(float(sys.argv[3])), int(sys.argv[4])
N, d = 4096, 10
X, y = bench.synthetic_training(N, d)
theta = bench.fixed_theta(d)
cuda = backend.CudaBackend(device=0)
s = cuda.open_session("Matern52", 1, True)
s.set_data(X, y)
s.set_screen_mode(3)
s.set_screen_pair(bool(pair))
s.set_overlap(bool(overlap))
s.factorize(theta)
xc = torch.empty((M, d), dtype=torch.float64)
bench.fill_candidates(xc.numpy(), 0, M, 10_000_000, d)
xd = xc.cuda()
stream = torch.cuda.current_stream().cuda_stream
print("start", pair, overlap, M, host, flush=True)
for i in range(3):
    t0 = time.perf_counter()
    r = s.ucb_argmax(xc.numpy(), bench.VARSIGMA) if host else s.ucb_argmax_dev(xd.data_ptr(), M, bench.VARSIGMA, stream)
    torch.cuda.synchronize()
    info = s.screen_info()
    print("call", i, round((time.perf_counter() - t0) * 1e3, 2), "ms", r[0], info["path"], info["survivors"], "prod_ms", round(info["screen_product_ms"], 2), "win", info["screen_windows"], flush=True)
s.close()
