"""LML+grad timing before / after a scoring phase has configured the persisting-L2 set-aside (same process, same GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from pygpso_b200 import backend

cuda = backend.CudaBackend(device=0)
def lml(N, d, tag):
    X, y = bench.synthetic_training(N, d)
    s = cuda.open_session("Matern52", 1, True); s.set_data(X, y)
    u = bench.pack_unconstrained(0.25 * np.sqrt(d), 1.0, 1e-3, 0.0)
    s.neg_lml_and_grad(u)
    ms = 0.0
    for i in range(5):
        s.neg_lml_and_grad(u + 1e-3 * (i + 1)); ms += s.last_timing_ms()[0]
    print(f"{tag}: N={N} LML+grad {ms / 5:.3f} ms"); s.close()
lml(4096, 10, "fresh process"); lml(8192, 20, "fresh process")
X, y = bench.synthetic_training(4096, 10)
p = cuda.open_session("Matern52", 1, True); p.set_data(X, y); p.factorize(bench.fixed_theta(10))
Xc = np.random.default_rng(0).random((200000, 10))
print("scoring", p.ucb_argmax(Xc, bench.VARSIGMA)[0])
lml(4096, 10, "after a scoring phase (set-aside released by the fit)"); lml(8192, 20, "after a scoring phase")
print("scoring again", p.ucb_argmax(Xc, bench.VARSIGMA)[0])
p.close()
