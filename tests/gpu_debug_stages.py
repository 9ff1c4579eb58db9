"""Stage-by-stage diagnostics of the CUDA path against the oracle + first timings (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import gpr_oracle as go
from pygpso_b200 import backend

def synthetic(N, d, seed=20240517):
    rng = np.random.default_rng(seed)
    X = rng.random((N, d)); y = np.sin(3 * X.sum(1)) + 0.01 * rng.standard_normal(N)
    return X, y[:, None]

cuda = backend.default_backend()
for N, d in [(6, 2), (129, 3), (300, 5), (1000, 10)]:
    X, y = synthetic(N, d)
    h = go.Hyper(0.25 * np.sqrt(d), 1.3, 1e-3, 0.1)
    s = cuda.open_session("Matern52", 1, True); s.set_data(X, y)
    theta = np.array([h.lengthscales[0], h.variance, h.noise_variance, h.mean_c])
    try:
        s.factorize(theta)
    except Exception as e:
        print(N, d, "factorize FAILED", repr(e)); continue
    K = go.kern("Matern52", X, None, h) + h.noise_variance * np.eye(N)
    Lr = np.linalg.cholesky(K)
    L = s.debug_fetch(1); Li = s.debug_fetch(2); al = s.debug_fetch(3)
    print(f"N={N} d={d}: |L-Lref|={np.abs(L-Lr).max():.2e} |Linv L - I|={np.abs(Li@Lr-np.eye(N)).max():.2e} "
          f"|alpha-ref|={np.abs(al-np.linalg.solve(K, y[:,0]-h.mean_c)).max():.2e} lml={s.log_marginal_likelihood():.12g} ref={go.lml('Matern52',X,y,h):.12g}")
    u = h.pack()
    f, g = s.neg_lml_and_grad(u); fr, gr = go.neg_lml_and_grad("Matern52", X, y, u, 1, True)
    print(f"   nlml {f:.12g} ref {fr:.12g}  grad {g} ref {gr}")
    s.factorize(theta)
    Xc = np.random.default_rng(1).random((777, d))
    m, v = s.predict_y(Xc); mr, vr = go.predict_y("Matern52", X, y, h, Xc)
    print(f"   predict: |dmean|={np.abs(m-mr[:,0]).max():.2e} |dvar|={np.abs(v-vr[:,0]).max():.2e}  argmax {s.ucb_argmax(Xc, go.VARSIGMA_DEFAULT)} ref {go.ucb_argmax(mr, vr)}")
    s.close()

print("---- timings")
for N, d in [(512, 2), (4096, 10), (8192, 20)]:
    X, y = synthetic(N, d)
    s = cuda.open_session("Matern52", 1, True); s.set_data(X, y)
    theta = np.array([0.25 * np.sqrt(d), 1.0, 1e-3, 0.0])
    s.factorize(theta)
    t = time.perf_counter(); s.factorize(theta); t1 = time.perf_counter() - t
    u = go.Hyper(theta[0], theta[1], theta[2], theta[3]).pack()
    s.neg_lml_and_grad(u)
    t = time.perf_counter(); f, g = s.neg_lml_and_grad(u); t2 = time.perf_counter() - t
    print(f"N={N} d={d}: factorize {t1*1e3:.2f} ms   neg_lml_grad {t2*1e3:.2f} ms (device {s.last_timing_ms()[0]:.2f} ms)  launches {s.launch_count()}  f={f:.10g}")
    s.factorize(theta)
    M = 100_000 if N <= 4096 else 20000
    Xc = np.random.default_rng(2).random((M, d))
    s.ucb_argmax(Xc[:1000], go.VARSIGMA_DEFAULT)
    t = time.perf_counter(); r = s.ucb_argmax(Xc, go.VARSIGMA_DEFAULT); t3 = time.perf_counter() - t
    dev = s.last_timing_ms()[0]
    print(f"   ucb_argmax M={M}: host {t3*1e3:.1f} ms device {dev:.1f} ms -> {M/dev*1e3:.3e} cand/s, {M*N*N/dev*1e-9:.2f} TFLOP/s(N^2 per cand)  result {r}")
    s.close()
