"""cProfile of config C5 (10-D Rastrigin, 500 evaluations, depth-12 exploration) through the public API on one GPU."""
import cProfile, pstats, io, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pygpso_b200 import GPRSurrogate, GPSOptimiser, ParameterSpace

def rastrigin(point):
    x = np.asarray(point)
    return -float(10 * x.size + np.sum(x * x - 10 * np.cos(2 * np.pi * x)))

space = ParameterSpace(parameter_names=[f"p{i}" for i in range(10)], parameter_bounds=[[-5.12, 5.12]] * 10)
opt = GPSOptimiser(parameter_space=space, exploration_method="tree", exploration_depth=12, budget=500, stopping_condition="evaluations")
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
opt.run(rastrigin)
pr.disable()
print("wall", time.perf_counter() - t0)
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
