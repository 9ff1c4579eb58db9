import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pygpso_b200 import backend
cuda = backend.default_backend()
d, depth = 10, 12
bounds = np.array([[0.0, 1.0/3.0]] + [[0.0, 1.0]]*(d-1))
for _ in range(3):
    out = cuda.grow_leaves(bounds, depth)
print(out.shape)
