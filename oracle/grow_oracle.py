"""
ORACLE -- TEST INFRASTRUCTURE ONLY (see gpr_oracle.py header for who may import this).

CPU restatements of ``LeafNode.grow(depth)`` (reference gpso/param_space.py:175-200) built on ``ternary_split``
(:257-307) and ``get_center_as_list`` (:202-217):

  * ``grow_literal``   -- per-node pure-Python floats, the reference's arithmetic statement by statement (small cases);
  * ``grow_by_level``  -- the same arithmetic vectorised per tree level with numpy (IEEE fp64 element-wise ops, no FMA),
                          used to check the CUDA generator at depth 12 in seconds.
Both return the ``[(3^depth - 1)/2, d]`` centre matrix in the reference's row order.
"""
import numpy as np


def _split(bounds):
    """bounds: list of (lo, hi) python numbers -> three child bound lists [l, c, r] (param_space.py:272-301)."""
    widths = [b[1] - b[0] for b in bounds]
    k = int(np.argmax(widths))
    delta = widths[k] / 3
    cuts = [bounds[k][0] + i * delta for i in range(4)]
    children = []
    for i in range(3):
        child = list(bounds)
        child[k] = (cuts[i], cuts[i + 1])
        children.append(child)
    return children


def grow_literal(bounds, depth):
    level = [[(b[0], b[1]) for b in bounds]]
    rows = []
    for _ in range(depth):
        rows.append(np.array([[np.mean(b) for b in node] for node in level]))
        level = [child for node in level for child in _split(node)]
    return np.vstack(rows)


def grow_by_level(bounds, depth):
    bounds = np.asarray(bounds, dtype=np.float64)
    lo = bounds[None, :, 0].copy()
    hi = bounds[None, :, 1].copy()
    rows = []
    for _ in range(depth):
        rows.append((lo + hi) / 2.0)
        n, d = lo.shape
        w = hi - lo
        k = np.argmax(w, axis=1)  # first maximum, like np.argmax on the python list
        ar = np.arange(n)
        delta = w[ar, k] / 3.0
        base = lo[ar, k]
        new_lo = np.repeat(lo, 3, axis=0)
        new_hi = np.repeat(hi, 3, axis=0)
        for i in range(3):
            new_lo[i::3][ar, k] = base + float(i) * delta
            new_hi[i::3][ar, k] = base + float(i + 1) * delta
        lo, hi = new_lo, new_hi
    return np.vstack(rows)
