"""Host logic of the partition tree (mirrors the reference's tests/test_param_space.py) -- CPU only."""
import numpy as np
import pytest
from sklearn.preprocessing import MinMaxScaler

from oracle import grow_oracle
from pygpso_b200.param_space import LeafNode, ParameterSpace, PreOrderIter
from pygpso_b200.utils import PointLabels

BOUNDS = [[-3, 5], [-3, 3], [2.0, 12.0]]
NAMES = ["x", "y", "z"]


def make_space():
    return ParameterSpace(parameter_bounds=BOUNDS, parameter_names=NAMES)


def test_init_and_scaler():
    space = make_space()
    assert isinstance(space, LeafNode) and isinstance(space.scaler, MinMaxScaler)
    assert space.ndim == 3 and space.depth == 0 and space.max_depth == 0 and space.is_leaf and space.is_root
    np.testing.assert_equal(space.scaler.data_min_, [-3, -3, 2])
    np.testing.assert_equal(space.scaler.data_max_, [5, 3, 12])
    assert space.norm_bounds == [(0, 1)] * 3
    assert space.name == "full_domain" and space.label == PointLabels.not_assigned and space.sampled is False
    assert space.get_center_as_list(normed=True) == [0.5, 0.5, 0.5]
    assert space.get_center_as_list(normed=False) == [1.0, 0.0, 7.0]
    assert space.get_center_as_dict(normed=False) == {"x": 1.0, "y": 0.0, "z": 7.0}


def test_bad_bounds_rejected():
    with pytest.raises(AssertionError):
        ParameterSpace(parameter_bounds=[[1, 0]], parameter_names=["a"])
    with pytest.raises(AssertionError):
        ParameterSpace(parameter_bounds=[[0, 1]], parameter_names=["a", "b"])


def test_normalise_round_trip_exact():
    space = make_space()
    rng = np.random.default_rng(0)
    orig = rng.random((50, 3)) * [8, 6, 10] + [-3, -3, 2]
    normed = space.normalise_coords(orig)
    assert normed.min() >= 0 and normed.max() <= 1
    np.testing.assert_allclose(space.denormalise_coords(normed), orig, rtol=0, atol=1e-14)
    unit = rng.random((20, 3))
    np.testing.assert_equal(space.normalise_coords(space.denormalise_coords(unit)).shape, unit.shape)


def test_ternary_split_bounds_and_order():
    space = make_space()
    kids = space.ternary_split()
    assert [k.name for k in kids] == ["full_domain->l", "full_domain->c", "full_domain->r"]
    assert space.children == tuple(kids) and all(k.parent is space and k.depth == 1 for k in kids)
    # first widest dimension (all equal -> dim 0) is cut into thirds with the reference's arithmetic
    third = 1 / 3
    assert kids[0].norm_bounds[0] == (0 + 0 * third, 0 + 1 * third)
    assert kids[1].norm_bounds[0] == (0 + 1 * third, 0 + 2 * third)
    assert kids[2].norm_bounds[0] == (0 + 2 * third, 0 + 3 * third)
    assert all(k.norm_bounds[1:] == [(0, 1), (0, 1)] for k in kids)
    np.testing.assert_array_equal(kids[1].center_array(), space.center_array())
    # next split goes to dim 1
    grand = kids[0].ternary_split()
    assert grand[0].norm_bounds[1] == (0.0, third) and space.max_depth == 2
    assert space[0][2] is grand[2]


def test_preorder_and_best_leaf():
    space = make_space()
    kids = space.ternary_split()
    grand = kids[1].ternary_split()
    order = [n.name for n in PreOrderIter(space)]
    assert order == ["full_domain", "full_domain->l", "full_domain->c", "full_domain->c->l", "full_domain->c->c",
                     "full_domain->c->r", "full_domain->r"]
    kids[0].score, kids[1].score, kids[2].score = 1.0, 3.0, 3.0
    assert space.get_best_score_leaf(depth=1) is kids[1]  # tie -> first in pre-order
    kids[1].sampled = True
    assert space.get_best_score_leaf(depth=1) is kids[2]
    assert space.get_best_score_leaf(depth=1, only_not_sampled=False) is kids[1]
    assert space.get_best_score_leaf(depth=2) is grand[0]
    assert space.get_best_score_leaf(depth=7) is None


def test_save_load_round_trip(tmp_path):
    space = make_space()
    kids = space.ternary_split()
    kids[2].ternary_split()
    kids[0].score, kids[0].label, kids[0].sampled = 2.5, PointLabels.evaluated, True
    path = str(tmp_path / "tree")
    space.save(path)
    loaded = ParameterSpace.from_file(path)
    assert isinstance(loaded, ParameterSpace) and loaded.max_depth == 2
    for a, b in zip(PreOrderIter(space), PreOrderIter(loaded)):
        for attr in LeafNode.COMPARE_ATTRS:
            assert getattr(a, attr) == getattr(b, attr), attr
    np.testing.assert_equal(loaded.scaler.data_max_, space.scaler.data_max_)


def test_sample_uniformly_inside_leaf():
    kid = make_space().ternary_split()[2]
    pts = kid.sample_uniformly(100, seed=3)
    assert pts.shape == (100, 3)
    lo = np.array([b[0] for b in kid.norm_bounds])
    hi = np.array([b[1] for b in kid.norm_bounds])
    assert np.all(pts >= lo) and np.all(pts <= hi)
    np.testing.assert_array_equal(pts, kid.sample_uniformly(100, seed=3))


@pytest.mark.parametrize("d,depth", [(2, 5), (3, 6), (5, 7)])
def test_grow_oracles_agree(d, depth):
    """The vectorised level-by-level restatement equals the literal per-node python arithmetic bit for bit."""
    space = ParameterSpace(parameter_bounds=[[0, 1]] * d, parameter_names=[f"p{i}" for i in range(d)])
    child = space.ternary_split()[0].ternary_split()[2]
    literal = grow_oracle.grow_literal(child.norm_bounds, depth)
    assert literal.shape == ((3 ** depth - 1) // 2, d)
    np.testing.assert_array_equal(literal, grow_oracle.grow_by_level(child.norm_bounds, depth))
    # row 0 is the leaf's own centre; rows 1..3 its children l, c, r
    np.testing.assert_array_equal(literal[0], child.center_array())
    kids = child.ternary_split()
    np.testing.assert_array_equal(literal[1:4], np.array([k.center_array() for k in kids]))
    np.testing.assert_array_equal(literal[2], literal[0])  # the middle child repeats its parent's centre


def test_grow_split_dimension_varies_per_node():
    """Appendix C of SURVEY.md: sibling widths differ by ulps, so the split dimension is a per-node decision."""
    rows = grow_oracle.grow_by_level([(0, 1)] * 3, 5)
    level3, level4 = rows[13:40], rows[40:121]
    changed = {int(np.flatnonzero(level4[3 * i] != level3[i])[0]) for i in range(27)}
    assert len(changed) > 1


def test_incremental_depth_index_equals_a_fresh_preorder_walk():
    """Splitting leaves in random order: the per-depth lists maintained by insertion must equal the lists a fresh pre-order
    traversal gives (this order decides ties in get_best_score_leaf), also after detaching a subtree."""
    from pygpso_b200.param_space import PreOrderIter

    rng = np.random.default_rng(3)
    space = ParameterSpace(parameter_names=["a", "b", "c"], parameter_bounds=[[0, 1], [-1, 1], [2, 5]])

    def fresh():
        by_depth = {}
        for node in PreOrderIter(space):
            by_depth.setdefault(node.depth, []).append(node)
        return by_depth

    assert space.max_depth == 0
    for step in range(120):
        leaves = [n for n in PreOrderIter(space) if n.is_leaf]
        leaf = leaves[rng.integers(len(leaves))]
        for child in leaf.ternary_split():
            child.score = float(rng.integers(4))  # many ties
        got, want = space._depth_index(), fresh()
        assert set(got) == set(want)
        for depth in want:
            assert [id(n) for n in got[depth]] == [id(n) for n in want[depth]]
        for depth in want:
            cands = [n for n in want[depth] if not n.sampled]
            best = space.get_best_score_leaf(depth)
            if cands:
                top = max(n.score for n in cands)
                assert best is next(n for n in cands if n.score == top)
        if step == 60:  # a structural change that is not an append: the index must rebuild itself
            victim = next(n for n in PreOrderIter(space) if n.depth == 2 and not n.is_leaf)
            victim.parent = None
            got, want = space._depth_index(), fresh()
            for depth in want:
                assert [id(n) for n in got[depth]] == [id(n) for n in want[depth]]


def test_fast_scaler_paths_are_bit_identical_to_sklearn():
    """normalise / denormalise skip sklearn's per-call validation: same two array operations, same bits; everything that is
    not a finite float64 matrix still goes through sklearn itself."""
    from pygpso_b200.param_space import _scaler_transform

    rng = np.random.default_rng(0)
    for d in (1, 2, 10, 33):
        bounds = np.sort(rng.normal(scale=50.0, size=(d, 2)), axis=1)
        bounds[:, 1] += 1e-3
        space = ParameterSpace(parameter_bounds=bounds.tolist(), parameter_names=[f"p{i}" for i in range(d)])
        for n in (1, 7, 500):
            x = rng.normal(scale=30.0, size=(n, d))
            u = rng.random((n, d))
            assert np.array_equal(space.normalise_coords(x), space.scaler.transform(x))
            assert np.array_equal(space.denormalise_coords(u), space.scaler.inverse_transform(u))
            assert space.normalise_coords(x) is not x  # a copy, like sklearn's
        # inputs outside the fast path: float32, integers, NaN
        x32 = rng.random((4, d)).astype(np.float32)
        assert np.array_equal(_scaler_transform(space.scaler, x32, inverse=True), space.scaler.inverse_transform(x32))
        xi = np.arange(3 * d).reshape(3, d)
        assert np.array_equal(_scaler_transform(space.scaler, xi, inverse=False), space.scaler.transform(xi))
        xn = rng.random((3, d))
        xn[1, 0] = np.nan
        assert np.array_equal(space.denormalise_coords(xn), space.scaler.inverse_transform(xn), equal_nan=True)
        with pytest.raises(ValueError):
            space.denormalise_coords(np.full((2, d), np.inf))
