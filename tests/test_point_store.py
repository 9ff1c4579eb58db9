"""GPListOfPoints keeps the reference's de-duplication semantics (gpso/gp_surrogate.py:68-101) -- CPU only."""
import numpy as np

from pygpso_b200.gp_surrogate import DUPLICATE_TOLERANCE, GPListOfPoints, GPPoint
from pygpso_b200.utils import PointLabels


def point(coord, mu=0.0, label=PointLabels.gp_based, ucb=0.0):
    return GPPoint(normed_coord=np.asarray(coord, dtype=float), score_mu=mu, score_sigma=0.0, score_ucb=ucb, label=label)


def literal_append(items, obj):
    """The reference's loop, statement for statement in behaviour, on a plain list."""
    add = True
    for idx in range(len(items)):
        if np.linalg.norm(items[idx].normed_coord - obj.normed_coord) < DUPLICATE_TOLERANCE:
            add = False
            if items[idx].label == PointLabels.evaluated:
                continue
            items[idx] = obj
    if add:
        items.append(obj)


def test_append_matches_literal_semantics():
    rng = np.random.default_rng(1)
    base = rng.random((40, 3))
    fast, slow = GPListOfPoints(), []
    for step in range(400):
        coord = base[rng.integers(40)] + (rng.random(3) < 0.3) * rng.normal(0, 3e-13, 3)
        label = PointLabels.evaluated if rng.random() < 0.4 else PointLabels.gp_based
        obj = point(coord, mu=float(step), label=label)
        fast.append(obj)
        literal_append(slow, obj)
        assert len(fast) == len(slow)
    for a, b in zip(fast, slow):
        assert a == b and a.label == b.label and a.score_mu == b.score_mu


def test_evaluated_is_never_overwritten_and_gp_is_replaced():
    pts = GPListOfPoints()
    pts.append(point([0.5, 0.5], 1.0, PointLabels.gp_based))
    pts.append(point([0.5, 0.5], 2.0, PointLabels.gp_based))
    assert len(pts) == 1 and pts[0].score_mu == 2.0
    pts.append(point([0.5, 0.5], 3.0, PointLabels.evaluated))
    assert len(pts) == 1 and pts[0].label == PointLabels.evaluated and pts[0].score_mu == 3.0
    pts.append(point([0.5, 0.5], 4.0, PointLabels.gp_based))
    assert len(pts) == 1 and pts[0].score_mu == 3.0
    pts.append(point([0.5, 0.5 + 5e-13], 5.0, PointLabels.gp_based))  # inside the tolerance
    assert len(pts) == 1
    pts.append(point([0.5, 0.5 + 2e-12], 6.0, PointLabels.gp_based))  # outside
    assert len(pts) == 2


def test_find_by_coords_and_mutation_keeps_index_fresh():
    pts = GPListOfPoints([point([0.1 * i, 0.2]) for i in range(10)])
    assert pts.find_by_coords(np.array([0.3, 0.2])) is pts[3]
    assert pts.find_by_coords(np.array([0.35, 0.2])) is None
    assert pts.index_by_coords(np.array([0.9, 0.2])) == 9
    del pts[3]
    assert pts.find_by_coords(np.array([0.3, 0.2])) is None
    pts.insert(0, point([0.3, 0.2], 9.0))
    assert pts.find_by_coords(np.array([0.3, 0.2])).score_mu == 9.0
    pts[1] = point([0.77, 0.77])
    assert pts.find_by_coords(np.array([0.77, 0.77])) is pts[1]
    pts.clear()
    assert pts.find_by_coords(np.array([0.77, 0.77])) is None and len(pts) == 0


def test_save_load(tmp_path):
    pts = GPListOfPoints([point([0.1, 0.9], 1.5, PointLabels.evaluated, 0.3), point([0.4, 0.2], -2.0, PointLabels.gp_based, 7.0)])
    path = str(tmp_path / "points")
    pts.save(path)
    loaded = GPListOfPoints.from_file(path)
    assert isinstance(loaded, GPListOfPoints) and list(loaded) == list(pts)
    assert loaded[0].label == PointLabels.evaluated and isinstance(loaded[0].normed_coord, np.ndarray)


def test_gppoint_equality():
    a = point([0.1, 0.2], 1.0)
    assert a == point([0.1, 0.2], 1.0)
    assert a != point([0.1, 0.2], 1.5)
    assert a != point([0.1, 0.25], 1.0)


def test_projection_index_matches_literal_semantics_large():
    """10-D, a few thousand operations: capacity growth of the mirror, replacements whose coordinates move inside the
    tolerance (the index entry must follow), lookups against the brute-force scan."""
    rng = np.random.default_rng(7)
    base = rng.random((600, 10))
    fast, slow = GPListOfPoints(), []
    for step in range(3000):
        coord = base[rng.integers(600)] + (rng.random(10) < 0.2) * rng.normal(0, 2.5e-13, 10)
        label = PointLabels.evaluated if rng.random() < 0.3 else PointLabels.gp_based
        obj = point(coord, mu=float(step), label=label)
        fast.append(obj)
        literal_append(slow, obj)
    assert len(fast) == len(slow)
    for a, b in zip(fast, slow):
        assert a == b and a.label == b.label
    for probe in base[:200]:
        want = next((i for i, p in enumerate(slow) if np.linalg.norm(p.normed_coord - probe) < DUPLICATE_TOLERANCE), None)
        assert fast.index_by_coords(probe) == want


def test_epoch_and_replace_at():
    pts = GPListOfPoints([point([0.1 * i, 0.2, 0.3]) for i in range(5)])
    epoch = pts.epoch
    pts.append(point([0.9, 0.9, 0.9]))            # appending keeps positions
    pts.append(point([0.1, 0.2, 0.3], mu=4.0))    # so does replacing in place
    assert pts.epoch == epoch and pts[1].score_mu == 4.0
    pts.replace_at([0, 5], [point([0.0, 0.2, 0.3], mu=7.0), point([0.9, 0.9, 0.9 + 4e-13], mu=8.0)])
    assert pts.epoch == epoch and pts[0].score_mu == 7.0 and pts[5].score_mu == 8.0
    assert pts.index_by_coords(np.array([0.9, 0.9, 0.9 + 4e-13])) == 5
    assert pts.index_by_coords(np.array([0.9, 0.9, 0.9 - 4e-13])) == 5   # still within the tolerance of the moved point
    pts.insert(0, point([0.55, 0.55, 0.55]))       # positions move: cached indices are invalid from here on
    assert pts.epoch != epoch and pts.index_by_coords(np.array([0.9, 0.9, 0.9 + 4e-13])) == 6


def test_label_positions_follow_every_kind_of_edit():
    """The incrementally maintained positions of evaluated / GP-based points equal a literal scan after appends, in-place
    replacements (GP -> GP, GP -> evaluated), bulk re-predictions and generic list edits, whenever they are asked for."""
    rng = np.random.default_rng(7)
    base = rng.random((60, 2))
    pts = GPListOfPoints()

    def check():
        for label in (PointLabels.evaluated, PointLabels.gp_based):
            assert pts.positions_with_label(label) == [i for i, p in enumerate(pts) if p.label == label]

    check()
    for step in range(600):
        action = rng.random()
        if action < 0.70 or len(pts) < 5:
            label = PointLabels.evaluated if rng.random() < 0.35 else PointLabels.gp_based
            pts.append(point(base[rng.integers(60)], mu=float(step), label=label))
        elif action < 0.85:
            positions = pts.positions_with_label(PointLabels.gp_based)[::2]
            pts.replace_at(positions, (point(pts[i].normed_coord, mu=-1.0 - step, label=PointLabels.gp_based) for i in positions),
                           coords_unchanged=True)
        elif action < 0.92:
            del pts[int(rng.integers(len(pts)))]          # generic edit: index rebuilt lazily
        elif action < 0.96:
            pts.insert(0, point(rng.random(2) + 5.0, label=PointLabels.evaluated))
        else:
            pts.extend([point(rng.random(2) + 9.0, label=PointLabels.gp_based)])
        if step % 7 == 0:
            check()
    check()
    import pickle

    clone = pickle.loads(pickle.dumps(pts))
    assert clone.positions_with_label(PointLabels.evaluated) == pts.positions_with_label(PointLabels.evaluated)
