"""
Parameter space and its ternary partition tree.

Public surface mirrors the reference's ``gpso/param_space.py``: ``LeafNode`` (:20-307) with ``ternary_split``,
``grow``, ``sample_uniformly``, ``get_center_as_list/dict``; ``ParameterSpace`` (:310-467) with ``normalise_coords``,
``denormalise_coords``, ``get_best_score_leaf``, ``max_depth``, ``save`` / ``from_file``.

Differences in construction, not in behaviour:
  * the tree is a plain parent/children structure of this module (the reference mixes in ``anytree.NodeMixin``); nodes
    expose the same ``parent``, ``children``, ``depth``, ``is_leaf`` and pre-order iteration;
  * the root keeps a per-depth index of the nodes in pre-order, so ``get_best_score_leaf`` ("first maximum in
    pre-order") is one scan of a level instead of a full traversal plus sort;
  * ``grow(depth)`` -- the leaf-coordinate batch of the exploitation step -- is generated on the GPU
    (``gpso_grow_leaves``), bit-identical to the reference's per-node Python arithmetic.
"""
import bisect
import pickle
from collections import OrderedDict

import numpy as np
from sklearn.preprocessing import MinMaxScaler

from . import backend as _backend
from .utils import PKL_EXT, PointLabels

NORM_PARAMS_BOUNDS = (0, 1)
_CHILD_TAGS = ("l", "c", "r")


def PreOrderIter(node, filter_=None):
    """Depth-first pre-order traversal (node, then children left to right), like ``anytree.PreOrderIter``."""
    stack = [node]
    while stack:
        current = stack.pop()
        if filter_ is None or filter_(current):
            yield current
        stack.extend(reversed(current.children))


class LeafNode:
    """
    One hyper-rectangle of the partition, in normalised coordinates.  ``norm_bounds`` is a list of (lo, hi) per
    dimension; ``score`` is the evaluated objective or the UCB of the GP at/inside the leaf.
    """

    COMPARE_ATTRS = ["norm_bounds", "parameter_names", "name", "ndim", "depth", "score", "sampled", "label"]
    INIT_ATTRS = ["label", "name", "norm_bounds", "parameter_names", "sampled", "scaler", "score", "children"]

    # ---- validation -------------------------------------------------------------------------------------------------
    @staticmethod
    def _validate_single_bound(single_bound):
        assert isinstance(single_bound, (list, tuple))
        assert len(single_bound) == 2
        assert single_bound[1] > single_bound[0]

    def _validate_param_bounds(self, param_bounds):
        assert param_bounds is not None
        assert isinstance(param_bounds, (list, tuple))
        for single_bound in param_bounds:
            self._validate_single_bound(single_bound)

    # ---- construction -----------------------------------------------------------------------------------------------
    def __init__(
        self,
        norm_bounds,
        scaler,
        parameter_names,
        score=0.0,
        sampled=False,
        label=PointLabels.not_assigned,
        name="",
        parent=None,
        children=None,
    ):
        assert isinstance(scaler, MinMaxScaler), "Scaler must be sklearn's `MinMaxScaler`"
        self.scaler = scaler
        self.name = name
        self.score = score
        self.sampled = sampled
        self.label = label
        self._parent = None
        self._children = []
        self._center_cache = None
        self.parent = parent
        if children:
            self.children = children

        self._validate_param_bounds(norm_bounds)
        assert len(norm_bounds) == self.ndim
        self.norm_bounds = norm_bounds
        assert len(parameter_names) == self.ndim
        assert all(isinstance(param_name, str) for param_name in parameter_names)
        self.parameter_names = parameter_names

    # ---- tree plumbing ----------------------------------------------------------------------------------------------
    @property
    def parent(self):
        return self._parent

    @parent.setter
    def parent(self, new_parent):
        if new_parent is self._parent:
            return
        if self._parent is not None:
            self._parent._children.remove(self)
            self._parent._on_children_changed()
        self._parent = new_parent
        if new_parent is not None:
            new_parent._children.append(self)
            new_parent._on_children_changed(appended=self)

    @property
    def children(self):
        return tuple(self._children)

    @children.setter
    def children(self, new_children):
        for child in list(self._children):
            child._parent = None
        self._children = []
        for child in new_children:
            child.parent = self
        self._on_children_changed()

    def _on_children_changed(self, appended=None):
        root = self.root
        if isinstance(root, ParameterSpace):
            # a childless node appended as the last child (what ``ternary_split`` does) goes straight into the per-depth
            # index; any other change of the tree marks the index for a rebuild
            if appended is not None and not appended._children and not root._index_dirty and root._index_insert(self, appended):
                return
            root._index_dirty = True

    @property
    def root(self):
        node = self
        while node._parent is not None:
            node = node._parent
        return node

    @property
    def depth(self):
        level, node = 0, self
        while node._parent is not None:
            node = node._parent
            level += 1
        return level

    @property
    def is_leaf(self):
        return not self._children

    @property
    def is_root(self):
        return self._parent is None

    @property
    def descendants(self):
        it = PreOrderIter(self)
        next(it)
        return tuple(it)

    def __getitem__(self, pos):
        return self.children[pos]

    def __str__(self):
        return (
            f"Leaf node `{self.name}`: score {self.score}; center at "
            f"{self.get_center_as_dict(normed=True)}; depth {self.depth}"
        )

    __repr__ = __str__

    @property
    def ndim(self):
        return self.scaler.data_max_.shape[0]

    # ---- geometry ---------------------------------------------------------------------------------------------------
    @property
    def norm_bounds(self):
        return self._norm_bounds

    @norm_bounds.setter
    def norm_bounds(self, value):
        self._norm_bounds = value
        self._center_cache = None

    def bounds_array(self):
        return np.array(self.norm_bounds, dtype=np.float64).reshape(self.ndim, 2)

    def center_array(self):
        """Normalised centre as a float64 vector, (lo + hi) / 2 per dimension (cached)."""
        if self._center_cache is None:
            b = self.bounds_array()
            self._center_cache = (b[:, 0] + b[:, 1]) / 2.0
        return self._center_cache

    def get_center_as_list(self, normed=False):
        centers = [float(c) for c in self.center_array()]
        if not normed:
            centers = np.around(_scaler_transform(self.scaler, np.array([centers]), inverse=True), decimals=5)[0].tolist()
        return centers

    def get_center_as_dict(self, normed=False):
        return dict(zip(self.parameter_names, self.get_center_as_list(normed=normed)))

    def sample_uniformly(self, n_points, seed=None):
        """``n_points`` uniform samples inside the leaf, [n_points, ndim] in normalised coordinates."""
        np.random.seed(seed)
        return np.random.uniform(
            low=[bound[0] for bound in self.norm_bounds],
            high=[bound[1] for bound in self.norm_bounds],
            size=(n_points, self.ndim),
        )

    def grow(self, depth):
        """
        Centres of the throw-away subtree of ``depth`` levels below (and including) this leaf, level by level,
        [(3^depth - 1)/2, ndim].  Nothing is attached to the tree.  Computed by the CUDA leaf generator.
        """
        coords = _backend.default_backend().grow_leaves(self.bounds_array(), depth)
        if self._children:
            self.children = list()
        return coords

    def _replace_normed_coord(self, index, new_coord):
        assert index < self.ndim
        self._validate_single_bound(new_coord)
        return [item if idx != index else new_coord for idx, item in enumerate(self.norm_bounds)]

    def ternary_split(self):
        """
        Split into three children along the widest dimension (first one on ties); the children are attached to this
        node and returned as [left, centre, right].  Cut points are ``lo + i * (width / 3)``, i = 0..3.
        """
        widths = [bound[1] - bound[0] for bound in self.norm_bounds]
        split_dim = int(np.argmax(widths))
        step = widths[split_dim] / 3
        low = self.norm_bounds[split_dim][0]
        cuts = [low + i * step for i in range(4)]
        kids = []
        for pos, tag in enumerate(_CHILD_TAGS):
            kids.append(
                LeafNode(
                    norm_bounds=self._replace_normed_coord(split_dim, (cuts[pos], cuts[pos + 1])),
                    scaler=self.scaler,
                    parameter_names=self.parameter_names,
                    name=self.name + "->" + tag,
                    parent=self,
                    children=None,
                )
            )
        if not np.allclose(kids[1].center_array(), self.center_array(), rtol=1e-7, atol=0.0):
            np.testing.assert_allclose(kids[1].center_array(), self.center_array())  # raises with the reference's message
        return kids


def _scaler_transform(scaler, coords, inverse):
    """``MinMaxScaler.transform`` / ``inverse_transform`` without sklearn's per-call input validation (0.4 ms per call, two
    calls per evaluated point): the same two float64 array operations in the same order (``X *= scale_; X += min_`` and
    ``X -= min_; X /= scale_``, sklearn/preprocessing/_data.py), so the result is bit-identical.  Anything but a finite
    float64 array takes sklearn's own path, with its conversions and error messages."""
    if isinstance(coords, np.ndarray) and coords.dtype == np.float64 and coords.ndim == 2 and not scaler.clip and np.isfinite(coords).all():
        out = np.array(coords, dtype=np.float64, order="C")
        if inverse:
            out -= scaler.min_
            out /= scaler.scale_
        else:
            out *= scaler.scale_
            out += scaler.min_
        return out
    return scaler.inverse_transform(coords) if inverse else scaler.transform(coords)


class ParameterSpace(LeafNode):
    """Root of the partition tree: the full (normalised) domain plus the scaler between original and unit coordinates."""

    def __init__(self, parameter_bounds, parameter_names):
        self._validate_param_bounds(parameter_bounds)
        scaler = MinMaxScaler(feature_range=NORM_PARAMS_BOUNDS)
        scaler.fit(np.array(parameter_bounds).T)
        parameter_names = parameter_names or ["" for _ in range(len(parameter_bounds))]
        assert len(parameter_names) == len(parameter_bounds)
        self._index_dirty = True
        self._by_depth = {}
        super().__init__(
            norm_bounds=[NORM_PARAMS_BOUNDS for _ in range(len(parameter_bounds))],
            scaler=scaler,
            parameter_names=parameter_names,
            name="full_domain",
            parent=None,
            children=None,
        )

    # ---- per-depth index --------------------------------------------------------------------------------------------
    # Nodes of one depth in pre-order (the order the reference's stable sort of ``PreOrderIter`` ties on,
    # param_space.py:412-420).  Among nodes of equal depth pre-order is the lexicographic order of the paths (child
    # positions from the root), so a split inserts its children by bisection instead of re-walking the tree
    # (the re-walk per split was the largest item of the host loop in a 500-evaluation run).
    def _depth_index(self):
        if getattr(self, "_index_dirty", True):
            by_depth, keys = {}, {}
            stack = [(self, ())]
            while stack:
                node, path = stack.pop()
                node._path = path
                by_depth.setdefault(len(path), []).append(node)
                keys.setdefault(len(path), []).append(path)
                kids = node._children
                stack.extend((kids[i], path + (i,)) for i in range(len(kids) - 1, -1, -1))
            self._by_depth, self._by_depth_keys = by_depth, keys
            self._index_dirty = False
        return self._by_depth

    def _index_insert(self, parent, child):
        path = getattr(parent, "_path", None)
        if path is None:
            return False
        path = path + (len(parent._children) - 1,)
        depth = len(path)
        keys = self._by_depth_keys.setdefault(depth, [])
        nodes = self._by_depth.setdefault(depth, [])
        at = bisect.bisect_right(keys, path)
        keys.insert(at, path)
        nodes.insert(at, child)
        child._path = path
        return True

    @property
    def max_depth(self):
        return max(self._depth_index())

    def get_best_score_leaf(self, depth, only_not_sampled=True):
        """
        Highest-scored node of a level (optionally only nodes not yet sampled); ties go to the first node in pre-order,
        which is what a stable descending sort of the pre-order traversal yields.
        """
        best = None
        for node in self._depth_index().get(depth, ()):  # lists are in pre-order
            if node.sampled and only_not_sampled:
                continue
            if best is None or node.score > best.score:
                best = node
        return best

    # ---- coordinates ------------------------------------------------------------------------------------------------
    def normalise_coords(self, orig_coords):
        assert orig_coords.ndim == 2
        assert orig_coords.shape[1] == self.ndim
        return _scaler_transform(self.scaler, orig_coords, inverse=False)

    def denormalise_coords(self, normed_coords):
        assert normed_coords.ndim == 2
        assert normed_coords.shape[1] == self.ndim
        return _scaler_transform(self.scaler, normed_coords, inverse=True)

    # ---- persistence (same file layout as the reference: pickled nested OrderedDict, keys sorted) --------------------
    @staticmethod
    def _export(node):
        data = OrderedDict()
        attrs = {
            "label": node.label,
            "name": node.name,
            "norm_bounds": node.norm_bounds,
            "parameter_names": node.parameter_names,
            "sampled": node.sampled,
            "scaler": node.scaler,
            "score": node.score,
        }
        for key in sorted(attrs):
            data[key] = attrs[key]
        if node.children:
            data["children"] = [ParameterSpace._export(child) for child in node.children]
        return data

    @staticmethod
    def _import(data, parent=None):
        kwargs = {key: value for key, value in data.items() if key in LeafNode.INIT_ATTRS and key != "children"}
        node = LeafNode(parent=parent, **kwargs)
        for child in data.get("children", ()):
            ParameterSpace._import(child, parent=node)
        return node

    def save(self, filename):
        if not filename.endswith(PKL_EXT):
            filename += PKL_EXT
        with open(filename, "wb") as handle:
            pickle.dump(self._export(self), handle, protocol=pickle.HIGHEST_PROTOCOL)

    @classmethod
    def from_file(cls, filename):
        if not filename.endswith(PKL_EXT):
            filename += PKL_EXT
        with open(filename, "rb") as handle:
            loaded = pickle.load(handle)
        root = cls._import(loaded)
        root.__class__ = ParameterSpace
        root._index_dirty = True
        root._by_depth = {}
        return root
