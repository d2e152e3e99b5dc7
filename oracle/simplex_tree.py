"""Dict-backed stand-in for ``gudhi.SimplexTree`` / ``gudhi.DelaunayComplex``.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  The reference keeps its Delaunay step
and its output container inside gudhi 3.11.0 (``/root/reference/flooder/core.py:130-138``
and ``:278-288``), which is not installable here.  This module restates the published
behaviour of exactly the calls the reference makes:

* ``DelaunayComplex(points).create_simplex_tree()``  -> the Delaunay triangulation with its
  whole face closure, every filtration value NaN (gudhi: "filtration=None").
  Triangulation comes from Qhull (``scipy.spatial.Delaunay``); equality of the simplex
  sets with gudhi's is pinned on the fixtures the reference ships
  (``docs/visualization/*/tetrahedra.csv``, see ``tests/test_oracle_golden.py``).
* ``get_simplices`` / ``assign_filtration`` / ``make_filtration_non_decreasing`` /
  ``insert`` / ``get_boundaries`` / ``filtration`` / ``num_simplices`` / ``num_vertices``.
"""
from __future__ import annotations

import itertools
import math
from typing import Dict, Iterable, Iterator, List, Tuple

import numpy as np


def delaunay_top_simplices(points: np.ndarray) -> np.ndarray:
    """Sorted vertex lists of the top-dimensional Delaunay cells (Qhull)."""
    from scipy.spatial import Delaunay

    pts = np.asarray(points, dtype=np.float64)
    if pts.shape[0] <= pts.shape[1]:
        # fewer points than needed for a full-dimensional cell: a single simplex
        return np.arange(pts.shape[0], dtype=np.int64)[None, :]
    if pts.shape[1] == 1:
        # 1-D: the Delaunay triangulation is the chain of consecutive points (Qhull needs >= 2-D)
        order = np.argsort(pts[:, 0], kind="stable")
        return np.sort(np.stack([order[:-1], order[1:]], axis=1).astype(np.int64), axis=1)
    tri = Delaunay(pts)
    return np.sort(tri.simplices.astype(np.int64), axis=1)


class DictSimplexTree:
    """Minimal simplicial complex container keyed by ascending vertex tuples."""

    def __init__(self) -> None:
        self._f: Dict[Tuple[int, ...], float] = {}

    # -- construction -----------------------------------------------------------------
    def insert(self, simplex: Iterable[int], filtration: float = 0.0) -> bool:
        """gudhi semantics: insert the simplex and all missing faces with ``filtration``;
        existing simplices keep min(old, new)."""
        key = tuple(sorted(int(v) for v in simplex))
        changed = False
        for k in range(1, len(key) + 1):
            for face in itertools.combinations(key, k):
                old = self._f.get(face)
                if old is None:
                    self._f[face] = float(filtration)
                    changed = True
                elif filtration < old:
                    self._f[face] = float(filtration)
                    changed = True
        return changed

    @classmethod
    def from_top_simplices(cls, tops: np.ndarray) -> "DictSimplexTree":
        st = cls()
        f = st._f
        nan = float("nan")
        for row in np.asarray(tops).tolist():
            for k in range(1, len(row) + 1):
                for face in itertools.combinations(row, k):
                    f[face] = nan
        return st

    # -- queries ----------------------------------------------------------------------
    def num_simplices(self) -> int:
        return len(self._f)

    def num_vertices(self) -> int:
        return sum(1 for k in self._f if len(k) == 1)

    def dimension(self) -> int:
        return max((len(k) for k in self._f), default=0) - 1

    def find(self, simplex: Iterable[int]) -> bool:
        return tuple(sorted(simplex)) in self._f

    def filtration(self, simplex: Iterable[int]) -> float:
        return self._f[tuple(sorted(simplex))]

    def get_simplices(self) -> Iterator[Tuple[List[int], float]]:
        for key in sorted(self._f, key=lambda k: (len(k), k)):
            yield list(key), self._f[key]

    def get_filtration(self) -> Iterator[Tuple[List[int], float]]:
        def order(k):
            v = self._f[k]
            return (math.inf if math.isnan(v) else v, len(k), k)

        for key in sorted(self._f, key=order):
            yield list(key), self._f[key]

    def get_boundaries(self, simplex: Iterable[int]) -> Iterator[Tuple[List[int], float]]:
        key = tuple(sorted(simplex))
        if len(key) <= 1:
            return
        for i in range(len(key)):
            face = key[:i] + key[i + 1:]
            yield list(face), self._f[face]

    # -- mutation ---------------------------------------------------------------------
    def assign_filtration(self, simplex: Iterable[int], filtration: float) -> None:
        key = tuple(sorted(simplex))
        if key not in self._f:
            raise KeyError(f"simplex {key} not in complex")
        self._f[key] = float(filtration)

    def make_filtration_non_decreasing(self) -> bool:
        """Raise every simplex to the max of its facets, by increasing dimension
        (gudhi ``Simplex_tree::make_filtration_non_decreasing``).  NaN counts as -inf:
        a NaN simplex takes the max of its facets, a NaN facet never raises a coface."""
        changed = False
        f = self._f
        for key in sorted(f, key=len):
            if len(key) == 1:
                continue
            best = -math.inf
            for i in range(len(key)):
                v = f[key[:i] + key[i + 1:]]
                if not math.isnan(v) and v > best:
                    best = v
            cur = f[key]
            if best > -math.inf and (math.isnan(cur) or cur < best):
                f[key] = best
                changed = True
        return changed


class DelaunayComplex:
    """Stand-in for ``gudhi.DelaunayComplex`` (reference call site ``core.py:130-132``)."""

    def __init__(self, points) -> None:
        if hasattr(points, "detach"):  # torch tensor, possibly on a CUDA device (core.py:130)
            points = points.detach().cpu().numpy()
        self._points = np.asarray(points, dtype=np.float64)

    def create_simplex_tree(self) -> DictSimplexTree:
        return DictSimplexTree.from_top_simplices(delaunay_top_simplices(self._points))
