"""Host-side complex container and Delaunay step.

The reference keeps both inside gudhi (``gudhi.DelaunayComplex(landmarks).create_simplex_tree()``,
``flooder/core.py:130-132``; ``assign_filtration`` / ``make_filtration_non_decreasing`` /
``get_simplices``, ``:278-288``).  When gudhi is importable it is used exactly like that and
``flood_complex(..., return_simplex_tree=True)`` returns a real ``gudhi.SimplexTree``.  The build
image has no gudhi, so this module also provides a small array-oriented ``SimplexTree`` with the
same method names, and a Qhull-based triangulation (``scipy.spatial.Delaunay``), which produces
the same simplex sets as gudhi/CGAL on every fixture the reference ships (tests/golden).
"""
from __future__ import annotations

import contextlib
import gc
import itertools
import math
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np

from .persistence import PersistenceMixin

try:
    import gudhi as _gudhi

    HAS_GUDHI = hasattr(_gudhi, "DelaunayComplex")
except Exception:  # noqa: BLE001  (gudhi is not in the build image; tests inject a stand-in)
    _gudhi = None
    HAS_GUDHI = False


def delaunay_cells(landmarks: np.ndarray) -> np.ndarray:
    """Top-dimensional Delaunay cells as ascending vertex index rows, shape (S, D+1).

    Qhull stand-in for ``gudhi.DelaunayComplex`` (``flooder/core.py:130-132``).  The landmarks
    are triangulated in the precision they are given in (the reference hands its landmark tensor
    to gudhi as is).  1-D landmarks need no Qhull: the cells are the consecutive pairs of the
    sorted coordinates.  Degenerate inputs are reported instead of silently changing the complex:
    affinely dependent landmarks raise, landmarks Qhull leaves out (duplicates) warn."""
    import warnings

    from scipy.spatial import Delaunay
    from scipy.spatial import QhullError

    pts = np.asarray(landmarks, dtype=np.float64)
    n, d = pts.shape
    if n <= d:  # not enough points for a full-dimensional cell
        return np.arange(n, dtype=np.int64)[None, :]
    if d == 1:
        order = np.argsort(pts[:, 0], kind="stable")
        cells = np.stack([order[:-1], order[1:]], axis=1).astype(np.int64)
        cells.sort(axis=1)
        return cells
    try:
        tri = Delaunay(pts)
    except QhullError as exc:
        raise RuntimeError(
            f"the {n} landmarks do not span {d} dimensions (Qhull: {str(exc).splitlines()[0]}); the Qhull "
            "stand-in cannot triangulate degenerate landmark sets -- install gudhi, which flooder_b200 "
            "uses for the Delaunay step when it is importable") from exc
    if tri.coplanar.size:
        warnings.warn(
            f"{tri.coplanar.shape[0]} of {n} landmarks are not vertices of the Delaunay triangulation "
            "(duplicate or degenerate points); they do not appear in the complex", RuntimeWarning, stacklevel=3)
    cells = tri.simplices.astype(np.int64)
    cells.sort(axis=1)
    return cells


def faces_of_cells(cells: np.ndarray, k: int) -> np.ndarray:
    """Unique (k-1)-dimensional faces (k vertices) of the given cells, rows ascending."""
    cells = np.asarray(cells, dtype=np.int64)
    width = cells.shape[1]
    if k == width:
        return np.unique(cells, axis=0)
    cols = np.array(list(itertools.combinations(range(width), k)), dtype=np.int64)
    faces = cells[:, cols].reshape(-1, k)
    return np.unique(faces, axis=0)


@contextlib.contextmanager
def _gc_paused():
    """Building tens of thousands of key tuples triggers a cascade of generational collections
    (measured: 9.0 ms with, 3.0 ms without, for the 25 k simplices of 1000 landmarks in 3-D); tuples
    of ints cannot be part of a reference cycle, so the collector is paused while they are made."""
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        yield
    finally:
        if was_enabled:
            gc.enable()


class FaceTable:
    """All faces of a pure simplicial complex given by its top cells, as arrays.

    For every face size k (1..K) it holds the unique faces as a lexicographically sorted
    ``(F_k, k)`` array and, for every top cell, the ids of its faces in the order of
    ``itertools.combinations(range(K), k)``.  Rows are packed into int64 keys (base = number of
    vertices) when they fit, which turns ``np.unique`` / facet lookups into 1-D sorts; this is the
    vectorised replacement of the reference's python loops over simplices
    (``flooder/core.py:135-138, 258-263, 278-288``).
    """

    def __init__(self, cells: np.ndarray, n_vertices: Optional[int] = None) -> None:
        cells = np.ascontiguousarray(cells, dtype=np.int64)
        self.cells = cells
        self.K = K = cells.shape[1]
        self.base = int(n_vertices if n_vertices is not None else (cells.max() + 1 if cells.size else 1))
        self.combos = {k: list(itertools.combinations(range(K), k)) for k in range(1, K + 1)}
        self.faces: Dict[int, np.ndarray] = {}
        self.keys: Dict[int, Optional[np.ndarray]] = {}
        self.cell_face: Dict[int, np.ndarray] = {}
        for k in range(1, K + 1):
            cols = np.asarray(self.combos[k], dtype=np.int64)
            rows = cells[:, cols].reshape(-1, k)
            uniq_rows, keys, inv = self._unique(rows)
            self.faces[k] = uniq_rows
            self.keys[k] = keys
            self.cell_face[k] = inv.reshape(cells.shape[0], len(self.combos[k]))

    def _packable(self, k: int) -> bool:
        return k * np.log2(max(self.base, 2)) < 62.0

    def _pack(self, rows: np.ndarray) -> np.ndarray:
        key = rows[:, 0].copy()
        for j in range(1, rows.shape[1]):
            key = key * self.base + rows[:, j]
        return key

    def _unpack(self, keys: np.ndarray, k: int) -> np.ndarray:
        rows = np.empty((len(keys), k), dtype=np.int64)
        rest = keys
        for j in range(k - 1, -1, -1):
            rest, rows[:, j] = np.divmod(rest, self.base)
        return rows

    def _unique(self, rows: np.ndarray):
        k = rows.shape[1]
        if self._packable(k):
            # (return_index would switch np.unique to a stable merge sort: twice the time; the
            # unique rows are unpacked from the keys instead)
            keys, inv = np.unique(self._pack(rows), return_inverse=True)
            return self._unpack(keys, k), keys, inv.reshape(-1)
        uniq, inv = np.unique(rows, axis=0, return_inverse=True)
        return uniq, None, inv.reshape(-1)

    def lookup(self, rows: np.ndarray) -> np.ndarray:
        """Ids of the given faces (all must exist) in ``faces[k]``."""
        k = rows.shape[1]
        if self.keys[k] is not None:
            return np.searchsorted(self.keys[k], self._pack(rows))
        table = {tuple(r): i for i, r in enumerate(self.faces[k].tolist())}
        return np.fromiter((table[tuple(r)] for r in rows.tolist()), dtype=np.int64, count=len(rows))

    def nan_values(self) -> Dict[int, np.ndarray]:
        return {k: np.full(len(self.faces[k]), np.nan) for k in self.faces}

    def make_non_decreasing(self, values: Dict[int, np.ndarray]) -> None:
        """In place: every face is raised to the largest value among its facets, by increasing
        size; NaN (unassigned) counts as minus infinity (np.fmax ignores NaN)."""
        for k in range(2, self.K + 1):
            f = self.faces[k]
            best = np.full(len(f), np.nan)
            for drop in range(k):
                facet = np.delete(f, drop, axis=1)
                best = np.fmax(best, values[k - 1][self.lookup(facet)])
            values[k] = np.fmax(values[k], best)


class SimplexTree(PersistenceMixin):
    """Filtered simplicial complex keyed by ascending vertex tuples.

    Mirrors the subset of ``gudhi.SimplexTree`` that the reference and its tests touch.
    Unassigned simplices carry NaN (what gudhi's DelaunayComplex hands out with
    ``filtration=None``).
    """

    def __init__(self) -> None:
        self._f: Dict[Tuple[int, ...], float] = {}

    # ---- construction -------------------------------------------------------------------
    @classmethod
    def from_cells(cls, cells: np.ndarray) -> "SimplexTree":
        st = cls()
        nan = float("nan")
        width = np.asarray(cells).shape[1]
        for k in range(1, width + 1):
            st._f.update(dict.fromkeys(map(tuple, faces_of_cells(cells, k).tolist()), nan))
        return st

    @classmethod
    def from_arrays(cls, faces: Dict[int, np.ndarray], values: Dict[int, np.ndarray],
                    keys: Optional[Dict[int, list]] = None) -> "SimplexTree":
        """``keys`` (optional): the key tuples per dimension as ``face_keys`` builds them -- they do
        not depend on the values, so a caller waiting for the device can prepare them meanwhile."""
        st = cls()
        if keys is None:
            keys = face_keys(faces)
        with _gc_paused():
            for k in sorted(faces):
                st._f.update(zip(keys[k], values[k].tolist()))
        return st

    def insert(self, simplex: Iterable[int], filtration: float = 0.0) -> bool:
        key = tuple(sorted(int(v) for v in simplex))
        filtration = float(filtration)
        changed = False
        for k in range(1, len(key) + 1):
            for face in itertools.combinations(key, k):
                old = self._f.get(face)
                if old is None or filtration < old:
                    self._f[face] = filtration
                    changed = True
        return changed

    # ---- queries ------------------------------------------------------------------------
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

    def get_skeleton(self, dimension: int) -> Iterator[Tuple[List[int], float]]:
        for key, val in self.get_simplices():
            if len(key) <= dimension + 1:
                yield key, val

    def get_filtration(self) -> Iterator[Tuple[List[int], float]]:
        def order(k):
            v = self._f[k]
            return (math.inf if math.isnan(v) else v, len(k), k)

        for key in sorted(self._f, key=order):
            yield list(key), self._f[key]

    def get_boundaries(self, simplex: Iterable[int]) -> Iterator[Tuple[List[int], float]]:
        key = tuple(sorted(simplex))
        if len(key) > 1:
            for i in range(len(key)):
                face = key[:i] + key[i + 1:]
                yield list(face), self._f[face]

    # ---- mutation -----------------------------------------------------------------------
    def assign_filtration(self, simplex: Iterable[int], filtration: float) -> None:
        key = tuple(sorted(simplex))
        if key not in self._f:
            raise KeyError(f"simplex {key} is not in the complex")
        self._f[key] = float(filtration)

    def assign_many(self, simplices: Sequence[Tuple[int, ...]], values: Sequence[float]) -> None:
        """Bulk ``assign_filtration`` for keys that are already ascending tuples."""
        self._f.update(zip(simplices, values))

    def make_filtration_non_decreasing(self) -> bool:
        """Every simplex is raised to the largest value among its facets, by increasing
        dimension; NaN (unassigned) counts as minus infinity."""
        f = self._f
        changed = False
        for key in sorted(f, key=len):
            n = len(key)
            if n == 1:
                continue
            best = -math.inf
            for i in range(n):
                v = f[key[:i] + key[i + 1:]]
                if v > best:  # False for NaN
                    best = v
            cur = f[key]
            if best > -math.inf and not cur >= best:  # cur < best or cur is NaN
                f[key] = best
                changed = True
        return changed

    def to_dict(self) -> Dict[Tuple[int, ...], float]:
        return {tuple(s): v for s, v in self.get_simplices()}

    def to_flat_dict(self) -> Dict[Tuple[int, ...], float]:
        """``{simplex: value}`` as ``flood_complex`` returns it (insertion order: by dimension, then
        lexicographic -- the order of ``get_simplices`` for a tree built by ``from_arrays``)."""
        return dict(self._f)


def face_keys(faces: Dict[int, np.ndarray]) -> Dict[int, list]:
    """Key tuples of the simplices per dimension (zip over the columns builds the tuples directly,
    no intermediate row lists)."""
    out = {}
    with _gc_paused():
        for k in sorted(faces):
            cols = [faces[k][:, j].tolist() for j in range(faces[k].shape[1])]
            out[k] = list(zip(*cols))
    return out


def delaunay_complex(landmarks: np.ndarray):
    """Host Delaunay step.  Returns ``(cells, gudhi_tree_or_None)``: with gudhi installed the
    triangulation and the container are gudhi's (as in the reference), otherwise Qhull's cells."""
    if HAS_GUDHI:
        tree = _gudhi.DelaunayComplex(np.asarray(landmarks)).create_simplex_tree()
        width = landmarks.shape[1] + 1
        cells = np.asarray([s for s, _ in tree.get_simplices() if len(s) == width], dtype=np.int64)
        if cells.size == 0:  # degenerate input: fall back to the maximal simplices present
            top = max(len(s) for s, _ in tree.get_simplices())
            cells = np.asarray([s for s, _ in tree.get_simplices() if len(s) == top], dtype=np.int64)
        return cells, tree
    return delaunay_cells(landmarks), None
