"""Persistent homology (Z/2) of a filtered simplicial complex, for the stand-in ``SimplexTree``.

The reference obtains persistence from gudhi (``stree.compute_persistence()`` followed by
``persistence_intervals_in_dimension(i)``, ``flooder/cli.py:473-476`` and
``tests/test_flooder.py:55-58``).  gudhi is not available in the build image, so the tree returned by
``flood_complex(..., return_simplex_tree=True)`` carries this small implementation with the same
method names: standard boundary-matrix reduction with the "twist" clearing optimisation, columns
processed from the top dimension down.  Complexes on this path have a few 10^4 ... 10^5 simplices
(Delaunay complex of the landmarks), for which a sparse pure-Python reduction takes well under a
second per 10^4 simplices.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np


def filtration_order(simplices: Sequence[Tuple[int, ...]], values: Sequence[float]) -> List[int]:
    """Indices sorted by (value, dimension, vertices): a filtration order in which every face
    precedes its cofaces, provided the values are non-decreasing along inclusions."""
    return sorted(range(len(simplices)), key=lambda i: (values[i], len(simplices[i]), simplices[i]))


def persistence_pairs(simplices: Sequence[Tuple[int, ...]], values: Sequence[float]):
    """Reduce the boundary matrix over Z/2.

    Returns ``(pairs, essential)`` with ``pairs`` = list of (birth index, death index) into
    ``simplices`` and ``essential`` = list of unpaired positive simplex indices."""
    order = filtration_order(simplices, values)
    position: Dict[Tuple[int, ...], int] = {simplices[i]: p for p, i in enumerate(order)}
    n = len(order)
    dims = [len(simplices[i]) - 1 for i in order]
    top = max(dims, default=0)
    by_dim: List[List[int]] = [[] for _ in range(top + 1)]
    for p, d in enumerate(dims):
        by_dim[d].append(p)

    pivot_owner: Dict[int, int] = {}      # low row -> column that owns it
    reduced: Dict[int, set] = {}          # non-zero reduced columns
    cleared = [False] * n                 # positive simplices already known to be paired
    paired_birth = [False] * n
    pairs: List[Tuple[int, int]] = []
    for d in range(top, 0, -1):
        for col in by_dim[d]:
            if cleared[col]:
                continue
            s = simplices[order[col]]
            column = {position[s[:i] + s[i + 1:]] for i in range(len(s))}
            while column:
                low = max(column)
                other = pivot_owner.get(low)
                if other is None:
                    break
                column ^= reduced[other]
            if column:
                low = max(column)
                pivot_owner[low] = col
                reduced[col] = column
                cleared[low] = True        # twist: the column of `low` reduces to zero
                paired_birth[low] = True
                pairs.append((order[low], order[col]))
    negative = {death for _, death in pairs}
    essential = [order[p] for p in range(n) if not paired_birth[p] and order[p] not in negative]
    return pairs, essential


class PersistenceMixin:
    """gudhi-style persistence API for a class exposing ``get_simplices()`` and ``dimension()``."""

    _intervals: Dict[int, np.ndarray]

    def compute_persistence(self, homology_coeff_field: int = 2, min_persistence: float = 0.0,
                            persistence_dim_max: bool = False) -> None:
        """Z/2 persistence of the current filtration.  Like gudhi, homology in the top dimension of
        the complex is skipped unless ``persistence_dim_max`` is set, and intervals of length
        <= ``min_persistence`` are dropped (``min_persistence < 0`` keeps everything)."""
        if homology_coeff_field != 2:
            raise NotImplementedError("only Z/2 coefficients are implemented")
        simplices, values = [], []
        for s, f in self.get_simplices():
            if math.isnan(f):
                raise ValueError(f"simplex {s} has no filtration value")
            simplices.append(tuple(s))
            values.append(float(f))
        pairs, essential = persistence_pairs(simplices, values)
        top = self.dimension()
        limit = top if persistence_dim_max else top - 1
        out: Dict[int, List[Tuple[float, float]]] = {}
        for b, d in pairs:
            dim = len(simplices[b]) - 1
            if dim <= limit and values[d] - values[b] > min_persistence:
                out.setdefault(dim, []).append((values[b], values[d]))
        for b in essential:
            dim = len(simplices[b]) - 1
            if dim <= limit:
                out.setdefault(dim, []).append((values[b], math.inf))
        self._intervals = {k: np.asarray(sorted(v), dtype=np.float64).reshape(-1, 2) for k, v in out.items()}

    def persistence_intervals_in_dimension(self, dimension: int) -> np.ndarray:
        if not hasattr(self, "_intervals"):
            raise RuntimeError("compute_persistence() must be called first")
        return self._intervals.get(dimension, np.empty((0, 2), dtype=np.float64))

    def persistence(self, homology_coeff_field: int = 2, min_persistence: float = 0.0,
                    persistence_dim_max: bool = False):
        self.compute_persistence(homology_coeff_field, min_persistence, persistence_dim_max)
        out = [(dim, (float(b), float(d))) for dim, arr in self._intervals.items() for b, d in arr]
        return sorted(out, key=lambda t: (-t[0], -(t[1][1] - t[1][0])))

    def betti_numbers(self) -> List[int]:
        if not hasattr(self, "_intervals"):
            raise RuntimeError("compute_persistence() must be called first")
        top = max(self._intervals, default=-1)
        return [int(np.isinf(self.persistence_intervals_in_dimension(d)[:, 1]).sum()) for d in range(top + 1)]
