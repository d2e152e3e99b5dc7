"""Alpha-complex filtration and bottleneck distance.  TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

The reference checks its Flood complex against ``gudhi.AlphaComplex`` persistence with
``gudhi.bottleneck_distance`` (``tests/test_flooder.py:24-75``: with landmarks == cloud the two
filtrations must give the same diagrams up to 5e-4).  gudhi is absent here, so the two published
algorithms are restated for small inputs:

* ``alpha_filtration``  -- Delaunay simplices with alpha values (not squared), the propagation
  scheme of gudhi's ``Alpha_complex::create_complex``: top-down, a face inherits the value of a
  coface whenever the coface's opposite vertex lies inside the face's smallest circumsphere
  (the face is not Gabriel), otherwise it gets its own circumradius.
* ``bottleneck_distance`` -- exact, by binary search over the candidate distances with a
  maximum bipartite matching (scipy Hopcroft-Karp).
"""
from __future__ import annotations

import itertools
from typing import Dict, Tuple

import numpy as np

from .simplex_tree import delaunay_top_simplices


def _circumsphere(p: np.ndarray) -> Tuple[np.ndarray, float]:
    """Centre (in the affine hull) and squared radius of the smallest sphere through the rows of p."""
    if p.shape[0] == 1:
        return p[0], 0.0
    a = p[1:] - p[0]
    gram = a @ a.T
    rhs = 0.5 * np.einsum("ij,ij->i", a, a)
    lam = np.linalg.solve(gram, rhs)
    off = lam @ a
    return p[0] + off, float(off @ off)


def alpha_filtration(points: np.ndarray) -> Dict[Tuple[int, ...], float]:
    pts = np.asarray(points, dtype=np.float64)
    cells = delaunay_top_simplices(pts)
    K = cells.shape[1]
    value: Dict[Tuple[int, ...], float] = {}
    by_size = {k: set() for k in range(1, K + 1)}
    for row in cells.tolist():
        for k in range(1, K + 1):
            by_size[k].update(itertools.combinations(row, k))
    for k in range(K, 1, -1):
        for s in by_size[k]:
            if s not in value:
                value[s] = _circumsphere(pts[list(s)])[1]
            vs = value[s]
            for i in range(k):
                face = s[:i] + s[i + 1:]
                if face in value:
                    if vs < value[face]:
                        value[face] = vs
                elif k > 2:
                    c, r2 = _circumsphere(pts[list(face)])
                    opp = pts[s[i]] - c
                    if opp @ opp < r2:          # not Gabriel: the face appears with its coface
                        value[face] = vs
    for v in by_size[1]:
        value[v] = 0.0
    return {s: float(np.sqrt(max(f, 0.0))) for s, f in value.items()}


def bottleneck_distance(a: np.ndarray, b: np.ndarray) -> float:
    """Bottleneck distance between two persistence diagrams ((n,2) arrays, L-infinity ground
    metric, points may be matched to the diagonal).  Essential classes (death = inf) are matched
    among themselves by sorted birth."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import maximum_bipartite_matching

    a = np.asarray(a, dtype=np.float64).reshape(-1, 2)
    b = np.asarray(b, dtype=np.float64).reshape(-1, 2)
    ia, ib = np.isinf(a[:, 1]), np.isinf(b[:, 1])
    if ia.sum() != ib.sum():
        return float("inf")
    ess = float(np.abs(np.sort(a[ia, 0]) - np.sort(b[ib, 0])).max()) if ia.any() else 0.0
    a, b = a[~ia], b[~ib]
    na, nb = len(a), len(b)
    if na + nb == 0:
        return ess
    da = 0.5 * (a[:, 1] - a[:, 0])          # distance to the diagonal
    db = 0.5 * (b[:, 1] - b[:, 0])
    cross = np.abs(a[:, None, :] - b[None, :, :]).max(axis=2) if na and nb else np.zeros((na, nb))
    # left nodes: a-points + nb diagonal copies; right nodes: b-points + na diagonal copies
    cost = np.zeros((na + nb, nb + na))
    cost[:na, :nb] = cross
    cost[:na, nb:] = np.inf
    cost[np.arange(na), nb + np.arange(na)] = da
    cost[na:, :nb] = np.inf
    cost[na + np.arange(nb), np.arange(nb)] = db
    cost[na:, nb:] = 0.0
    cand = np.unique(np.concatenate([cross.ravel(), da, db, [0.0]]))

    def feasible(eps: float) -> bool:
        match = maximum_bipartite_matching(csr_matrix(cost <= eps), perm_type="column")
        return bool((match >= 0).all())

    lo, hi = 0, len(cand) - 1
    while lo < hi:
        mid = (lo + hi) // 2
        if feasible(cand[mid]):
            hi = mid
        else:
            lo = mid + 1
    return max(ess, float(cand[lo]))
