"""numpy/scipy restatement of the reference's Flood-complex CPU path.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``): this is the checker, never the product.

Every function cites the reference lines it follows (paths relative to
``/root/reference``).  The distance step is the reference's own CPU back-end: an exact
nearest-neighbour query per sample point on a ``scipy.spatial.KDTree``
(``flooder/core.py:128`` and ``:197-199``), float64 distances of float32 inputs.

Pinned by ``tests/test_oracle_golden.py`` against
  * ``docs/animation/{points,landmarks,edges,triangles}.csv`` (shipped by the reference
    authors, produced by ``docs/animation/generate_csvs.py`` with the CPU path), and
  * ``tests/golden/*.npz`` (the reference's ``core.py`` executed unmodified in the
    authoring container through ``oracle/ref_shims.py``; generator:
    ``tests/golden/make_golden.py``).
"""
from __future__ import annotations

import itertools
from math import comb
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .simplex_tree import DelaunayComplex, DictSimplexTree


# ----------------------------------------------------------------------------------------
# sample-point generators
# ----------------------------------------------------------------------------------------
def generate_grid(n: int, dim: int, dtype=np.float32):
    """Barycentric lattice with ``n`` points per edge on a ``dim``-simplex.

    Follows ``flooder/core.py:346-402``: rows are the compositions of ``n-1`` into
    ``dim+1`` parts in the order induced by ``itertools.combinations`` ("stars and
    bars", ``:369-380``), weights are ``k/(n-1)`` (``:400-401``); for every ``k`` and
    every ``k``-subset ``Z`` of coordinates ``face_idxs[k][j]`` lists the rows whose
    coordinates in ``Z`` are all zero and ``vertex_idxs[k][j]`` the complementary
    coordinates (``:386-399``).
    """
    bars = np.array(list(itertools.combinations(range(n + dim - 1), dim)), dtype=np.int64)
    bars = bars.reshape(-1, dim)
    rows = bars.shape[0]
    fenced = np.concatenate(
        [np.full((rows, 1), -1, np.int64), bars, np.full((rows, 1), n + dim - 1, np.int64)], axis=1
    )
    counts = np.diff(fenced, axis=1) - 1  # (C, dim+1) integer lattice coordinates
    face_idxs: List[np.ndarray] = []
    vertex_idxs: List[np.ndarray] = []
    axes = np.arange(dim + 1)
    for k in range(dim + 1):
        rows_k, verts_k = [], []
        for zero_set in itertools.combinations(range(dim + 1), k):
            if k == 0:
                sel = np.ones(rows, dtype=bool)
            else:
                sel = (counts[:, list(zero_set)] == 0).all(axis=1)
            rows_k.append(np.nonzero(sel)[0])
            verts_k.append(axes[~np.isin(axes, zero_set)])
        face_idxs.append(np.stack(rows_k))
        vertex_idxs.append(np.stack(verts_k))
    # torch.divide(int64 tensor, python int, out=<dtype>) (core.py:400-401) divides in the promoted
    # type of its inputs -- float32, torch's default -- and only then converts to the dtype of
    # `out`: float64 weights of the reference carry float32-rounded quotients.
    weights = counts.astype(np.float32) / np.float32(n - 1)
    return weights.astype(dtype), vertex_idxs, face_idxs


def generate_uniform_weights(num_rand: int, dim: int, dtype=np.float32) -> np.ndarray:
    """Dirichlet(1,..,1) weights, ``flooder/core.py:405-427``.

    The reference draws ``torch.rand(num_rand, dim+1)`` from torch's *global CPU*
    generator (``:425``), so the same call is made here to consume the same stream.
    """
    import torch

    if dim == 0:
        return np.ones((num_rand, 1), dtype=dtype)
    w = -torch.log(1 - torch.rand(num_rand, dim + 1))
    w = w.to(dtype=torch.float64 if dtype == np.float64 else torch.float32)
    w = w / w.sum(dim=1, keepdim=True)
    return w.numpy().astype(dtype)


# ----------------------------------------------------------------------------------------
# bounding balls  (candidate rule; result-neutral on CPU, defines the work count E)
# ----------------------------------------------------------------------------------------
def bounding_balls(simplex_vertices: np.ndarray, d: int) -> Tuple[np.ndarray, np.ndarray]:
    """``flooder/core.py:156-172``: centre = midpoint of the longest edge (first maximum
    of the flattened ``(d+1)x(d+1)`` distance matrix), radius = farthest vertex from the
    centre times 1.42 (1.01 for d<=1) plus 1e-3.  Computed in the input dtype."""
    v = np.asarray(simplex_vertices)
    dt = v.dtype.type
    S, K, D = v.shape

    def norms(diff):
        # squares rounded in the input dtype and added in coordinate order (no pairwise
        # re-association, no fused multiply-add): the operation sequence is part of the
        # contract, near-ties between edges must resolve the same way everywhere
        acc = np.zeros(diff.shape[:-1], dtype=v.dtype)
        for a in range(diff.shape[-1]):
            acc = acc + diff[..., a] * diff[..., a]
        return np.sqrt(acc)

    pair = norms(v[:, :, None, :] - v[:, None, :, :])          # (S,K,K)
    flat = pair.reshape(S, K * K).argmax(axis=1)               # first maximum, row-major
    i0, i1 = np.divmod(flat, K)
    ar = np.arange(S)
    centers = (v[ar, i0] + v[ar, i1]) / dt(2.0)
    far = norms(v - centers[:, None, :]).max(axis=1)
    radii = far * dt(1.42 if d > 1 else 1.01) + dt(1e-3)
    return centers.astype(v.dtype), radii.astype(v.dtype)


def ball_candidate_counts(points: np.ndarray, centers: np.ndarray, radii: np.ndarray) -> np.ndarray:
    """Number of cloud points with ``sum_i (p_i - c_i)^2 <= r^2`` per ball -- the predicate
    of the reference's mask kernel (``flooder/triton_kernels.py:137-148``), float32, the
    squares accumulated in coordinate order with fused multiply-add (Triton's default
    contraction).  Uses the plain-C routine for exact float32/fmaf arithmetic."""
    from . import native

    return native.ball_counts(points, centers, radii)


# ----------------------------------------------------------------------------------------
# farthest-point sampling
# ----------------------------------------------------------------------------------------
def fps_exact(points: np.ndarray, n_samples: int, start_idx: int = 0) -> np.ndarray:
    """Exact farthest-point sampling; restates what the reference obtains from
    ``fpsample.bucket_fps_kdline_sampling`` (fpsample 0.3.3, ``flooder/core.py:337-342``;
    bucket-FPS is an exact acceleration of this recurrence):

        idx[0] = start_idx;  idx[k+1] = argmax_i min_{j<=k} |p_i - p_idx[j]|^2

    float32, squared distance summed in coordinate order without contraction, first
    maximum wins.  Pinned on ``docs/animation/landmarks.csv`` (25 of 200 points)."""
    p = np.ascontiguousarray(points, dtype=np.float32)
    n = p.shape[0]
    n_samples = min(int(n_samples), n)
    out = np.empty(n_samples, dtype=np.int64)
    mind = np.full(n, np.inf, dtype=np.float32)
    cur = int(start_idx)
    for k in range(n_samples):
        out[k] = cur
        diff = p - p[cur]
        sq = diff[:, 0] * diff[:, 0]
        for j in range(1, p.shape[1]):
            sq = sq + diff[:, j] * diff[:, j]
        np.minimum(mind, sq, out=mind)
        cur = int(np.argmax(mind))
    return out


def generate_landmarks(points: np.ndarray, n_lms: int, start_idx: Optional[int] = 0) -> np.ndarray:
    """``flooder/core.py:291-343``: validate, clamp, FPS, gather coordinates in FPS order."""
    if n_lms <= 0:
        raise RuntimeError(f"Number of landmarks ({n_lms}) must be positive")
    n_lms = min(n_lms, len(points))
    if start_idx is None:
        start_idx = int(np.random.randint(len(points)))
    idx = fps_exact(points, n_lms, start_idx)
    return np.asarray(points)[idx]


# ----------------------------------------------------------------------------------------
# the filtration
# ----------------------------------------------------------------------------------------
def sample_points(weights: np.ndarray, simplex_vertices: np.ndarray) -> np.ndarray:
    """``flooder/core.py:188``: ``weights[None] @ simplex_vertices`` -> (S, R, D) in the
    input dtype, accumulated over the K = d+1 vertices in order."""
    w = np.asarray(weights)
    v = np.asarray(simplex_vertices)
    if v.dtype == np.float32 and w.dtype == np.float32 and v.shape[0] > 0:
        # float32: the contract is a fused multiply-add chain over k ascending (first term a plain
        # product).  The plain-C routine uses fmaf; emulating the fusion through float64 double-rounds
        # in about one of 1e7 operations, which is visible at 1 M points (one ulp of a coordinate of
        # magnitude 4 is 4.8e-7).
        from . import native

        return native.sample_points(w, v)
    out = np.zeros((v.shape[0], w.shape[0], v.shape[2]), dtype=v.dtype)
    for k in range(v.shape[1]):
        out = out + w[None, :, k, None] * v[:, None, k, :]
    return out


def simplices_by_dimension(stree, max_dimension: int) -> List[List[Tuple[int, ...]]]:
    """``flooder/core.py:135-138``."""
    buckets: List[List[Tuple[int, ...]]] = [[] for _ in range(max_dimension + 1)]
    for simplex, _ in stree.get_simplices():
        if len(simplex) <= max_dimension + 1:
            buckets[len(simplex) - 1].append(tuple(simplex))
    return buckets


def flood_complex(
    points,
    landmarks,
    max_dimension: Optional[int] = None,
    points_per_edge: Optional[int] = 30,
    num_rand: Optional[int] = None,
    return_simplex_tree: bool = False,
    start_idx: Optional[int] = 0,
    workers: int = 1,
    stats: Optional[dict] = None,
    injected_samples: Optional[Dict[int, np.ndarray]] = None,
):
    """CPU path of ``flooder/core.py:32-288`` (``landmarks.is_cpu`` branch).

    ``workers`` is forwarded to ``KDTree.query`` (the reference passes none, i.e. 1).
    ``stats``, when given, receives per-dimension intermediates (sorted simplices, sample
    points, per-sample distances, balls) for the differential tests.
    ``injected_samples[d]`` (S_d, R, D) replaces ``weights @ vertices`` so that the GPU
    path and the oracle can be compared on bit-identical sample points.
    """
    from scipy.spatial import KDTree

    pts = np.asarray(points)
    if isinstance(landmarks, (int, np.integer)):
        landmarks = generate_landmarks(pts, min(int(landmarks), pts.shape[0]), start_idx)
    lms = np.asarray(landmarks)
    if lms.dtype != pts.dtype:
        raise RuntimeError(f"landmarks.dtype ({lms.dtype}) != points.dtype ({pts.dtype})")
    if pts.dtype not in (np.float32, np.float64):
        raise TypeError(f"dtype ({pts.dtype}) not supported")
    if max_dimension is None:
        max_dimension = pts.shape[1]
    dtype = pts.dtype.type

    tree = KDTree(pts)                                           # core.py:128
    stree = DelaunayComplex(lms).create_simplex_tree()           # core.py:130-132
    buckets = simplices_by_dimension(stree, max_dimension)       # core.py:135-138
    axis = int(np.argmax(pts.max(axis=0) - pts.min(axis=0)))     # core.py:140-142

    out: Dict[Tuple[int, ...], float] = {}
    for d in range(max_dimension + 1):                           # core.py:146
        if num_rand is None and d < max_dimension:               # core.py:149-150
            continue
        if len(buckets[d]) == 0:
            continue
        d_simplices = np.asarray(buckets[d], dtype=np.int64)
        verts = lms[d_simplices]                                 # (S, d+1, D)
        centers, radii = bounding_balls(verts, d)                # core.py:156-172
        order = np.argsort(centers[:, axis], kind="stable")      # core.py:175-179
        verts, centers, radii, d_simplices = verts[order], centers[order], radii[order], d_simplices[order]

        if num_rand is None:                                     # core.py:182-187
            weights, vertex_idxs, face_idxs = generate_grid(points_per_edge, max_dimension, dtype)
        else:
            weights = generate_uniform_weights(num_rand, d, dtype)
        if injected_samples is not None and d in injected_samples:
            x = np.asarray(injected_samples[d])
        else:
            x = sample_points(weights, verts)                    # core.py:188
        dist, _ = tree.query(x.reshape(-1, x.shape[-1]), workers=workers)   # core.py:197-199
        dist = dist.reshape(x.shape[0], x.shape[1])

        if stats is not None:
            stats[d] = dict(simplices=d_simplices, vertices=verts, centers=centers, radii=radii,
                            weights=weights, samples=x, distances=dist)

        if num_rand is None:                                     # core.py:251-263
            for rows, vsel in zip(face_idxs, vertex_idxs):
                faces = d_simplices[:, vsel].reshape(-1, vsel.shape[1])
                vals = dist[:, rows].max(axis=2).reshape(-1)
                out.update(zip(map(tuple, faces.tolist()), vals.tolist()))
        else:                                                    # core.py:269-276
            vals = dist.max(axis=1)
            out.update(zip(map(tuple, d_simplices.tolist()), vals.tolist()))

    for simplex, value in out.items():                           # core.py:278-279
        stree.assign_filtration(simplex, value)
    stree.make_filtration_non_decreasing()                       # core.py:280
    if return_simplex_tree:
        return stree
    return dict((tuple(s), f) for s, f in stree.get_simplices())  # core.py:285-288


def algorithmic_evals(points: np.ndarray, landmarks: np.ndarray, max_dimension: Optional[int] = None,
                      points_per_edge: Optional[int] = 30, num_rand: Optional[int] = None) -> int:
    """Work count of SURVEY.md section 8(d): ``E = sum_s R_s * |ball(s) & cloud|`` over the
    simplices the call processes (grid mode: top dimension only, R = C(ppe+D-1, D);
    random mode: every dimension, R = num_rand)."""
    pts = np.asarray(points)
    lms = np.asarray(landmarks)
    if max_dimension is None:
        max_dimension = pts.shape[1]
    stree = DelaunayComplex(lms).create_simplex_tree()
    buckets = simplices_by_dimension(stree, max_dimension)
    total = 0
    for d in range(max_dimension + 1):
        if num_rand is None and d < max_dimension:
            continue
        if not buckets[d]:
            continue
        verts = lms[np.asarray(buckets[d], dtype=np.int64)]
        centers, radii = bounding_balls(verts, d)
        R = comb(points_per_edge + max_dimension - 1, max_dimension) if num_rand is None else num_rand
        total += int(R) * int(ball_candidate_counts(pts, centers, radii).sum())
    return total
