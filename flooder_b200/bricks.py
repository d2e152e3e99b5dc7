"""Sample ordering for the covering-radius kernel.

The kernel keeps the sample points of a simplex in registers, one *brick* (up to 8 groups of 32
consecutive samples) per warp, and prunes candidate cloud points against the bounding box of each
brick (``csrc/covering_kernels.cuh``).  The reference lists the barycentric lattice in
lexicographic order (``flooder/core.py:369-380``), where 256 consecutive rows form a thin slab
with a large box.  The sample weights are shared by all simplices (``core.py:182-188``) and the
result does not depend on their order, so the host is free to permute them: ``brick_order``
returns a permutation under which every brick -- and every CTA's block of bricks -- is a compact
piece of the simplex (recursive bisection along the widest barycentric axis, honouring the brick
sizes reported by ``flood_covering_bricks``).  Per-face maxima only need the support masks to be
permuted alongside; per-sample outputs are un-permuted on request.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def split_directions(k: int) -> np.ndarray:
    """Unit directions (rows) in barycentric coordinates along which a set of samples may be cut:
    the K altitudes and the K(K-1)/2 edge directions of the regular simplex (the barycentric
    coordinates are an isometric picture of it, up to scale)."""
    rows = []
    for i in range(k):
        v = np.full(k, -1.0 / k)
        v[i] += 1.0
        rows.append(v / np.linalg.norm(v))
    for i in range(k):
        for j in range(i + 1, k):
            v = np.zeros(k)
            v[i], v[j] = 1.0, -1.0
            rows.append(v / np.sqrt(2.0))
    return np.asarray(rows)


def _bisect(coords: np.ndarray, idx: np.ndarray, sizes: Sequence[int], out: List[np.ndarray]) -> None:
    """Order ``idx`` so that consecutive segments of the given sizes are spatially compact.
    ``coords`` holds the projections of the samples on the candidate cut directions."""
    if len(sizes) == 1 or idx.size == 0:
        out.append(idx)
        return
    half = len(sizes) // 2
    n_left = int(sum(sizes[:half]))
    pts = coords[idx]
    axis = int(np.argmax(pts.max(axis=0) - pts.min(axis=0)))
    order = idx[np.argsort(pts[:, axis], kind="stable")]
    _bisect(coords, order[:n_left], sizes[:half], out)
    _bisect(coords, order[n_left:], sizes[half:], out)


def _segments(total: int, unit: int) -> List[int]:
    """``total`` split into pieces of ``unit`` (the last one may be short)."""
    full, rest = divmod(total, unit)
    return [unit] * full + ([rest] if rest else [])


def brick_order(weights: np.ndarray, brick_groups: Sequence[int], bricks_per_block: int, group: int = 32) -> np.ndarray:
    """Permutation ``perm`` (int64, length R): ``weights[perm]`` is the order to hand to the kernel.

    ``brick_groups[i]`` = number of ``group``-sample groups of brick i, ``bricks_per_block``
    consecutive bricks form one CTA's sample block (``flood_covering_bricks``).  Nested
    bisections: sample blocks, bricks inside a block, pairs of groups inside a brick, the two
    groups of a pair."""
    w = np.asarray(weights, dtype=np.float64)
    R = w.shape[0]
    w = np.round(w @ split_directions(w.shape[1]).T, 9)      # rounding keeps lattice ties exact
    sizes, left = [], R
    for g in brick_groups:                        # samples per brick; the tail brick may be short
        take = min(int(g) * group, left)
        sizes.append(take)
        left -= take
    if left != 0:
        raise ValueError(f"brick layout covers {R - left} of {R} samples")
    blocks = [sizes[i:i + bricks_per_block] for i in range(0, len(sizes), bricks_per_block)]
    block_idx: List[np.ndarray] = []
    _bisect(w, np.arange(R, dtype=np.int64), [sum(b) for b in blocks], block_idx)
    pieces: List[np.ndarray] = []
    for idx_b, bricks in zip(block_idx, blocks):
        brick_idx: List[np.ndarray] = []
        _bisect(w, idx_b, bricks, brick_idx)
        for idx_k in brick_idx:
            if idx_k.size:
                # pairs of groups first (the kernel's second-level boxes are those of groups
                # (0,1), (2,3), ... of a brick), then the two groups of a pair
                pair_idx: List[np.ndarray] = []
                _bisect(w, idx_k, _segments(idx_k.size, 2 * group), pair_idx)
                for idx_p in pair_idx:
                    _bisect(w, idx_p, _segments(idx_p.size, group), pieces)
    perm = np.concatenate(pieces) if pieces else np.zeros(0, dtype=np.int64)
    assert perm.size == R
    return perm
