"""Host side of the Flood-complex path: the reference's public API on top of the sm_100a kernels.

``flood_complex`` and ``generate_landmarks`` keep the signatures, argument meaning, return types
and exceptions of ``flooder/core.py:32-43`` and ``:291-296`` of the reference.  What differs is
everything between "list of Delaunay simplices" and "per-simplex filtration values": it runs in
``libflood_b200.so`` (see ``include/flood_b200.h``) through the torch extension
``flooder_b200._native.ext()``.  There is no CPU, Triton or eager-PyTorch fallback: inputs that
are not on a CUDA device are rejected.
"""
from __future__ import annotations

import functools
import itertools
import os
import warnings
from numbers import Integral
from typing import Dict, List, Optional, Tuple, Union

import numpy as np
import torch

from . import _native
from . import distributed as fdist
from .bricks import brick_order
from .simplex_tree import FaceTable, SimplexTree, delaunay_complex, face_keys

_SUPPORTED_DTYPES = (torch.float32, torch.float64)

# Optional stage timing of flood_complex (diagnostics; adds a device synchronisation per stage and
# therefore removes the host/device overlap -- never enabled inside a timed benchmark region).
PROFILE_STAGES = os.environ.get("FLOODER_B200_PROFILE", "0") == "1"
# PROFILE_STAGES = "host": host-side time per stage without the synchronisations (the overlap stays;
# a stage that waits for the device shows the wait)
last_stage_seconds: Dict[str, float] = {}


class _Stage:
    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if PROFILE_STAGES:
            import time

            if PROFILE_STAGES != "host":
                torch.cuda.synchronize()
            self.t0 = time.perf_counter()
        return self

    def __exit__(self, *exc):
        if PROFILE_STAGES:
            import time

            if PROFILE_STAGES != "host":
                torch.cuda.synchronize()
            last_stage_seconds[self.name] = last_stage_seconds.get(self.name, 0.0) + time.perf_counter() - self.t0
        return False


# ----------------------------------------------------------------------------------------------
# sample-point generators (host, tiny) -- same weights as the reference
# ----------------------------------------------------------------------------------------------
@functools.lru_cache(maxsize=32)
def _lattice(n: int, dim: int) -> np.ndarray:
    """Integer barycentric lattice: all (dim+1)-part compositions of n-1 in the order of
    ``itertools.combinations`` over bar positions (reference ``core.py:369-380``)."""
    bars = np.fromiter(
        itertools.chain.from_iterable(itertools.combinations(range(n + dim - 1), dim)), dtype=np.int64
    ).reshape(-1, dim)
    edge = np.empty((bars.shape[0], dim + 2), dtype=np.int64)
    edge[:, 0] = -1
    edge[:, 1:-1] = bars
    edge[:, -1] = n + dim - 1
    return np.diff(edge, axis=1) - 1


def generate_grid(n: int, dim: int, device, dtype=torch.float32):
    """Grid of ``n`` points per edge on the unit ``dim``-simplex (reference ``core.py:346-402``).

    Returns ``(weights (C, dim+1), vertex_idxs, face_idxs)`` with the reference's meaning:
    ``face_idxs[k][j]`` are the rows whose weights vanish on the j-th ``k``-subset of vertices,
    ``vertex_idxs[k][j]`` the remaining vertex positions.
    """
    counts = _lattice(int(n), int(dim))
    axes = np.arange(dim + 1)
    face_idxs, vertex_idxs = [], []
    for k in range(dim + 1):
        rows_k, verts_k = [], []
        for zero_set in itertools.combinations(range(dim + 1), k):
            sel = np.ones(len(counts), dtype=bool) if k == 0 else (counts[:, list(zero_set)] == 0).all(axis=1)
            rows_k.append(torch.as_tensor(np.nonzero(sel)[0], device=device))
            verts_k.append(torch.as_tensor(axes[~np.isin(axes, zero_set)], device=device))
        face_idxs.append(torch.stack(rows_k))
        vertex_idxs.append(torch.stack(verts_k))
    # divided on the CPU and then moved: on a CUDA device torch.divide by a scalar multiplies by the
    # rounded reciprocal (one ulp off for some k); the CPU quotient is what the reference's CPU path
    # sees (see _grid_weights_cached)
    weights = torch.empty((counts.shape[0], dim + 1), dtype=dtype)
    torch.divide(torch.as_tensor(counts), n - 1, out=weights)
    return weights.to(device), vertex_idxs, face_idxs


def generate_uniform_weights(num_rand: int, dim: int, device, dtype=torch.float32) -> torch.Tensor:
    """Uniform (Dirichlet(1,..,1)) weights on the unit simplex (reference ``core.py:405-427``);
    drawn on the CPU from torch's global generator so that CPU and GPU callers see the same
    stream."""
    if dim == 0:
        return torch.ones((num_rand, 1), device=device, dtype=dtype)
    w = -torch.log(1 - torch.rand(num_rand, dim + 1)).to(device, dtype=dtype)
    return w / w.sum(dim=1, keepdim=True)


@functools.lru_cache(maxsize=16)
def _grid_weights_cached(n: int, dim: int, device_str: str) -> torch.Tensor:
    # The quotients k / (n - 1) are formed on the CPU (correctly rounded float32 division, what the
    # reference's CPU path uses) and then copied: on a CUDA device torch.divide by a
    # Python scalar multiplies by the rounded reciprocal, which is off by one ulp for some k -- one
    # ulp of a weight moves a sample point by up to one ulp of its coordinates, visible against the
    # exact KD-tree distances at 1 M points.
    counts = torch.as_tensor(_lattice(n, dim))
    weights = torch.empty(counts.shape, dtype=torch.float32)
    torch.divide(counts, n - 1, out=weights)
    weights = weights.to(device_str)
    weights._flood_cached = True
    return weights


def _grid_weights(n: int, dim: int, device, dtype=torch.float32) -> torch.Tensor:
    """The weights of ``generate_grid`` only (cached per device; treated as read-only)."""
    if dtype is torch.float64:
        counts = torch.as_tensor(_lattice(int(n), int(dim)))
        weights = torch.empty(counts.shape, dtype=torch.float64)
        torch.divide(counts, n - 1, out=weights)
        return weights.to(device)
    return _grid_weights_cached(int(n), int(dim), str(device))


def _support_masks(weights: torch.Tensor) -> torch.Tensor:
    """bit k of mask[r] <=> weights[r, k] != 0."""
    K = weights.shape[1]
    bits = (weights != 0).to(torch.int32) << torch.arange(K, device=weights.device, dtype=torch.int32)
    return bits.sum(dim=1).to(torch.int32).contiguous()


# Sample order handed to the kernel: every warp's brick of samples spatially compact (bricks.py).
# FLOODER_B200_BRICKS=0 keeps the reference's order (diagnostics).
USE_BRICKS = os.environ.get("FLOODER_B200_BRICKS", "1") != "0"
_brick_cache: Dict[tuple, tuple] = {}


def _kernel_sample_order(weights: torch.Tensor, dim: int, cache_key=None):
    """``(perm, weights[perm])`` for the brick layout the kernel reports for these options, or
    ``(None, weights)``.  ``cache_key`` identifies read-only weights (the cached lattice)."""
    R, K = weights.shape
    if not USE_BRICKS or K < 2 or R <= 32:
        return None, weights
    groups, per_block = _native.ext().covering_bricks(int(R), int(dim))
    key = None
    if cache_key is not None:
        key = (cache_key, str(weights.device), tuple(groups), per_block)
        hit = _brick_cache.get(key)
        if hit is not None:
            return hit
    perm_np = brick_order(weights.detach().cpu().numpy(), groups, per_block)
    perm = torch.as_tensor(perm_np, device=weights.device)
    out = (perm, weights[perm].contiguous())
    if key is not None:
        if len(_brick_cache) > 64:
            _brick_cache.clear()
        _brick_cache[key] = out
    return out


# ----------------------------------------------------------------------------------------------
# landmarks
# ----------------------------------------------------------------------------------------------
def _require_cuda(t: torch.Tensor, what: str) -> None:
    if t.device.type != "cuda":
        raise RuntimeError(
            f"Device not supported: {what} is on '{t.device}'. flooder_b200 runs on CUDA (sm_100a) "
            "devices only and has no CPU fallback."
        )


def _fps_register_capacity(dim: int) -> int:
    """Clouds up to this size stay register-resident in the brute-force FPS kernel (148 SMs x 1024
    threads x 8 points for D <= 4, x 4 points above); measured on B200 it then needs ~4 us per
    landmark, which the bucketed kernel only beats on larger clouds."""
    return 148 * 1024 * (8 if dim <= 4 else 4)



def fps_indices(points: torch.Tensor, n_lms: int, start_idx: int = 0, method: str = "auto",
                cloud: Optional["PreparedCloud"] = None) -> torch.Tensor:
    """Indices chosen by exact farthest-point sampling, int64 on ``points.device``.

    ``method``: ``"brute"`` = every point every iteration (``flood_fps_f32``), ``"grid"`` = bucketed
    on the cell grid of a prepared cloud (``flood_fps_grid_f32``), ``"auto"`` picks by size.  Both
    return the same indices bit for bit."""
    _require_cuda(points, "points")
    pts32 = points.detach().to(torch.float32).contiguous()
    if method == "auto":
        n, dim = pts32.shape
        method = "grid" if (dim >= 2 and n > _fps_register_capacity(dim)) else "brute"
    if method == "brute":
        return _native.ext().fps(pts32, int(n_lms), int(start_idx))
    if method != "grid":
        raise ValueError(f"unknown FPS method {method!r}")
    if cloud is None:
        cloud = PreparedCloud(pts32)
    return _native.ext().fps_grid(cloud.workspace, pts32, int(n_lms), int(start_idx))


def generate_landmarks(
    points: torch.Tensor,
    n_lms: int,
    fps_h: Union[None, int] = None,
    start_idx: Union[int, None] = None,
    *,
    _cloud: Optional["PreparedCloud"] = None,
) -> torch.Tensor:
    """Selects landmarks using farthest-point sampling.

    Same contract as the reference (``flooder/core.py:291-343``): returns ``points[idx]`` for the
    FPS index sequence ``idx`` on the device and in the dtype of ``points``.  The sampling itself
    is the persistent sm_100a kernel of ``csrc/fps.cu`` (exact FPS, float32).  ``fps_h`` (the
    KD-tree height of the reference's bucket-FPS) is accepted for compatibility and has no
    effect on the result.  ``start_idx=None`` picks the start index with ``np.random``.
    """
    del fps_h
    if n_lms <= 0:
        raise RuntimeError(f"Number of landmarks ({n_lms}) must be positive")
    n_pts = len(points)
    n_lms = min(int(n_lms), n_pts)
    if start_idx is None:
        start_idx = int(np.random.randint(n_pts))
    dev_points = points
    if points.device.type != "cuda":
        if not torch.cuda.is_available():
            _require_cuda(points, "points")
        dev_points = points.cuda()
    index_set = fps_indices(dev_points, n_lms, start_idx, cloud=_cloud).to(points.device)
    return points[index_set]


# ----------------------------------------------------------------------------------------------
# the filtration
# ----------------------------------------------------------------------------------------------
class PreparedCloud:
    """Cell-sorted device copy of a point cloud (``flood_cloud_build_f32``)."""

    def __init__(self, points: torch.Tensor, points_per_cell: int = 0):
        _require_cuda(points, "points")
        pts32 = points.detach().to(torch.float32).contiguous()
        self.n, self.d = int(pts32.shape[0]), int(pts32.shape[1])
        self.device = pts32.device
        self.workspace = _native.ext().cloud_build(pts32, int(points_per_cell))


def _slab_rows(S: int, R: int, device) -> int:
    """Simplices per kernel call such that the (rows, R) float32 ``min_dist2`` buffer stays within
    a quarter of the free device memory (at most 8 GiB).  The reference bounds the same buffer with
    ``batch_size`` (``flooder/core.py:193-226``); here one call normally takes every simplex."""
    override = os.environ.get("FLOODER_B200_SLAB_BYTES")
    if override:
        budget = int(override)
    elif 4 * S * max(R, 1) <= (2 << 30):
        return S            # small enough not to need a look at the free memory (cudaMemGetInfo stalls
                            # the submitting thread for tens of milliseconds every few calls)
    else:
        free, _total = torch.cuda.mem_get_info(device)
        budget = min(free // 4, 8 << 30)
    return max(1, min(S, int(budget // (4 * max(R, 1)))))


def covering_values(
    cloud: PreparedCloud,
    simplex_vertices: torch.Tensor,
    weights: torch.Tensor,
    grid_mode: bool,
    samples: Optional[torch.Tensor] = None,
    return_details: bool = False,
    stats: Optional[dict] = None,
):
    """Device part of one dimension pass: bounding balls -> covering-radius kernel -> face maxima.

    Returns a float32 tensor ``(S, 2^K - 1)`` (grid mode; column ``m-1`` belongs to the face whose
    vertex positions are the set bits of ``m``) or ``(S, 1)`` (random mode).  ``stats``, when
    given, receives the device-side work counters of the call (``evals`` = algorithmic count E,
    ``executed`` = evaluations actually performed; summed over slabs) without changing what is
    enqueued.
    """
    ext = _native.ext()
    verts = simplex_vertices.to(torch.float32).contiguous()
    w = weights.to(torch.float32).contiguous()
    S, K = verts.shape[0], verts.shape[1]
    perm = None
    if samples is None:
        # the lattice weights are cached per device and never modified: key the order by identity
        cache_key = ("lattice", tuple(w.shape), w.data_ptr()) if getattr(weights, "_flood_cached", False) else None
        perm, w = _kernel_sample_order(w, cloud.d, cache_key)
    support = _support_masks(w) if grid_mode else None
    centers, radii = ext.bounding_balls(verts)
    # Largest balls first: the kernel's work queue follows the simplex order, so the big items
    # are dealt out early and the small ones fill the tail of the launch.
    order = None
    if samples is None and not return_details and S > 1:
        order = torch.argsort(radii, descending=True)
        verts, centers, radii = verts[order].contiguous(), centers[order].contiguous(), radii[order].contiguous()
    rows = S if return_details else _slab_rows(S, w.shape[0], verts.device)
    if rows >= S:
        min_d2, counts, evals, executed = ext.covering_radius(cloud.workspace, cloud.n, cloud.d, verts, w, samples,
                                                              centers, radii)
        values = ext.face_max(min_d2, support, K)
        if stats is not None:
            stats["evals"], stats["executed"] = evals, executed
    else:
        # memory-bounded: the simplices go through the kernel in slabs, only the face values stay
        parts = []
        for lo in range(0, S, rows):
            hi = min(S, lo + rows)
            min_d2, _c, evals, executed = ext.covering_radius(cloud.workspace, cloud.n, cloud.d, verts[lo:hi], w,
                                                              None if samples is None else samples[lo:hi].contiguous(),
                                                              centers[lo:hi], radii[lo:hi])
            parts.append(ext.face_max(min_d2, support, K))
            del min_d2
            if stats is not None:
                stats["evals"] = evals if lo == 0 else stats["evals"] + evals
                stats["executed"] = executed if lo == 0 else stats["executed"] + executed
        values = torch.cat(parts, dim=0)
    if order is not None:
        values = torch.empty_like(values).index_copy_(0, order, values)
    if return_details:
        if perm is not None:            # per-sample output back in the caller's sample order
            min_d2 = torch.empty_like(min_d2).index_copy_(1, perm, min_d2)
        return values, dict(min_dist2=min_d2, cand_count=counts, evals=evals, executed=executed,
                            centers=centers, radii=radii)
    return values


def covering_cost(cloud: PreparedCloud, simplex_vertices: torch.Tensor) -> torch.Tensor:
    """Per-simplex cost estimate (length of the candidate stream, ``flood_covering_plan_f32``)."""
    ext = _native.ext()
    centers, radii = ext.bounding_balls(simplex_vertices.to(torch.float32).contiguous())
    return ext.covering_plan(cloud.workspace, cloud.n, cloud.d, centers, radii).to(torch.float32)


def covering_values_f64(cloud: PreparedCloud, points64: torch.Tensor, simplex_vertices: torch.Tensor,
                        weights: torch.Tensor, grid_mode: bool, stats: Optional[dict] = None) -> torch.Tensor:
    """float64 variant of ``covering_values`` (the reference evaluates float64 inputs in float64,
    ``flooder/triton_kernels.py:226-229``): bounding balls, ball predicate, sample points and
    distances in float64 on ``points64``; ``cloud`` (the float32-rounded prepared cloud) only
    enumerates candidates.  Plain kernels, no pruning (``csrc/f64.cu``)."""
    ext = _native.ext()
    verts = simplex_vertices.to(torch.float64).contiguous()
    w = weights.to(torch.float64).contiguous()
    S, K = verts.shape[0], verts.shape[1]
    support = _support_masks(w) if grid_mode else None
    centers, radii = ext.bounding_balls_f64(verts)
    rows = _slab_rows(S, 2 * w.shape[0], verts.device)
    parts = []
    for lo in range(0, S, rows):
        hi = min(S, lo + rows)
        min_d2, _counts, evals = ext.covering_radius_f64(cloud.workspace, points64, verts[lo:hi].contiguous(), w,
                                                         centers[lo:hi].contiguous(), radii[lo:hi].contiguous())
        parts.append(ext.face_max_f64(min_d2, support, K))
        del min_d2
        if stats is not None:
            stats["evals"] = evals if lo == 0 else stats["evals"] + evals
            stats["executed"] = stats["evals"]
    return parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)


def device_pass(cloud: PreparedCloud, simplex_vertices: torch.Tensor, weights: torch.Tensor, grid_mode: bool,
                shard: Optional["fdist.Shard"] = None, stats: Optional[dict] = None,
                points64: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Everything ``flood_complex`` enqueues on the device for one dimension pass, given the
    prepared cloud: [cost plan + partition over ranks] -> bounding balls -> covering-radius kernel
    -> face maxima -> [all-gather of the per-simplex values].  Asynchronous; returns the device
    tensor of values for all simplices.  ``bench.py`` times exactly this function (plus the cloud
    build), so the benchmarked path cannot drift from the product path."""
    if points64 is not None:
        def compute(v):
            return covering_values_f64(cloud, points64, v, weights, grid_mode, stats=stats)
    else:
        def compute(v):
            return covering_values(cloud, v, weights, grid_mode=grid_mode, stats=stats)
    if shard is None:
        return compute(simplex_vertices)
    return fdist.sharded_covering_values(shard, simplex_vertices, compute,
                                         cost=covering_cost(cloud, simplex_vertices))


def _scatter_face_values(table: FaceTable, cell_values: np.ndarray, values: Dict[int, np.ndarray]) -> None:
    """Grid mode: spread the (S, 2^K-1) per-cell face values over the unique faces.  Column
    ``m-1`` of ``cell_values`` belongs to the face made of the vertex positions set in ``m``.  A
    face shared by several cells is computed once per coface from bit-identical sample points;
    the smallest value is kept (it has seen the union of the cofaces' candidate balls)."""
    for k, combos in table.combos.items():
        cols = [sum(1 << p for p in combo) - 1 for combo in combos]
        best = np.full(len(table.faces[k]), np.inf)
        np.minimum.at(best, table.cell_face[k].reshape(-1), cell_values[:, cols].astype(np.float64).reshape(-1))
        values[k] = best


def flood_complex(
    points: torch.Tensor,
    landmarks: Union[int, torch.Tensor],
    max_dimension: Union[None, int] = None,
    points_per_edge: Union[None, int] = 30,
    num_rand: int = None,
    batch_size: Union[None, int] = 64,
    use_triton: Optional[bool] = None,
    return_simplex_tree: bool = False,
    fps_h: Union[None, int] = None,
    start_idx: Union[int, None] = 0,
) -> Union[dict, "SimplexTree"]:
    """Constructs a Flood complex from witness points and landmarks.

    Drop-in for the reference's ``flooder.flood_complex`` (``flooder/core.py:32-288``): same
    arguments, same result (``{tuple(landmark indices): covering radius}`` for every simplex of the
    Delaunay complex of the landmarks, or the simplex tree itself).  ``points`` and ``landmarks``
    must live on a CUDA device.  ``batch_size``, ``use_triton`` and ``fps_h`` are accepted for
    compatibility and ignored: there is no batching (the kernel never materialises a mask) and
    no Triton path.  float64 inputs are accepted with the reference's ``RuntimeWarning`` and
    computed in float32.

    When ``torch.distributed`` is initialised with more than one rank, the simplex list is
    sharded over the ranks (cloud replicated, see ``flooder_b200.distributed``) and the values
    are all-gathered, so every rank returns the complete complex.
    """
    del batch_size, use_triton
    if max_dimension is None:
        max_dimension = points.shape[1]
    if points.dim() != 2 or points.shape[1] < 1 or points.shape[1] > 8:
        raise RuntimeError(f"points must have shape (N, D) with 1 <= D <= 8, got {tuple(points.shape)}")
    if PROFILE_STAGES:
        last_stage_seconds.clear()
    cloud = None
    landmarks_arg = landmarks
    if isinstance(landmarks, Integral):
        if points.device.type == "cuda" and points.dtype in _SUPPORTED_DTYPES:
            with _Stage("cloud_build"):
                cloud = PreparedCloud(points)      # shared by the bucketed FPS and the covering pass
        if start_idx is None and fdist.current_shard() is not None:
            # a random start index is drawn by rank 0 and shared: every rank must select the
            # same landmarks (FPS itself is deterministic)
            start_idx = fdist.broadcast_int(fdist.current_shard(), int(np.random.randint(len(points))), points.device)
        with _Stage("fps"):
            landmarks = generate_landmarks(points, min(landmarks, points.shape[0]), fps_h, start_idx=start_idx,
                                           _cloud=cloud)
    if landmarks.device != points.device:
        raise RuntimeError(f"landmarks.device ({landmarks.device}) != points.device ({points.device})")
    if landmarks.dtype != points.dtype:
        raise RuntimeError(f"landmarks.dtype ({landmarks.dtype}) != points.device ({points.dtype})")
    device, dtype = points.device, points.dtype
    if dtype not in _SUPPORTED_DTYPES:
        raise TypeError(f"dtype ({dtype}) not supported")
    if dtype is torch.float64:
        warnings.warn(
            "Using float64 kernels is slow on B200 (plain FP64 kernels, no pruning)",
            RuntimeWarning,
            stacklevel=2,
        )
    _require_cuda(points, "points")
    torch.cuda.set_device(device)

    shard = fdist.current_shard()
    del landmarks_arg
    lms32 = landmarks.detach().to(torch.float32)
    # float64 inputs are evaluated in float64 (as the reference does): original coordinates
    points64 = points.detach().contiguous() if dtype is torch.float64 else None
    lms64 = landmarks.detach().contiguous() if dtype is torch.float64 else None
    with _Stage("delaunay"):
        # host, as in the reference; triangulated in the precision the landmarks were given in
        cells, gudhi_tree = delaunay_complex(landmarks.detach().cpu().numpy())
    K = cells.shape[1]
    grid_mode = num_rand is None
    max_dimension = min(max_dimension, K - 1)          # degenerate inputs have lower-dimensional cells
    if cloud is None:
        with _Stage("cloud_build"):
            cloud = PreparedCloud(points)
    if shard is not None:
        # every rank must hold the same complex, or the all-gather below mixes rows up
        digest_input = np.concatenate([np.asarray(cells.shape, dtype=np.int64), cells.reshape(-1),
                                       lms32.cpu().numpy().view(np.int32).reshape(-1).astype(np.int64)])
        fdist.assert_same_on_all_ranks(shard, digest_input, device, "landmarks / Delaunay cells")

    def launch(d_simplices_np: np.ndarray, weights: torch.Tensor) -> torch.Tensor:
        """Enqueue one dimension pass (asynchronous); returns the device tensor of values."""
        index = torch.as_tensor(d_simplices_np, device=device)
        if points64 is not None:
            return device_pass(cloud, lms64[index], weights, grid_mode, shard, points64=points64)
        return device_pass(cloud, lms32[index], weights, grid_mode, shard)

    # Grid mode on full-dimensional cells needs nothing but the cells: enqueue the kernels first
    # and build the face table on the host while the GPU works.
    pending = None
    if grid_mode and max_dimension == K - 1 and cells.shape[0] > 0:
        with _Stage("kernels"):
            pending = launch(cells, _grid_weights(points_per_edge, max_dimension, device, dtype))
    with _Stage("face_table"):
        table = FaceTable(cells, n_vertices=lms32.shape[0])
        values = table.nan_values()       # NaN = not assigned (simplices above max_dimension)
    key_lists = None
    if pending is not None and gudhi_tree is None:
        with _Stage("face_table"):
            key_lists = face_keys(table.faces)      # value-independent part of the result, built while the GPU works
    if pending is not None:
        with _Stage("d2h_scatter"):
            _scatter_face_values(table, pending.cpu().numpy(), values)
    else:
        for d in range(max_dimension + 1):
            if grid_mode and d < max_dimension:
                continue
            # grid mode evaluates the max_dimension-faces and reads all lower faces off the same
            # samples; random mode evaluates every dimension on its own.  Rows follow
            # table.faces[d + 1] (sorted, unique), the order values[d + 1] is indexed in.
            d_cells = table.faces[d + 1]
            if d_cells.shape[0] == 0:
                continue
            if grid_mode:
                sub = FaceTable(d_cells, n_vertices=table.base)
                host_values = launch(d_cells, _grid_weights(points_per_edge, max_dimension, device, dtype)).cpu().numpy()
                _scatter_face_values(sub, host_values, values)
            else:
                weights = generate_uniform_weights(num_rand, d, device, dtype)
                if shard is not None:
                    # the weights come from each process's own CPU generator (core.py:423-425 of the
                    # reference): rank 0's draw is used everywhere so that a value does not depend
                    # on which rank evaluated it
                    fdist.broadcast_tensor(shard, weights)
                values[d + 1] = launch(d_cells, weights).cpu().numpy()[:, 0].astype(np.float64)

    with _Stage("assemble"):
        # grid mode on full-dimensional cells is monotone by construction (face samples are a subset
        # of coface samples, and the minimum over cofaces preserves that); everything else goes
        # through make_filtration_non_decreasing like the reference (core.py:280)
        monotone = pending is not None
        return _write_back(table, values, gudhi_tree, return_simplex_tree, monotone, key_lists)


def _write_back(table: FaceTable, values: Dict[int, np.ndarray], gudhi_tree, return_simplex_tree: bool,
                monotone: bool = False, key_lists: Optional[Dict[int, list]] = None):
    """Reference ``core.py:278-288``: assign the values, make the filtration non-decreasing, return
    the tree or ``{tuple(simplex): value}``.  With gudhi the container is gudhi's own tree (the
    reference's contract); otherwise the array-backed stand-in."""
    if gudhi_tree is not None:
        stree = gudhi_tree
        for k, faces in table.faces.items():
            for simplex, value in zip(faces.tolist(), values[k].tolist()):
                if value == value:          # NaN = not assigned (above max_dimension)
                    stree.assign_filtration(simplex, value)
        stree.make_filtration_non_decreasing()
        if return_simplex_tree:
            return stree
        return dict((tuple(simplex), filtr) for (simplex, filtr) in stree.get_simplices())
    if not monotone:
        table.make_non_decreasing(values)
    stree = SimplexTree.from_arrays(table.faces, values, keys=key_lists)
    if return_simplex_tree:
        return stree
    return stree._f          # the tree is dropped: hand its dictionary over instead of copying it
