"""GPU parity tests of the individual kernels, through the C ABI (ctypes and the torch loader),
against the CPU oracle on the same seeded inputs.  Bit-exact unless stated."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import flood_oracle, native
from tests.helpers import seed_all

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ext():
    from flooder_b200 import _native

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return _native.ext()


def _cloud(kind, n, d, seed=0):
    g = torch.Generator().manual_seed(seed)
    if kind == "gauss":
        return torch.randn(n, d, generator=g)
    if kind == "uniform":
        return torch.rand(n, d, generator=g)
    if kind == "torus":
        import flooder_b200 as fb

        torch.manual_seed(seed)
        return fb.generate_noisy_torus_points_3d(n)
    raise ValueError(kind)


# ------------------------------------------------------------------------------------------
# FPS: indices bit-exact
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,d,n_lms,start", [
    (1000, 2, 64, 0), (1000, 3, 1000, 7), (10_000, 3, 100, 0), (50_000, 3, 500, 123),
    (200_000, 3, 300, 0), (20_000, 5, 128, 5), (5_000, 6, 64, 0), (300, 1, 10, 0), (1, 3, 1, 0),
])
def test_fps_indices(ext, n, d, n_lms, start):
    pts = _cloud("gauss", n, d, seed=n + d)
    want = native.fps(pts.numpy(), n_lms, start)
    got = ext.fps(pts.cuda(), n_lms, start).cpu().numpy()
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("kind,n,d,n_lms,start,ppc", [
    ("gauss", 1000, 2, 64, 0, 0), ("gauss", 1000, 3, 1000, 7, 0), ("torus", 50_000, 3, 500, 123, 0),
    ("gauss", 200_000, 3, 300, 0, 0), ("uniform", 20_000, 5, 128, 5, 0), ("uniform", 5_000, 6, 64, 0, 0),
    ("uniform", 3_000, 8, 40, 1, 0), ("torus", 30_000, 3, 200, 0, 1), ("torus", 30_000, 3, 200, 0, 500),
    ("gauss", 400_000, 4, 100, 0, 2), ("gauss", 2, 3, 2, 1, 0),
])
def test_fps_grid_indices(ext, kind, n, d, n_lms, start, ppc):
    """Bucketed FPS on the cell grid == exact FPS, for several cell sizes (ppc = points per cell;
    1 -> many lanes' worth of cells per warp, 500 -> a handful of big cells)."""
    pts = _cloud(kind, n, d, seed=n + d)
    want = native.fps(pts.numpy(), n_lms, start)
    dev = pts.cuda()
    ws = ext.cloud_build(dev, ppc)
    got = ext.fps_grid(ws, dev, n_lms, start).cpu().numpy()
    np.testing.assert_array_equal(got, want)


def test_fps_grid_duplicates(ext):
    base = _cloud("uniform", 50, 3, seed=9)
    pts = base.repeat(4, 1)
    want = native.fps(pts.numpy(), 120, 0)
    dev = pts.cuda()
    ws = ext.cloud_build(dev, 4)
    np.testing.assert_array_equal(ext.fps_grid(ws, dev, 120, 0).cpu().numpy(), want)


def test_fps_streaming_mode(ext):
    """Clouds beyond the register-resident capacity use the global-scratch variant."""
    pts = _cloud("gauss", 30_000, 3, seed=3)
    want = native.fps(pts.numpy(), 200, 0)
    ext.set_option("fps_stream", 1)
    try:
        got = ext.fps(pts.cuda(), 200, 0).cpu().numpy()
    finally:
        ext.set_option("fps_stream", 0)
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("mode", [0, 1])
def test_fps_exchange_modes(ext, mode):
    """Grid-wide argmax via cooperative-groups grid sync (0, the default) or the counter barrier (1)."""
    pts = _cloud("gauss", 40_000, 3, seed=4)
    want = native.fps(pts.numpy(), 150, 0)
    dev = pts.cuda()
    prev = ext.set_option("fps_barrier", mode)
    try:
        got = ext.fps(dev, 150, 0).cpu().numpy()
        ws = ext.cloud_build(dev, 0)
        got_grid = ext.fps_grid(ws, dev, 150, 0).cpu().numpy()
        ext.set_option("fps_stream", 1)
        got_stream = ext.fps(dev, 150, 0).cpu().numpy()
    finally:
        ext.set_option("fps_stream", 0)
        ext.set_option("fps_barrier", prev)
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(got_grid, want)
    np.testing.assert_array_equal(got_stream, want)


def test_fps_duplicates(ext):
    """More landmarks than distinct points: ties resolve to the first index, as in the oracle."""
    base = _cloud("uniform", 50, 3, seed=9)
    pts = base.repeat(4, 1)
    want = native.fps(pts.numpy(), 120, 0)
    got = ext.fps(pts.cuda(), 120, 0).cpu().numpy()
    np.testing.assert_array_equal(got, want)


# ------------------------------------------------------------------------------------------
# bounding balls
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d,K", [(2, 3), (3, 4), (3, 2), (3, 1), (5, 6), (8, 9), (1, 2), (6, 4)])
def test_bounding_balls(ext, d, K):
    pts = _cloud("gauss", 500, d, seed=d * 10 + K)
    g = torch.Generator().manual_seed(1)
    verts = pts[torch.randint(0, 500, (3000, K), generator=g)]
    c, r = ext.bounding_balls(verts.cuda().contiguous())
    c0, r0 = flood_oracle.bounding_balls(verts.numpy(), K - 1)
    # bit-exact: same operation sequence, first maximum of the flattened K x K matrix wins
    # (the random vertex picks contain repeated vertices, i.e. exact ties between edges)
    np.testing.assert_array_equal(c.cpu().numpy(), c0)
    np.testing.assert_array_equal(r.cpu().numpy(), r0)


# ------------------------------------------------------------------------------------------
# covering radius: per-sample minima and candidate counts, bit-exact vs the brute-force oracle
# ------------------------------------------------------------------------------------------
def _covering_case(ext, pts, verts, weights, samples=None, ppc=0):
    dev = torch.device("cuda")
    P = pts.to(dev).contiguous()
    V = verts.to(dev).contiguous()
    W = weights.to(dev).contiguous()
    ws = ext.cloud_build(P, ppc)
    c, r = ext.bounding_balls(V)
    smp = None if samples is None else samples.to(dev).contiguous()
    md2, cnt, ev, executed = ext.covering_radius(ws, P.shape[0], P.shape[1], V, W, smp, c, r)
    torch.cuda.synchronize()
    return md2.cpu().numpy(), cnt.cpu().numpy(), int(ev.item()), c.cpu().numpy(), r.cpu().numpy()


_ORACLE_CACHE = {}


def _check_against_bruteforce(pts, verts, weights, md2, cnt, ev, c, r, samples=None, cache_key=None):
    """``cache_key``: the oracle's answer is kept per key (the option tests run one input under
    many kernel configurations; the balls (c, r) are compared before a cached answer is used)."""
    hit = _ORACLE_CACHE.get(cache_key) if cache_key is not None else None
    if hit is not None and np.array_equal(hit[0], c) and np.array_equal(hit[1], r):
        want, want_cnt = hit[2], hit[3]
    else:
        x = native.sample_points(weights.numpy(), verts.numpy()) if samples is None else samples.numpy()
        want = native.min_dist(pts.numpy(), x, c, r)            # sqrt(min d2) restricted to the ball
        want_cnt = native.ball_counts(pts.numpy(), c, r)
        if cache_key is not None:
            _ORACLE_CACHE[cache_key] = (c.copy(), r.copy(), want, want_cnt)
    got = np.sqrt(md2)
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(cnt, want_cnt)
    assert ev == int(want_cnt.sum()) * weights.shape[0]


@pytest.mark.parametrize("kind,n,d,S,ppe", [
    ("torus", 5000, 3, 60, 8), ("gauss", 20000, 3, 40, 12), ("uniform", 3000, 2, 50, 20),
    ("uniform", 4000, 4, 30, 5), ("uniform", 3000, 5, 20, 4), ("uniform", 2500, 6, 12, 3),
    ("gauss", 777, 3, 9, 30), ("torus", 4000, 3, 20, 20), ("uniform", 3000, 2, 10, 130),
    ("uniform", 2000, 3, 10, 2), ("uniform", 2000, 5, 10, 6), ("uniform", 2000, 6, 10, 4),
    ("uniform", 1500, 7, 6, 3), ("uniform", 1500, 8, 6, 2),
])
def test_covering_bruteforce(ext, kind, n, d, S, ppe):
    pts = _cloud(kind, n, d, seed=n)
    g = torch.Generator().manual_seed(2)
    lms = pts[torch.randperm(n, generator=g)[: max(4 * (d + 1), 30)]]
    cells = flood_oracle.delaunay_top_simplices(lms.numpy()) if False else None
    from oracle.simplex_tree import delaunay_top_simplices

    cells = delaunay_top_simplices(lms.numpy())[:S]
    verts = lms[torch.as_tensor(cells)]
    w = torch.as_tensor(flood_oracle.generate_grid(ppe, d)[0])
    md2, cnt, ev, c, r = _covering_case(ext, pts, verts, w)
    _check_against_bruteforce(pts, verts, w, md2, cnt, ev, c, r)


@pytest.mark.parametrize("option,value", [("chunk", 256), ("warps", 8), ("warps", 3), ("warps", 13),
                                          ("tile_cap", 700), ("ctas_per_sm", 1), ("warps", 4), ("rows_per_chunk_factor", 0),
                                          ("points_per_cell", 1), ("points_per_cell", 64),
                                          ("level2", 0), ("level2", 1), ("slab_cull", 0),
                                          ("flush", 128), ("l2_bypass", 0), ("l2_bypass", 9),
                                          ("seg", 64)])
def test_covering_options(ext, option, value):
    """Chunk splitting (atomicMin merge), other CTA shapes (several sample blocks, uneven
    groups per warp, odd group counts -> scalar leftover path), small tiles, extreme cell
    sizes and every variant of the pruned sweep (one / two levels, with and without prefetch,
    slab culling, whole-brick bypass always / never) give the same bits."""
    pts = _cloud("torus", 30000, 3, seed=11)
    g = torch.Generator().manual_seed(5)
    lms = pts[torch.randperm(30000, generator=g)[:40]]
    from oracle.simplex_tree import delaunay_top_simplices

    cells = delaunay_top_simplices(lms.numpy())[:25]
    verts = lms[torch.as_tensor(cells)]
    w = torch.as_tensor(flood_oracle.generate_grid(30, 3)[0])
    prev = ext.set_option(option, value)
    try:
        md2, cnt, ev, c, r = _covering_case(ext, pts, verts, w)
    finally:
        ext.set_option(option, prev)
    _check_against_bruteforce(pts, verts, w, md2, cnt, ev, c, r, cache_key="covering_options")


@pytest.mark.parametrize("kind,n,d,ppe", [("torus", 60_000, 3, 30), ("gauss", 50_000, 3, 12), ("uniform", 20_000, 2, 40),
                                         ("uniform", 20_000, 5, 4),
                                         # wide shape (more than two bricks: second pruning level) in 2-D with two
                                         # sample blocks, 4-D, and with the 32-byte records of 5-D / 6-D
                                         ("uniform", 20_000, 2, 130), ("uniform", 20_000, 4, 10),
                                         ("uniform", 15_000, 5, 8), ("uniform", 10_000, 6, 7)])
def test_pruning_is_exact(ext, kind, n, d, ppe):
    """Pruned sweep (with and without the seed pass) == exhaustive sweep, bit for bit, and the
    work counters are unchanged; the pruned modes execute fewer evaluations."""
    pts = _cloud(kind, n, d, seed=n)
    g = torch.Generator().manual_seed(3)
    lms = pts[torch.randperm(n, generator=g)[:60]]
    from oracle.simplex_tree import delaunay_top_simplices

    cells = delaunay_top_simplices(lms.numpy())[:120]
    verts = lms[torch.as_tensor(cells)].cuda().contiguous()
    w = torch.as_tensor(flood_oracle.generate_grid(ppe, d)[0]).cuda()
    P = pts.cuda()
    ws = ext.cloud_build(P, 0)
    c, r = ext.bounding_balls(verts)
    results = {}
    for name, opts in {"exhaustive": {"prune": 0}, "pruned": {"prune": 1, "seed_stride": 16},
                       "pruned_noseed": {"prune": 1, "seed_stride": 1},
                       "pruned_one_level": {"prune": 1, "level2": 0, "slab_cull": 0}}.items():
        prev = {k: ext.set_option(k, v) for k, v in opts.items()}
        try:
            md2, cnt, ev, executed = ext.covering_radius(ws, n, d, verts, w, None, c, r)
            torch.cuda.synchronize()
            results[name] = (md2.cpu().numpy(), cnt.cpu().numpy(), int(ev.item()), int(executed.item()))
        finally:
            for k, v in prev.items():
                ext.set_option(k, v)
    ref = results["exhaustive"]
    assert ref[3] >= ref[2]                                   # exhaustive executes every evaluation (+ padding)
    for name in ("pruned", "pruned_noseed", "pruned_one_level"):
        got = results[name]
        np.testing.assert_array_equal(got[0], ref[0])
        np.testing.assert_array_equal(got[1], ref[1])
        assert got[2] == ref[2]
        assert got[3] < ref[3]


def test_covering_explicit_samples_and_random_weights(ext):
    seed_all(3)
    pts = _cloud("gauss", 8000, 3, seed=21)
    lms = pts[:40]
    from oracle.simplex_tree import delaunay_top_simplices

    cells = delaunay_top_simplices(lms.numpy())[:30]
    verts = lms[torch.as_tensor(cells)]
    w = torch.as_tensor(flood_oracle.generate_uniform_weights(300, 3))
    samples = torch.as_tensor(native.sample_points(w.numpy(), verts.numpy()))
    a = _covering_case(ext, pts, verts, w)
    b = _covering_case(ext, pts, verts, w, samples=samples)
    np.testing.assert_array_equal(a[0], b[0])
    _check_against_bruteforce(pts, verts, w, *a)


def test_covering_empty_ball_and_outside(ext):
    """A simplex far away from the cloud has no candidates: +inf, count 0 (the reference's
    Triton path has the same +inf initial value, triton_kernels.py:70)."""
    pts = _cloud("uniform", 2000, 3, seed=1)
    verts = torch.tensor([[[10.0, 10, 10], [10.1, 10, 10], [10, 10.1, 10], [10, 10, 10.1]],
                          [[0.5, 0.5, 0.5], [0.6, 0.5, 0.5], [0.5, 0.6, 0.5], [0.5, 0.5, 0.6]]])
    w = torch.as_tensor(flood_oracle.generate_grid(5, 3)[0])
    md2, cnt, ev, c, r = _covering_case(ext, pts, verts, w)
    assert np.isinf(md2[0]).all() and cnt[0] == 0
    assert np.isfinite(md2[1]).all() and cnt[1] > 0
    _check_against_bruteforce(pts, verts, w, md2, cnt, ev, c, r)


def test_face_max(ext):
    rng = np.random.default_rng(0)
    w = flood_oracle.generate_grid(7, 3)
    weights, vertex_idxs, face_idxs = w
    R = weights.shape[0]
    md2 = rng.random((50, R)).astype(np.float32)
    from flooder_b200.core import _support_masks

    sup = _support_masks(torch.as_tensor(weights)).cuda()
    out = ext.face_max(torch.as_tensor(md2).cuda(), sup, 4).cpu().numpy()
    # reference formulation: distances[:, face_idx].amax(dim=2)  (core.py:251-257)
    for rows, vsel in zip(face_idxs, vertex_idxs):
        for j in range(rows.shape[0]):
            mask = sum(1 << int(k) for k in vsel[j])
            want = np.sqrt(md2[:, rows[j]].max(axis=1))
            np.testing.assert_array_equal(out[:, mask - 1], want)
    flat = ext.face_max(torch.as_tensor(md2).cuda(), None, 4).cpu().numpy()
    np.testing.assert_array_equal(flat[:, 0], np.sqrt(md2.max(axis=1)))


def test_c_abi_direct_ctypes():
    """Drive the C ABI without the torch extension: raw device pointers + stream handle."""
    from flooder_b200 import _native

    lib = _native.cdll()
    pts = _cloud("gauss", 6000, 3, seed=8).cuda().contiguous()
    n, d, n_lms = 6000, 3, 50
    out = torch.empty(n_lms, dtype=torch.int64, device="cuda")
    wsb = lib.flood_fps_workspace_bytes(n, d, n_lms)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.flood_fps_f32(pts.data_ptr(), n, d, n_lms, 0, out.data_ptr(), ws.data_ptr(), wsb, stream)
    assert rc == 0, lib.flood_last_error()
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy(), native.fps(pts.cpu().numpy(), n_lms, 0))
    # error path: workspace too small -> negative code + message, nothing launched
    rc = lib.flood_fps_f32(pts.data_ptr(), n, d, n_lms, 0, out.data_ptr(), ws.data_ptr(), 8, stream)
    assert rc == -2 and b"workspace" in lib.flood_last_error()
    sm, khz = ctypes.c_int(), ctypes.c_int()
    assert lib.flood_device_info(ctypes.byref(sm), ctypes.byref(khz)) == 0 and sm.value > 0
