"""GPU parity of the public API against the CPU oracle and the golden fixtures.

Tolerance: the oracle measures exact nearest-neighbour distances in float64 on the same
float32 sample points (the CUDA kernel builds them with the same FMA chain); the kernel's
direct-difference float32 distance then carries a few ulps -> rtol 1e-5 as north_star states,
plus atol 1e-7 for values that are exactly 0 in one arithmetic and ~1e-8 in the other.
"""
import numpy as np
import pytest
import torch

import flooder_b200 as fb
from oracle import flood_oracle
from tests.helpers import (REF_CASES, assert_close_dict, golden_dict, golden_kwargs, load_golden,
                           seed_all)

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-7
DEV = torch.device("cuda")


@pytest.mark.parametrize("case", REF_CASES)
def test_golden_reference_runs(case):
    """flood_complex == what the reference itself returned for these inputs (tests/golden)."""
    g = load_golden("ref_" + case)
    pts = torch.as_tensor(g["points"]).to(DEV)
    lms = torch.as_tensor(g["landmarks"]).to(DEV)
    seed_all()
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        got = fb.flood_complex(pts, lms, **golden_kwargs(g))
    if case.endswith("f64"):
        # float64 inputs are evaluated in float64 (csrc/f64.cu), like the reference does
        assert_close_dict(got, golden_dict(g), rtol=1e-10, atol=1e-12, what=case)
    else:
        assert_close_dict(got, golden_dict(g), rtol=RTOL, atol=ATOL, what=case)


def test_shipped_animation_csv():
    g = load_golden("shipped_animation")
    got = fb.flood_complex(torch.as_tensor(g["points"]).to(DEV), torch.as_tensor(g["landmarks"]).to(DEV),
                           points_per_edge=int(g["points_per_edge"]))
    for a, b, f in g["edges"]:
        assert abs(got[(int(a), int(b))] - f) <= RTOL * f + 2e-8
    for a, b, c, f in g["triangles"]:
        assert abs(got[(int(a), int(b), int(c))] - f) <= RTOL * f + 2e-8


@pytest.mark.parametrize("num_witnesses", [1000, 10_000])
@pytest.mark.parametrize("num_landmarks", [20, 701, 2000])
@pytest.mark.parametrize("use_rand", [True, False])
def test_kdtree_vs_gpu(num_witnesses, num_landmarks, use_rand):
    """Mirror of the reference's test_kdtree_vs_triton (tests/test_flooder.py:119-157) with the
    tighter tolerance; includes n_landmarks > n_points (clamping)."""
    kwargs = {"num_rand": 512, "points_per_edge": None} if use_rand else {"num_rand": None, "points_per_edge": 20}
    seed_all()
    X = fb.generate_noisy_torus_points_3d(num_witnesses).to(DEV)
    L = fb.generate_landmarks(X, num_landmarks, start_idx=0)
    assert L.shape == (min(num_landmarks, num_witnesses), 3) and L.device == X.device
    seed_all()
    got = fb.flood_complex(X, L, **kwargs)
    seed_all()
    want = flood_oracle.flood_complex(X.cpu().numpy(), L.cpu().numpy(), **kwargs)
    assert_close_dict(got, want, rtol=RTOL, atol=ATOL)


def test_landmarks_from_int_and_fps_parity():
    seed_all()
    X = fb.generate_noisy_torus_points_3d(20_000).to(DEV)
    got = fb.flood_complex(X, 150, points_per_edge=10)
    lms = flood_oracle.generate_landmarks(X.cpu().numpy(), 150, 0)
    np.testing.assert_array_equal(fb.generate_landmarks(X, 150, start_idx=0).cpu().numpy(), lms)
    want = flood_oracle.flood_complex(X.cpu().numpy(), lms, points_per_edge=10)
    assert_close_dict(got, want, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("num_witnesses", [1000, 10_000])
@pytest.mark.parametrize("num_landmarks", [20, 1000])
@pytest.mark.parametrize("return_simplex_tree", [True, False])
def test_filtration_condition(num_witnesses, num_landmarks, return_simplex_tree):
    """Reference tests/test_flooder.py:160-211: the result is a filtered complex."""
    seed_all()
    X = fb.generate_noisy_torus_points_3d(num_witnesses).to(DEV)
    L = fb.generate_landmarks(X, num_landmarks)
    if return_simplex_tree:
        st = fb.flood_complex(X, L, return_simplex_tree=True)
    else:
        fc = fb.flood_complex(X, L, return_simplex_tree=False)
        st = fb.SimplexTree()
        for simplex in fc:
            st.insert(simplex, float("inf"))
            st.assign_filtration(simplex, fc[simplex])
    for simplex, filtration in st.get_simplices():
        faces = list(st.get_boundaries(simplex))
        assert len(faces) == (len(simplex) if len(simplex) > 1 else 0)
        for face, face_filtration in faces:
            assert face_filtration <= filtration


@pytest.mark.parametrize("pointcloud", ["torus", "cheese"])
def test_float64_inputs(pointcloud):
    """Reference tests/test_flooder.py:214-246 (f32 vs f64 within 3e-6)."""
    seed_all()
    pts = fb.generate_noisy_torus_points_3d(50_000) if pointcloud == "torus" else fb.generate_swiss_cheese_points(50_000)[0]
    seed_all()
    lms = fb.generate_landmarks(pts.to(DEV), 500)
    flood32 = fb.flood_complex(pts.to(DEV), lms)
    with pytest.warns(RuntimeWarning):
        flood64 = fb.flood_complex(pts.to(DEV, torch.float64), lms.to(torch.float64))
    for s in flood32:
        assert abs(flood32[s] - flood64[s]) < 3e-6


def test_vs_exact_alpha_free_property_2d():
    """With L = X every landmark is a cloud point: all vertex values are exactly 0 and every
    edge value is at most half its length (the midpoint is within len/2 of both endpoints)."""
    seed_all()
    X = fb.generate_figure_eight_points_2d(1000).to(DEV)
    fc = fb.flood_complex(X, X, points_per_edge=31, batch_size=8)
    Xc = X.cpu().numpy().astype(np.float64)
    for s, f in fc.items():
        if len(s) == 1:
            assert f == 0.0
        elif len(s) == 2:
            assert f <= 0.5 * np.linalg.norm(Xc[s[0]] - Xc[s[1]]) * (1 + 1e-5) + 1e-7


def test_errors_match_reference():
    X = torch.rand(100, 3, device=DEV)
    with pytest.raises(RuntimeError, match="must be positive"):
        fb.generate_landmarks(X, 0)
    with pytest.raises(RuntimeError, match="landmarks.device"):
        fb.flood_complex(X, X[:10].cpu())
    with pytest.raises(RuntimeError, match="landmarks.dtype"):
        fb.flood_complex(X, X[:10].double())
    with pytest.raises(TypeError, match="not supported"):
        fb.flood_complex(X.half(), X[:10].half())
    with pytest.raises(RuntimeError, match="Device not supported"):
        fb.flood_complex(X.cpu(), X[:10].cpu())


def _diagrams(tree_or_dict, dims):
    if isinstance(tree_or_dict, dict):
        st = fb.SimplexTree()
        for s, f in tree_or_dict.items():
            st.insert(s, f)
    else:
        st = tree_or_dict
    st.compute_persistence()
    return [st.persistence_intervals_in_dimension(d) for d in dims]


@pytest.mark.parametrize("use_rand", [True, False])
@pytest.mark.parametrize("batch_size", [8, 23])
def test_vs_alpha(use_rand, batch_size):
    """Reference tests/test_flooder.py:24-75: with L = X the Flood complex has the persistence of
    the Alpha complex (bottleneck distance < 5e-4 in dimensions 0 and 1).  Alpha filtration and
    bottleneck distance come from the CPU oracle (gudhi is not installed)."""
    from oracle import alpha

    seed_all()
    X = fb.generate_figure_eight_points_2d(1000).to(DEV)
    kwargs = {"num_rand": 20_000, "points_per_edge": None} if use_rand else {"num_rand": None, "points_per_edge": 130}
    stree = fb.flood_complex(X, X, return_simplex_tree=True, batch_size=batch_size, **kwargs)
    flood = _diagrams(stree, range(2))
    alpha_tree = fb.SimplexTree()
    for s, f in alpha.alpha_filtration(X.cpu().numpy()).items():
        alpha_tree.insert(s, f)
    ref = _diagrams(alpha_tree, range(2))
    for dim in range(2):
        dist = alpha.bottleneck_distance(flood[dim], ref[dim])
        assert dist < 5e-4, f"bottleneck distance {dist} in dimension {dim} (use_rand={use_rand})"


def test_persistence_diagrams_match_oracle():
    """north_star: persistence diagrams (dims 0-2) of the GPU filtration and of the reference CPU
    filtration agree within 1e-5 relative."""
    seed_all()
    X = fb.generate_noisy_torus_points_3d(10_000).to(DEV)
    L = fb.generate_landmarks(X, 150, start_idx=0)
    got = _diagrams(fb.flood_complex(X, L, points_per_edge=15), range(3))
    want = _diagrams(flood_oracle.flood_complex(X.cpu().numpy(), L.cpu().numpy(), points_per_edge=15), range(3))
    for dim in range(3):
        assert got[dim].shape == want[dim].shape, f"dimension {dim}: different number of intervals"
        fin = np.isfinite(want[dim])
        np.testing.assert_allclose(got[dim][fin], want[dim][fin], rtol=1e-5, atol=1e-7)
        assert np.array_equal(np.isinf(got[dim]), np.isinf(want[dim]))


def test_landmarks_off_the_cloud_and_small_inputs():
    """Landmarks need not be cloud points (vertex values become > 0); tiny clouds and fewer
    landmarks than a full-dimensional cell needs still give the oracle's answer."""
    seed_all(5)
    X = torch.rand(3000, 3)
    L = torch.rand(30, 3) * 1.2 - 0.1                      # partly outside the cloud's box
    got = fb.flood_complex(X.to(DEV), L.to(DEV), points_per_edge=8)
    want = flood_oracle.flood_complex(X.numpy(), L.numpy(), points_per_edge=8)
    assert_close_dict(got, want, rtol=RTOL, atol=ATOL)
    assert min(v for s, v in got.items() if len(s) == 1) > 0
    # three landmarks in 3-D: one triangle, no tetrahedron
    got = fb.flood_complex(X.to(DEV), L[:3].to(DEV), points_per_edge=8)
    want = flood_oracle.flood_complex(X.numpy(), L[:3].numpy(), points_per_edge=8, max_dimension=2)
    assert set(got) == {(0,), (1,), (2,), (0, 1), (0, 2), (1, 2), (0, 1, 2)}
    assert_close_dict(got, want, rtol=RTOL, atol=ATOL)
    # a five-point cloud
    tiny = X[:5].to(DEV)
    got = fb.flood_complex(tiny, tiny, points_per_edge=5)
    want = flood_oracle.flood_complex(X[:5].numpy(), X[:5].numpy(), points_per_edge=5)
    assert_close_dict(got, want, rtol=RTOL, atol=ATOL)


def test_max_dimension_one_in_2d_and_random_mode_dims():
    seed_all(6)
    X = fb.generate_figure_eight_points_2d(4000)
    L = fb.generate_landmarks(X.to(DEV), 50, start_idx=0).cpu()
    for kwargs in ({"max_dimension": 1, "points_per_edge": 12},
                   {"max_dimension": 1, "points_per_edge": None, "num_rand": 77}):
        seed_all(7)
        got = fb.flood_complex(X.to(DEV), L.to(DEV), **kwargs)
        seed_all(7)
        want = flood_oracle.flood_complex(X.numpy(), L.numpy(), **kwargs)
        assert_close_dict(got, want, rtol=RTOL, atol=ATOL, what=str(kwargs))


# ------------------------------------------------------------------------------------------
# round 2
# ------------------------------------------------------------------------------------------
def test_one_dimensional_cloud():
    """The reference is dimension-generic (flooder/core.py:146-188); 1-D clouds go through the same
    kernels (records padded with a zero coordinate) and give the oracle's values."""
    seed_all(8)
    X = torch.rand(5000, 1)
    L = fb.generate_landmarks(X.to(DEV), 40, start_idx=0)
    np.testing.assert_array_equal(L.cpu().numpy(), flood_oracle.generate_landmarks(X.numpy(), 40, 0))
    for kwargs in ({"points_per_edge": 30}, {"points_per_edge": None, "num_rand": 100}):
        seed_all(9)
        got = fb.flood_complex(X.to(DEV), L, **kwargs)
        seed_all(9)
        want = flood_oracle.flood_complex(X.numpy(), L.cpu().numpy(), **kwargs)
        assert len(got) == 40 + 39
        assert_close_dict(got, want, rtol=RTOL, atol=ATOL, what=str(kwargs))


def test_gudhi_branch_end_to_end(monkeypatch):
    """flood_complex with a gudhi module present (the oracle's stand-in injected, gudhi itself is
    not in the image): Delaunay step and container are "gudhi's", values equal the default path
    (reference flooder/core.py:130-132, 278-288)."""
    import types

    from flooder_b200 import simplex_tree as st
    from oracle import simplex_tree as ost

    seed_all(10)
    X = fb.generate_noisy_torus_points_3d(8000).to(DEV)
    L = fb.generate_landmarks(X, 60, start_idx=0)
    plain = {kw: fb.flood_complex(X, L, **dict(kw)) for kw in
             ((("points_per_edge", 9),), (("points_per_edge", None), ("num_rand", 50), ("max_dimension", 2)))}
    g = types.ModuleType("gudhi")
    g.DelaunayComplex, g.SimplexTree = ost.DelaunayComplex, ost.DictSimplexTree
    monkeypatch.setattr(st, "_gudhi", g)
    monkeypatch.setattr(st, "HAS_GUDHI", True)
    for kw, want in plain.items():
        torch.manual_seed(11)
        tree = fb.flood_complex(X, L, return_simplex_tree=True, **dict(kw))
        assert isinstance(tree, ost.DictSimplexTree)
        torch.manual_seed(11)
        got = fb.flood_complex(X, L, **dict(kw))
        assert set(got) == set(want)
        if dict(kw).get("num_rand") is None:                 # random mode draws fresh weights per call
            for s, v in want.items():
                assert got[s] == v or (np.isnan(v) and np.isnan(got[s])), s
            assert dict((tuple(s), f) for s, f in tree.get_simplices()) == got


def test_memory_bounded_slabs(monkeypatch):
    """The (S, R) per-sample buffer is processed in slabs when it would not fit (the role of the
    reference's batch_size, flooder/core.py:193-226): same bits."""
    seed_all(12)
    X = fb.generate_noisy_torus_points_3d(20_000).to(DEV)
    L = fb.generate_landmarks(X, 80, start_idx=0)
    want = fb.flood_complex(X, L, points_per_edge=10)
    monkeypatch.setenv("FLOODER_B200_SLAB_BYTES", str(37 * 220 * 4))      # 37 simplices per slab
    got = fb.flood_complex(X, L, points_per_edge=10)
    assert got == want


@pytest.mark.parametrize("dim,kwargs", [(2, {"points_per_edge": 8}), (3, {"points_per_edge": 12}),
                                        (3, {"points_per_edge": None, "num_rand": 200}), (5, {"points_per_edge": 3}),
                                        (1, {"points_per_edge": 10})])
def test_float64_kernels_match_oracle(dim, kwargs):
    """float64 inputs run float64 kernels (reference: flooder/triton_kernels.py:226-229): values
    agree with the float64 CPU path of the oracle to float64 accuracy, far inside the 3e-6 the
    reference's own test_float64 allows between precisions.  Landmarks are triangulated in float64."""
    seed_all(13)
    X = torch.rand(6000, dim, dtype=torch.float64)
    L = X[:60].clone()
    seed_all(14)
    with pytest.warns(RuntimeWarning):
        got = fb.flood_complex(X.to(DEV), L.to(DEV), **kwargs)
    seed_all(14)
    want = flood_oracle.flood_complex(X.numpy(), L.numpy(), **kwargs)
    assert_close_dict(got, want, rtol=1e-10, atol=1e-12, what=f"{dim}-D {kwargs}")
    # and the float32 path stays within the reference's own bound of it
    seed_all(14)
    got32 = fb.flood_complex(X.to(DEV, torch.float32), L.to(DEV, torch.float32), **kwargs)
    if kwargs.get("num_rand") is None:
        for s, v in got.items():
            assert abs(got32[s] - v) < 3e-6
