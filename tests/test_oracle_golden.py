"""Pin the CPU oracle (``oracle/``) before anything is checked against it.

* the CSVs shipped by the reference authors (inputs + filtration values + FPS order),
* outputs of the reference's own code executed in the authoring container
  (``tests/golden/make_golden.py``),
* the gudhi Delaunay simplex sets shipped with the reference's visualisations.
"""
import numpy as np
import pytest

from oracle import flood_oracle, native
from oracle.simplex_tree import DictSimplexTree, delaunay_top_simplices
from tests.helpers import (REF_CASES, golden_dict, golden_kwargs, load_golden, seed_all,
                           assert_close_dict)


def test_shipped_animation_filtration_values():
    """docs/animation/{edges,triangles}.csv, written with %.8f by the reference authors."""
    g = load_golden("shipped_animation")
    res = flood_oracle.flood_complex(g["points"], g["landmarks"], points_per_edge=int(g["points_per_edge"]))
    edges, tris = g["edges"], g["triangles"]
    assert {k for k in res if len(k) == 2} == {(int(a), int(b)) for a, b, _ in edges}
    assert {k for k in res if len(k) == 3} == {(int(a), int(b), int(c)) for a, b, c, _ in tris}
    for a, b, f in edges:
        assert abs(res[(int(a), int(b))] - f) < 2e-8
    for a, b, c, f in tris:
        assert abs(res[(int(a), int(b), int(c))] - f) < 2e-8
    for k in res:
        if len(k) == 1:
            assert res[k] == 0.0  # landmarks are cloud points


def test_shipped_animation_fps_order():
    """docs/animation/landmarks.csv are 25 of the 200 points in fpsample's order."""
    g = load_golden("shipped_animation")
    pts, lms = g["points"], g["landmarks"]
    start = int(np.nonzero((np.abs(pts - lms[0]) < 1e-7).all(axis=1))[0][0])
    for fps in (native.fps, flood_oracle.fps_exact):
        idx = fps(pts, len(lms), start)
        assert len(set(idx.tolist())) == len(lms)
        np.testing.assert_allclose(pts[idx], lms, atol=1e-7, rtol=0)


@pytest.mark.parametrize("name", ["virus", "coral", "lockwasher"])
def test_shipped_delaunay_simplex_sets(name):
    """Qhull == gudhi/CGAL Delaunay on the 1000-landmark fixtures the reference ships."""
    g = load_golden("shipped_" + name)
    tops = delaunay_top_simplices(g["landmarks"])
    assert {tuple(r) for r in tops.tolist()} == {tuple(r) for r in g["tetrahedra"].tolist()}
    st = DictSimplexTree.from_top_simplices(tops)
    keys = {tuple(s) for s, _ in st.get_simplices()}
    assert {k for k in keys if len(k) == 3} == {tuple(r) for r in g["triangles"].tolist()}
    assert {k for k in keys if len(k) == 2} == {tuple(r) for r in g["edges"].tolist()}
    # shipped values are monotone; the tree fix-up must leave them alone
    for row, f in zip(g["tetrahedra"].tolist(), g["tetrahedra_f"].tolist()):
        st.assign_filtration(row, f)
    for row, f in zip(g["triangles"].tolist(), g["triangles_f"].tolist()):
        st.assign_filtration(row, f)
    for row, f in zip(g["edges"].tolist(), g["edges_f"].tolist()):
        st.assign_filtration(row, f)
    for v in range(len(g["landmarks"])):
        st.assign_filtration([v], 0.0)
    assert st.make_filtration_non_decreasing() is False


@pytest.mark.parametrize("case", REF_CASES)
def test_reference_runs(case):
    """Oracle == the reference's own flood_complex (CPU path) on seeded inputs."""
    g = load_golden("ref_" + case)
    seed_all()
    got = flood_oracle.flood_complex(g["points"], g["landmarks"], **golden_kwargs(g))
    # identical algorithm and arithmetic: only libm / BLAS build differences may show.  The float64
    # fixture pins the reference's weight arithmetic (torch.divide rounds the quotients to float32
    # before widening them, flooder/core.py:400-401): float64 accuracy, not 1e-6.
    if case.endswith("f64"):
        assert_close_dict(got, golden_dict(g), rtol=1e-11, atol=1e-13, what=case)
    else:
        assert_close_dict(got, golden_dict(g), rtol=1e-6, atol=1e-7, what=case)


@pytest.mark.parametrize("case", ["torus3d_grid", "fig8_2d_grid", "uniform4d_grid"])
def test_reference_fps(case):
    """The fixtures' landmarks were produced by FPS from index 0."""
    g = load_golden("ref_" + case)
    idx = native.fps(g["points"], len(g["landmarks"]), 0)
    np.testing.assert_array_equal(g["points"][idx], g["landmarks"])
    np.testing.assert_array_equal(flood_oracle.fps_exact(g["points"], len(g["landmarks"]), 0), idx)


def test_grid_shapes():
    """Shapes quoted in SURVEY.md section 8(a) for generate_grid (core.py:346-402)."""
    w, vidx, fidx = flood_oracle.generate_grid(30, 3)
    assert w.shape == (4960, 4) and w.dtype == np.float32
    assert [tuple(f.shape) for f in fidx] == [(1, 4960), (4, 465), (6, 30), (4, 1)]
    assert [tuple(v.shape) for v in vidx] == [(1, 4), (4, 3), (6, 2), (4, 1)]
    np.testing.assert_array_equal(w[0], [0, 0, 0, 1])
    np.testing.assert_allclose(w.sum(axis=1), 1.0, atol=1e-6)


def test_native_matches_numpy():
    rng = np.random.default_rng(0)
    pts = rng.normal(size=(3000, 3)).astype(np.float32)
    verts = pts[rng.integers(0, 3000, size=(20, 4))]
    c, r = flood_oracle.bounding_balls(verts, 3)
    w, _, _ = flood_oracle.generate_grid(6, 3)
    x_np = flood_oracle.sample_points(w, verts)
    np.testing.assert_array_equal(native.sample_points(w, verts), x_np)
    cnt = native.ball_counts(pts, c, r)
    d2 = ((pts[None].astype(np.float64) - c[:, None]) ** 2).sum(-1)
    approx = (d2 <= (r.astype(np.float64) ** 2)[:, None]).sum(1)
    assert np.abs(cnt - approx).max() <= 2
    from scipy.spatial import KDTree

    ball = native.min_dist(pts, x_np, c, r)
    exact, _ = KDTree(pts).query(x_np.reshape(-1, 3))
    np.testing.assert_allclose(ball.reshape(-1), exact, rtol=2e-6, atol=1e-7)
