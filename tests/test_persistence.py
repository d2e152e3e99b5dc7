"""Persistence of the stand-in SimplexTree (CPU): known complexes, and the Alpha-complex oracle."""
import itertools

import numpy as np
import pytest

import flooder_b200 as fb
from flooder_b200.simplex_tree import SimplexTree
from oracle import alpha
from tests.helpers import seed_all


def test_hollow_triangle_and_filled():
    st = SimplexTree()
    for v in range(3):
        st.insert([v], 0.0)
    st.insert([0, 1], 1.0)
    st.insert([1, 2], 2.0)
    st.insert([0, 2], 3.0)
    st.compute_persistence(persistence_dim_max=True)
    h0 = st.persistence_intervals_in_dimension(0)
    h1 = st.persistence_intervals_in_dimension(1)
    np.testing.assert_array_equal(h0, [[0, 1], [0, 2], [0, np.inf]])
    np.testing.assert_array_equal(h1, [[3, np.inf]])
    assert st.betti_numbers() == [1, 1]
    st.insert([0, 1, 2], 5.0)
    st.compute_persistence(persistence_dim_max=True)
    np.testing.assert_array_equal(st.persistence_intervals_in_dimension(1), [[3, 5]])
    st.compute_persistence()                       # gudhi default: top dimension skipped
    assert len(st.persistence_intervals_in_dimension(2)) == 0


def test_sphere_boundary_of_tetrahedron():
    st = SimplexTree()
    for tri in itertools.combinations(range(4), 3):
        st.insert(tri, 1.0)
    st.compute_persistence(persistence_dim_max=True)
    assert st.betti_numbers() == [1, 0, 1]
    assert [d for d, _ in st.persistence(persistence_dim_max=True)] == [2, 0]   # zero-length pairs dropped


def test_unassigned_values_are_rejected():
    st = SimplexTree.from_cells(np.array([[0, 1, 2]]))
    with pytest.raises(ValueError):
        st.compute_persistence()


def test_alpha_oracle_annulus():
    """Alpha persistence of an annulus: one connected component, one dominant loop whose death is
    about the inner radius, everything else short-lived."""
    seed_all(1)
    pts = fb.generate_annulus_points_2d(400, radius=1.0, width=0.2).numpy().astype(np.float64)
    filt = alpha.alpha_filtration(pts)
    st = SimplexTree()
    for s, f in filt.items():
        st.insert(s, f)
    assert not st.make_filtration_non_decreasing()       # alpha values are already monotone
    st.compute_persistence()
    h0, h1 = st.persistence_intervals_in_dimension(0), st.persistence_intervals_in_dimension(1)
    assert np.isinf(h0[:, 1]).sum() == 1
    pers = np.sort(h1[:, 1] - h1[:, 0])[::-1]
    assert pers[0] > 0.5 and pers[1] < 0.15
    big = h1[np.argmax(h1[:, 1] - h1[:, 0])]
    assert 0.7 < big[1] < 0.85


def test_bottleneck_distance():
    a = np.array([[0.0, 1.0], [0.0, 0.1], [0.0, np.inf]])
    b = np.array([[0.0, 1.2], [0.05, np.inf]])
    assert alpha.bottleneck_distance(a, a) == 0.0
    assert abs(alpha.bottleneck_distance(a, b) - 0.2) < 1e-12       # (0,1)~(0,1.2); (0,.1)->diag .05
    assert alpha.bottleneck_distance(a, np.array([[0.0, 1.0]])) == float("inf")   # essential count
    c = np.array([[0.0, 1.0], [0.0, 0.6]])
    d = np.array([[0.0, 1.05]])
    assert abs(alpha.bottleneck_distance(c, d) - 0.3) < 1e-12       # (0,.6) goes to the diagonal
