"""CPU-side checks: the C-ABI library loads and exports every symbol include/flood_b200.h
declares (no compute without a GPU), the host logic of the package, and the no-fallback rule."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import flooder_b200 as fb
from flooder_b200 import _native, core
from flooder_b200.simplex_tree import FaceTable, SimplexTree, delaunay_cells, faces_of_cells
from oracle import flood_oracle
from oracle.simplex_tree import DictSimplexTree, delaunay_top_simplices
from tests.helpers import load_golden, seed_all

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "flood_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(flood_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    declared = _declared_symbols()
    assert sorted(_native.EXPORTED_SYMBOLS) == declared
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in flood_b200.h but not exported"
    assert _native.cdll().flood_abi_version() == 1


def test_abi_argument_validation_without_gpu():
    """Argument checks run before any CUDA call, so they are testable on a CPU box."""
    lib = _native.cdll()
    assert lib.flood_cloud_workspace_bytes(0, 3) == 0
    assert lib.flood_cloud_workspace_bytes(1000, 9) == 0
    assert lib.flood_cloud_workspace_bytes(1_000_000, 3) > 16_000_000
    assert lib.flood_fps_workspace_bytes(1000, 3, 10) >= 1000 * 4
    rc = lib.flood_fps_f32(None, 10, 3, 5, 0, None, None, 0, None)
    assert rc == -1 and b"bad arguments" in lib.flood_last_error()
    rc = lib.flood_covering_radius_f32(None, 10, 3, None, 5, 4, None, 10, None, None, None, None, None, None, None, 0, None)
    assert rc == -1
    rc = lib.flood_bounding_balls_f32(None, 5, 4, 3, None, None, None)
    assert rc == -1


def test_torch_extension_loads():
    ext = _native.ext()
    assert ext.abi_version() == 1
    for fn in ("fps", "cloud_build", "bounding_balls", "covering_radius", "face_max", "kernel_ms"):
        assert hasattr(ext, fn)


def test_no_cpu_fallback():
    """The product path refuses CPU tensors instead of silently computing elsewhere."""
    x = torch.rand(50, 3)
    with pytest.raises(RuntimeError, match="Device not supported|CUDA"):
        fb.flood_complex(x, x[:8])
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        _native.ext().fps(x, 4, 0)
    src = open(os.path.join(ROOT, "flooder_b200", "core.py")).read()
    assert "oracle" not in src and "KDTree" not in src


def test_product_does_not_import_oracle():
    for fn in os.listdir(os.path.join(ROOT, "flooder_b200")):
        if fn.endswith(".py"):
            text = open(os.path.join(ROOT, "flooder_b200", fn)).read()
            assert "import oracle" not in text and "from oracle" not in text, fn


def test_generate_grid_matches_oracle():
    for n, dim in [(30, 3), (130, 2), (20, 3), (6, 5), (4, 6), (5, 1)]:
        w, vidx, fidx = fb.generate_grid(n, dim, "cpu")
        wo, vo, fo = flood_oracle.generate_grid(n, dim)
        np.testing.assert_array_equal(w.numpy(), wo)
        for a, b in zip(fidx, fo):
            np.testing.assert_array_equal(a.numpy(), b)
        for a, b in zip(vidx, vo):
            np.testing.assert_array_equal(a.numpy(), b)


def test_uniform_weights_match_oracle():
    seed_all(7)
    a = fb.generate_uniform_weights(100, 3, "cpu").numpy()
    seed_all(7)
    b = flood_oracle.generate_uniform_weights(100, 3)
    np.testing.assert_array_equal(a, b)
    assert (fb.generate_uniform_weights(5, 0, "cpu").numpy() == 1).all()


def test_support_masks_select_reference_faces():
    """Samples with support inside face m == the reference's face_idxs rows for that face."""
    w, vidx, fidx = fb.generate_grid(9, 3, "cpu")
    sup = core._support_masks(w).numpy()
    for rows, vsel in zip(fidx, vidx):
        for j in range(rows.shape[0]):
            m = sum(1 << int(k) for k in vsel[j])
            mine = np.nonzero((sup & ~m) == 0)[0]
            np.testing.assert_array_equal(mine, rows[j].numpy())


def test_scatter_face_values_min_over_cofaces():
    cells = np.array([[0, 1, 2], [1, 2, 3]])
    vals = np.arange(14, dtype=np.float32).reshape(2, 7)
    table = FaceTable(cells, n_vertices=4)
    values = table.nan_values()
    core._scatter_face_values(table, vals, values)
    out = {tuple(f): v for k in table.faces for f, v in zip(table.faces[k].tolist(), values[k].tolist())}
    assert out[(0, 1, 2)] == 6 and out[(1, 2, 3)] == 13
    assert out[(1, 2)] == min(vals[0, 0b110 - 1], vals[1, 0b011 - 1])
    assert out[(0,)] == vals[0, 0] and out[(3,)] == vals[1, 0b100 - 1]
    assert len(out) == 4 + 5 + 2


@pytest.mark.parametrize("n_vertices", [40, 3_000_000_000])
def test_face_table_matches_dict_tree(n_vertices):
    """Array face table (packed int64 keys, and the unpackable fallback) == the dict tree,
    including the monotone fix-up with unassigned (NaN) simplices."""
    rng = np.random.default_rng(3)
    pts = rng.random((40, 4))
    cells = delaunay_cells(pts)
    table = FaceTable(cells, n_vertices=n_vertices)
    assert (table.keys[5] is None) == (n_vertices > 1e9)
    ref = DictSimplexTree.from_top_simplices(cells)
    keys = [tuple(s) for s, _ in ref.get_simplices()]
    mine = [tuple(f) for k in sorted(table.faces) for f in table.faces[k].tolist()]
    assert mine == keys
    for k in table.faces:                       # cell -> face ids are consistent
        cols = np.asarray(table.combos[k])
        np.testing.assert_array_equal(table.faces[k][table.cell_face[k]], cells[:, cols])
    values = table.nan_values()
    for k in (1, 2, 4):                          # triangles (3) and cells (5) stay unassigned
        values[k] = rng.random(len(table.faces[k]))
        for f, v in zip(table.faces[k].tolist(), values[k].tolist()):
            ref.assign_filtration(f, v)
    table.make_non_decreasing(values)
    ref.make_filtration_non_decreasing()
    st = SimplexTree.from_arrays(table.faces, values)
    for (s, f), (s2, f2) in zip(st.get_simplices(), ref.get_simplices()):
        assert s == s2 and f == f2
    # keys prepared ahead of the values (flood_complex builds them while the GPU works) give the
    # same tree, in the same insertion order
    from flooder_b200.simplex_tree import face_keys

    st2 = SimplexTree.from_arrays(table.faces, values, keys=face_keys(table.faces))
    assert list(st2.to_flat_dict().items()) == list(st.to_flat_dict().items())


@pytest.mark.parametrize("name", ["virus", "coral", "lockwasher"])
def test_delaunay_matches_shipped_gudhi_sets(name):
    g = load_golden("shipped_" + name)
    cells = delaunay_cells(g["landmarks"])
    assert {tuple(r) for r in cells.tolist()} == {tuple(r) for r in g["tetrahedra"].tolist()}
    assert {tuple(r) for r in faces_of_cells(cells, 3).tolist()} == {tuple(r) for r in g["triangles"].tolist()}
    assert {tuple(r) for r in faces_of_cells(cells, 2).tolist()} == {tuple(r) for r in g["edges"].tolist()}


def test_simplex_tree_matches_oracle_tree():
    rng = np.random.default_rng(1)
    pts = rng.random((60, 3))
    cells = delaunay_cells(pts)
    np.testing.assert_array_equal(np.unique(cells, axis=0), np.unique(delaunay_top_simplices(pts), axis=0))
    a, b = SimplexTree.from_cells(cells), DictSimplexTree.from_top_simplices(cells)
    assert a.num_simplices() == b.num_simplices() and a.num_vertices() == 60 and a.dimension() == 3
    keys = [tuple(s) for s, _ in a.get_simplices()]
    assert keys == [tuple(s) for s, _ in b.get_simplices()]
    vals = rng.random(len(keys))
    for k, v in zip(keys, vals):
        if len(k) != 3:            # leave the triangles unassigned (NaN)
            a.assign_filtration(k, v)
            b.assign_filtration(k, v)
    assert a.make_filtration_non_decreasing() == b.make_filtration_non_decreasing()
    for (s, f), (s2, f2) in zip(a.get_simplices(), b.get_simplices()):
        assert s == s2 and (f == f2 or (np.isnan(f) and np.isnan(f2)))
    for s, f in a.get_simplices():
        for face, ff in a.get_boundaries(s):
            assert not ff > f
    with pytest.raises(KeyError):
        a.assign_filtration((0, 59, 58, 57, 56), 1.0)


def test_synthetic_generators_match_reference_bytes():
    g = load_golden("ref_generators")
    seed_all()
    np.testing.assert_array_equal(fb.generate_noisy_torus_points_3d(64).numpy(), g["torus"])
    seed_all()
    np.testing.assert_array_equal(fb.generate_figure_eight_points_2d(64).numpy(), g["fig8"])
    seed_all()
    np.testing.assert_array_equal(fb.generate_swiss_cheese_points(64)[0].numpy(), g["cheese"])
    seed_all()
    np.testing.assert_array_equal(fb.generate_annulus_points_2d(64).numpy(), g["annulus"])


def test_landmark_argument_errors():
    with pytest.raises(RuntimeError, match="must be positive"):
        fb.generate_landmarks(torch.rand(10, 2), 0)
    with pytest.raises(RuntimeError, match="must be positive"):
        fb.generate_landmarks(torch.rand(10, 2), -3)


# ------------------------------------------------------------------------------------------
# round 2: 1-D / degenerate landmark sets, the gudhi branch, brick ordering
# ------------------------------------------------------------------------------------------
def test_delaunay_cells_1d_and_degenerate():
    x = np.array([[0.3], [0.1], [0.9], [0.5]])
    cells = delaunay_cells(x)
    assert sorted(map(tuple, cells.tolist())) == [(0, 1), (0, 3), (2, 3)]
    np.testing.assert_array_equal(np.sort(cells, axis=0), np.sort(delaunay_top_simplices(x), axis=0))
    # affinely dependent landmarks: a clear error instead of a raw QhullError
    flat = np.random.default_rng(0).random((20, 3))
    flat[:, 2] = 0.0
    with pytest.raises(RuntimeError, match="do not span"):
        delaunay_cells(flat)
    # duplicated landmarks are left out by Qhull: reported, not silent
    pts = np.random.default_rng(1).random((30, 2))
    dup = np.concatenate([pts, pts[:3]])
    with pytest.warns(RuntimeWarning, match="not vertices"):
        cells = delaunay_cells(dup)
    assert cells.max() < 33


def _fake_gudhi():
    import types

    from oracle import simplex_tree as ost

    g = types.ModuleType("gudhi")
    g.DelaunayComplex = ost.DelaunayComplex
    g.SimplexTree = ost.DictSimplexTree
    return g


def test_gudhi_branch_with_injected_module(monkeypatch):
    """The reference's container contract (flooder/core.py:130-132, 278-288): with gudhi importable
    the Delaunay step and the returned tree are gudhi's.  gudhi is not in the image, so the
    oracle's stand-in module is injected in its place."""
    from flooder_b200 import simplex_tree as st

    monkeypatch.setattr(st, "_gudhi", _fake_gudhi())
    monkeypatch.setattr(st, "HAS_GUDHI", True)
    rng = np.random.default_rng(3)
    lms = rng.random((40, 3))
    cells, tree = st.delaunay_complex(lms)
    assert tree is not None and isinstance(tree, DictSimplexTree)
    np.testing.assert_array_equal(np.unique(cells, axis=0), np.unique(delaunay_cells(lms), axis=0))
    table = FaceTable(cells, n_vertices=40)
    values = {k: rng.random(len(f)) for k, f in table.faces.items()}
    values[4][:] = np.nan                                     # "above max_dimension": stays unassigned
    want_vals = {k: v.copy() for k, v in values.items()}
    got_tree = core._write_back(table, {k: v.copy() for k, v in values.items()}, tree, True)
    assert got_tree is tree                                   # gudhi's own tree is returned
    got = core._write_back(table, {k: v.copy() for k, v in values.items()},
                           st.delaunay_complex(lms)[1], False)
    # same as the stand-in path
    table.make_non_decreasing(want_vals)
    want = SimplexTree.from_arrays(table.faces, want_vals).to_flat_dict()
    assert set(got) == set(want)
    for key, v in want.items():
        assert (np.isnan(v) and np.isnan(got[key])) or got[key] == v, key
    for simplex, f in tree.get_simplices():                   # filtered complex
        for _face, ff in tree.get_boundaries(simplex):
            assert not (ff > f)


def test_brick_order_is_a_compact_permutation():
    from flooder_b200.bricks import brick_order

    w, _, _ = flood_oracle.generate_grid(30, 3)
    groups = [8, 8, 8, 7] * 5                                 # flood_covering_bricks for R = 4960, narrow shape
    perm = brick_order(w, groups, 4)
    assert sorted(perm.tolist()) == list(range(len(w)))

    # axis-aligned boxes of the bricks on a (rotated) regular tetrahedron with unit edges
    verts = np.array([[1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]], float) / np.sqrt(8)
    q, _ = np.linalg.qr(np.random.default_rng(0).normal(size=(3, 3)))
    x = w.astype(np.float64) @ verts @ q.T

    def mean_box_volume(order):
        vol, pos = [], 0
        for g in groups:
            idx = order[pos:pos + g * 32]
            pos += g * 32
            vol.append(np.prod(x[idx].max(axis=0) - x[idx].min(axis=0) + 0.1))
        return np.mean(vol)

    assert mean_box_volume(perm) < 0.25 * mean_box_volume(np.arange(len(w)))

    # the kernel's second pruning level boxes the pairs of groups (0,1), (2,3), ... of a brick:
    # under the brick order a pair is a compact piece of its brick (mean pair box well below half
    # of the mean brick box: a brick of 8 groups holds 4 pairs)
    def mean_pair_volume(order):
        vol, pos = [], 0
        for g in groups:
            for j in range(0, g, 2):
                idx = order[pos + j * 32: pos + min(j + 2, g) * 32]
                vol.append(np.prod(x[idx].max(axis=0) - x[idx].min(axis=0)))
            pos += g * 32
        return np.mean(vol)

    def mean_brick_volume(order):
        vol, pos = [], 0
        for g in groups:
            idx = order[pos:pos + g * 32]
            pos += g * 32
            vol.append(np.prod(x[idx].max(axis=0) - x[idx].min(axis=0)))
        return np.mean(vol)

    assert mean_pair_volume(perm) < 0.4 * mean_brick_volume(perm)
    # ragged tail, tiny sets, random weights
    for R, gs, per in [(252, [2, 2, 2, 2], 4), (33, [1, 1], 2), (5000, [8] * 19 + [5], 4)]:
        ww = np.random.default_rng(R).dirichlet(np.ones(4), size=R).astype(np.float32)
        p = brick_order(ww, gs, per)
        assert sorted(p.tolist()) == list(range(R))
    with pytest.raises(ValueError):
        brick_order(w, [8] * 3, 4)


def test_covering_bricks_query_without_gpu():
    """Host-side layout query of the C ABI (no device work)."""
    ext = _native.ext()
    groups, per_block = ext.covering_bricks(4960, 3)
    assert sum(groups) == 155 and len(groups) % per_block == 0 and max(groups) <= 8
    groups, per_block = ext.covering_bricks(252, 5)
    assert sum(groups) == 8
    groups, per_block = ext.covering_bricks(1, 3)
    assert sum(groups) == 1
    lib = _native.cdll()
    assert lib.flood_covering_bricks(0, 3, None, 0, None) == -1


def test_library_error_codes_become_python_exceptions():
    """An error code returned across the C ABI must reach Python as a RuntimeError carrying
    flood_last_error() -- not crash the interpreter.  Run in a subprocess so that a crash is a test
    failure instead of the end of the session."""
    import subprocess
    import sys

    code = ("import sys; sys.path.insert(0, %r)\n"
            "from flooder_b200 import _native\n"
            "ext = _native.ext()\n"
            "ext._raise_for_code(0)\n"
            "try:\n"
            "    ext._raise_for_code(-4)\n"
            "except RuntimeError as exc:\n"
            "    assert 'failed (-4)' in str(exc)\n"
            "    print('ok')\n") % ROOT
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert res.returncode == 0 and "ok" in res.stdout, (res.returncode, res.stderr[-500:])
