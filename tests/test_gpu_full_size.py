"""Parity at BASELINE.json's full size (noisy torus 1 M points, 1 k landmarks, 30 points per edge)
through size-independent properties, plus spot checks against the CPU oracle on a sample of
simplices (the full oracle run takes minutes; a KD-tree query for a few dozen simplices seconds)."""
import numpy as np
import pytest
import torch

import flooder_b200 as fb
from flooder_b200 import _native
from flooder_b200.simplex_tree import delaunay_cells
from oracle import flood_oracle, native
from tests.helpers import seed_all

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda")


@pytest.fixture(scope="module")
def torus_1m():
    seed_all()
    pts = fb.generate_noisy_torus_points_3d(1_000_000)
    dev = pts.to(DEV)
    idx = fb.fps_indices(dev, 1000, 0)
    return pts, dev, idx


def test_fps_full_size_matches_oracle(torus_1m):
    pts, dev, idx = torus_1m
    want = native.fps(pts.numpy(), 1000, 0)                       # ~2 s of scalar C
    np.testing.assert_array_equal(idx.cpu().numpy(), want)
    grid = fb.fps_indices(dev, 1000, 0, method="grid")
    assert torch.equal(grid, idx)


def test_flood_complex_full_size_properties(torus_1m):
    pts, dev, idx = torus_1m
    lms = dev[idx]
    st = fb.flood_complex(dev, lms, return_simplex_tree=True)
    fc = dict(st._f)
    assert len(fc) == st.num_simplices() and st.num_vertices() == 1000
    vals = np.array(list(fc.values()))
    assert np.isfinite(vals).all() and (vals >= 0).all()
    for s, f in fc.items():                                        # landmarks are cloud points
        if len(s) == 1:
            assert f == 0.0
    for s, f in st.get_simplices():                                # filtered complex
        for face, ff in st.get_boundaries(s):
            assert ff <= f
    # idempotence: a second call returns the same bits (order of atomics does not matter)
    again = fb.flood_complex(dev, lms)
    assert again == fc
    # integer landmarks argument == explicit landmarks
    assert fb.flood_complex(dev, 1000) == fc

    # spot check against the oracle: exact KD-tree distances for a sample of tetrahedra
    from scipy.spatial import KDTree

    host_lms = lms.cpu().numpy()
    cells = delaunay_cells(host_lms)
    rng = np.random.default_rng(0)
    sel = rng.choice(len(cells), size=24, replace=False)
    weights, vertex_idxs, face_idxs = flood_oracle.generate_grid(30, 3)
    x = flood_oracle.sample_points(weights, host_lms[cells[sel]])
    dist, _ = KDTree(pts.numpy()).query(x.reshape(-1, 3))
    dist = dist.reshape(len(sel), -1)
    for rows, vsel in zip(face_idxs, vertex_idxs):
        for j in range(rows.shape[0]):
            want = dist[:, rows[j]].max(axis=1)
            for i, cell in enumerate(cells[sel]):
                key = tuple(int(v) for v in cell[vsel[j]])
                got = fc[key]
                # the complex keeps the smallest value over the cofaces of a shared face; the
                # sampled coface gives an upper bound, equal for the cell itself
                if len(key) == 4:
                    assert abs(got - want[i]) <= 1e-5 * want[i] + 1e-7
                else:
                    assert got <= want[i] * (1 + 1e-5) + 1e-7


def test_work_count_full_size(torus_1m):
    """The kernel's own counters (E and per-simplex candidate counts) against the plain-C ball
    count on a sample of simplices."""
    pts, dev, idx = torus_1m
    ext = _native.ext()
    lms = dev[idx]
    cells = delaunay_cells(lms.cpu().numpy())
    verts = lms[torch.as_tensor(cells, device=DEV)].contiguous()
    w = fb.core._grid_weights(30, 3, DEV)
    ws = ext.cloud_build(dev, 0)
    c, r = ext.bounding_balls(verts)
    md2, cnt, ev, executed = ext.covering_radius(ws, dev.shape[0], 3, verts, w, None, c, r)
    assert int(ev.item()) == int(cnt.sum().item()) * w.shape[0]
    rng = np.random.default_rng(1)
    sel = rng.choice(len(cells), size=40, replace=False)
    want = native.ball_counts(pts.numpy(), c[sel].cpu().numpy(), r[sel].cpu().numpy())
    np.testing.assert_array_equal(cnt[sel].cpu().numpy(), want)
    # per-sample minima of the same simplices: bit-exact against the brute-force oracle
    x = native.sample_points(w.cpu().numpy(), verts[sel[:4]].cpu().numpy())
    ref = native.min_dist(pts.numpy(), x, c[sel[:4]].cpu().numpy(), r[sel[:4]].cpu().numpy())
    np.testing.assert_array_equal(np.sqrt(md2[sel[:4]].cpu().numpy()), ref)
