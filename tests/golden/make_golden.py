"""Generate the golden fixtures under ``tests/golden/``.

Run ONCE in the authoring container (needs ``/root/reference``; the GPU box has no copy):

    python tests/golden/make_golden.py

Two kinds of fixtures are written:

``shipped_*.npz``    data files the reference authors ship with their documentation, stored
                     verbatim as arrays: ``docs/animation/*.csv`` (inputs AND filtration
                     values, produced by ``docs/animation/generate_csvs.py`` with the CPU
                     path at 31 points per edge) and ``docs/visualization/*/`` (1000
                     landmarks + the gudhi Delaunay simplices with their values; the 1 M
                     point clouds are not shipped, so these pin the simplex set only).
``ref_*.npz``        outputs of the reference's own ``flooder.flood_complex`` executed
                     unmodified here (CPU path) through ``oracle/ref_shims.py`` on small
                     seeded inputs.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_shims  # noqa: E402

REF = ref_shims.REFERENCE_ROOT


def pack(result: dict) -> dict:
    """dict{tuple: float} -> arrays per simplex size."""
    out = {}
    for k in sorted({len(s) for s in result}):
        keys = sorted(s for s in result if len(s) == k)
        out[f"simplices_{k}"] = np.asarray(keys, dtype=np.int32).reshape(len(keys), k)
        out[f"values_{k}"] = np.asarray([result[s] for s in keys], dtype=np.float64)
    return out


def shipped():
    a = os.path.join(REF, "docs", "animation")
    load = lambda n: np.loadtxt(os.path.join(a, n), delimiter=",")  # noqa: E731
    np.savez_compressed(
        os.path.join(HERE, "shipped_animation.npz"),
        points=load("points.csv").astype(np.float32),
        landmarks=load("landmarks.csv").astype(np.float32),
        edges=load("edges.csv"), triangles=load("triangles.csv"),
        points_per_edge=np.int64(31),
    )
    for name in ("virus", "coral", "lockwasher"):
        v = os.path.join(REF, "docs", "visualization", name)
        load = lambda n: np.loadtxt(os.path.join(v, n), delimiter=",")  # noqa: E731
        tets, tris, edges = load("tetrahedra.csv"), load("triangles.csv"), load("edges.csv")
        np.savez_compressed(
            os.path.join(HERE, f"shipped_{name}.npz"),
            landmarks=load("landmarks.csv").astype(np.float32),
            tetrahedra=tets[:, :4].astype(np.int32), tetrahedra_f=tets[:, 4].astype(np.float32),
            triangles=tris[:, :3].astype(np.int32), triangles_f=tris[:, 3].astype(np.float32),
            edges=edges[:, :2].astype(np.int32), edges_f=edges[:, 2].astype(np.float32),
        )


def reference_runs():
    fl = ref_shims.import_reference()

    def seed():
        torch.manual_seed(42)
        np.random.seed(42)

    def run(tag, pts, n_lms, **kw):
        seed()
        lms = fl.generate_landmarks(pts, n_lms, start_idx=0)
        seed()
        res = fl.flood_complex(pts, lms, **kw)
        meta = {k: np.int64(-1 if v is None else v) for k, v in kw.items()}
        np.savez_compressed(os.path.join(HERE, f"ref_{tag}.npz"), points=pts.numpy(),
                            landmarks=lms.numpy(), **meta, **pack(res))
        print(tag, tuple(pts.shape), n_lms, kw, "->", len(res), "simplices")

    seed()
    torus = fl.generate_noisy_torus_points_3d(3000)
    run("torus3d_grid", torus, 40, points_per_edge=12)
    run("torus3d_rand", torus, 40, points_per_edge=None, num_rand=96)
    run("torus3d_maxdim2", torus, 40, points_per_edge=9, max_dimension=2)
    run("torus3d_f64", torus.double(), 40, points_per_edge=8)
    seed()
    fig8 = fl.generate_figure_eight_points_2d(1500)
    run("fig8_2d_grid", fig8, 60, points_per_edge=25)
    run("fig8_2d_rand", fig8, 60, points_per_edge=None, num_rand=200)
    seed()
    cheese = fl.generate_swiss_cheese_points(4000)[0]
    run("cheese3d_grid", cheese, 50, points_per_edge=10)
    seed()
    run("uniform4d_grid", torch.rand(1200, 4), 24, points_per_edge=5)
    seed()
    run("uniform5d_rand", torch.rand(800, 5), 16, points_per_edge=None, num_rand=48)
    # generators: first rows, for the byte-exact check of flooder_b200.synthetic
    seed()
    t = fl.generate_noisy_torus_points_3d(64)
    seed()
    f8 = fl.generate_figure_eight_points_2d(64)
    seed()
    ch = fl.generate_swiss_cheese_points(64)[0]
    seed()
    an = fl.generate_annulus_points_2d(64)
    np.savez_compressed(os.path.join(HERE, "ref_generators.npz"), torus=t.numpy(), fig8=f8.numpy(),
                        cheese=ch.numpy(), annulus=an.numpy())


if __name__ == "__main__":
    shipped()
    reference_runs()
