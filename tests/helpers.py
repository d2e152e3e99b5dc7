"""Shared helpers for the test-suite (fixture loading, seeded inputs)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

REF_CASES = [
    "torus3d_grid", "torus3d_rand", "torus3d_maxdim2", "torus3d_f64",
    "fig8_2d_grid", "fig8_2d_rand", "cheese3d_grid", "uniform4d_grid", "uniform5d_rand",
]


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_kwargs(g):
    kw = {}
    for key in ("points_per_edge", "num_rand", "max_dimension"):
        if key in g.files:
            v = int(g[key])
            kw[key] = None if v < 0 else v
    return kw


def golden_dict(g):
    out = {}
    k = 1
    while f"simplices_{k}" in g.files:
        for s, v in zip(g[f"simplices_{k}"].tolist(), g[f"values_{k}"].tolist()):
            out[tuple(s)] = v
        k += 1
    return out


def seed_all(seed=42):
    import torch

    torch.manual_seed(seed)
    np.random.seed(seed)


def assert_close_dict(got, want, rtol, atol, what=""):
    assert set(got) == set(want), f"{what}: simplex sets differ"
    worst = 0.0
    for s, w in want.items():
        g = got[s]
        if np.isnan(w):
            assert np.isnan(g), f"{what}: {s} expected NaN, got {g}"
            continue
        err = abs(g - w)
        tol = atol + rtol * abs(w)
        assert err <= tol, f"{what}: simplex {s}: got {g!r}, want {w!r}, |diff| {err:.3e} > {tol:.3e}"
        worst = max(worst, err)
    return worst
