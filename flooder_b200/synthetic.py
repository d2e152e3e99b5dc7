"""Synthetic point clouds of the shapes the reference's tests, examples and benchmarks use.

Re-implemented (not imported: the reference tree does not exist on the GPU box) so that the same
seed yields the same cloud as ``flooder/synthetic_data_generators.py``; pinned byte-for-byte by
``tests/test_synthetic.py`` against ``tests/golden/ref_generators.npz``.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import numpy as np
import torch


def generate_noisy_torus_points_3d(n: int = 1000, R: float = 3.0, r: float = 1.0, noise_std: float = 0.02,
                                   seed: Optional[int] = None) -> torch.Tensor:
    """Uniform angles on a torus (major radius R, tube radius r) plus isotropic Gaussian noise.
    Draw order: theta, phi, noise -- all from torch's global CPU generator
    (reference ``synthetic_data_generators.py:220-269``)."""
    if seed is not None:
        torch.manual_seed(seed)
    theta = torch.rand(n) * 2 * torch.pi
    phi = torch.rand(n) * 2 * torch.pi
    ring = R + r * torch.cos(phi)
    pts = torch.stack((ring * torch.cos(theta), ring * torch.sin(theta), r * torch.sin(phi)), dim=1)
    return pts + torch.randn_like(pts) * noise_std


def generate_figure_eight_points_2d(n: int = 1000, r_bounds: Tuple[float, float] = (0.2, 0.3),
                                    centers=((0.3, 0.5), (0.7, 0.5)), noise_std: float = 0.0,
                                    noise_kind: str = "gaussian", seed: Optional[int] = None) -> torch.Tensor:
    """Two annular lobes, area-uniform in radius; numpy's global generator, draw order lobe,
    radius, angle, noise (reference ``synthetic_data_generators.py:13-69``)."""
    if seed is not None:
        np.random.seed(seed)
    lobe = np.random.randint(0, 2, size=n)
    cx, cy = np.asarray(centers).T
    lo, hi = r_bounds
    rad = np.sqrt(np.random.uniform(lo ** 2, hi ** 2, size=n))
    ang = np.random.uniform(0.0, 2 * np.pi, size=n)
    x = cx[lobe] + rad * np.cos(ang)
    y = cy[lobe] + rad * np.sin(ang)
    if noise_std > 0:
        if noise_kind == "gaussian":
            x += np.random.normal(0.0, noise_std, size=n)
            y += np.random.normal(0.0, noise_std, size=n)
        elif noise_kind == "uniform":
            x += np.random.uniform(-noise_std, noise_std, size=n)
            y += np.random.uniform(-noise_std, noise_std, size=n)
        else:
            raise ValueError("noise_kind must be 'gaussian' or 'uniform'")
    return torch.tensor(np.stack((x, y), axis=1), dtype=torch.float32)


def generate_annulus_points_2d(n: int = 1000, center: torch.Tensor = torch.tensor([0.0, 0.0]),
                               radius: float = 1.0, width: float = 0.2, seed: Optional[int] = None) -> torch.Tensor:
    """Area-uniform annulus; draw order angles, radii (reference ``:175-217``)."""
    assert center.shape == (2,), "Center must be a 2D point."
    assert radius > 0 and width > 0, "Radius and width must be positive."
    if seed is not None:
        torch.manual_seed(seed)
    ang = torch.rand(n) * 2 * torch.pi
    rad = radius - width + width * torch.sqrt(torch.rand(n))
    return torch.stack((center[0] + rad * torch.cos(ang), center[1] + rad * torch.sin(ang)), dim=1)


@torch.no_grad()
def generate_swiss_cheese_points(n: int = 1000, rect_min: Sequence[float] = (0.0, 0.0, 0.0),
                                 rect_max: Sequence[float] = (1.0, 1.0, 1.0), k: int = 6,
                                 void_radius_range: Tuple[float, float] = (0.1, 0.2), seed: Optional[int] = None,
                                 *, device="cpu", batch_factor: int = 4):
    """Box minus k disjoint balls, by rejection sampling with the generator of ``device``
    (reference ``:72-172``; like there, ``seed=0`` does not seed).  Returns (points, centres, radii)."""
    if seed:
        torch.manual_seed(seed)
    assert len(rect_min) == len(rect_max), "rect_min and rect_max must have the same dimension."
    dim = len(rect_min)
    r_lo, r_hi = void_radius_range
    lo = torch.tensor(rect_min, dtype=torch.float32, device=device)
    hi = torch.tensor(rect_max, dtype=torch.float32, device=device)
    centres = torch.empty((0, dim), device=device)
    radii = torch.empty((0,), device=device)
    while centres.shape[0] < k:
        batch = max(8, 2 * (k - centres.shape[0]))
        cand_c = (lo + r_hi) + (hi - lo - 2 * r_hi) * torch.rand(batch, dim, device=device)
        cand_r = r_lo + (r_hi - r_lo) * torch.rand(batch, device=device)
        if centres.numel() == 0:
            ok = torch.ones(batch, dtype=torch.bool, device=device)
        else:
            ok = (torch.cdist(cand_c, centres) >= (cand_r[:, None] + radii[None, :])).all(dim=1)
        keep = ok.nonzero(as_tuple=False).squeeze()[: k - centres.shape[0]]
        centres = torch.cat([centres, cand_c[keep]], dim=0)
        radii = torch.cat([radii, cand_r[keep]], dim=0)
    pts = torch.empty((0, dim), dtype=lo.dtype, device=device)
    todo = n
    while todo:
        cand = lo + (hi - lo) * torch.rand(batch_factor * todo, dim, device=device)
        if k:
            good = (torch.cdist(cand, centres) >= radii[None, :]).all(dim=1)
        else:
            good = torch.ones(cand.shape[0], dtype=torch.bool, device=device)
        pts = torch.cat([pts, cand[good][:todo]], dim=0)
        todo = n - pts.shape[0]
    return pts, centres, radii
