"""Stage timings for the BASELINE.json configurations (development helper; one GPU).

    python tools/config_bench.py torus1m cheese1m gauss10m uni5d uni6d [key=value ...]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flooder_b200 as fb
from flooder_b200 import _native
from flooder_b200.simplex_tree import delaunay_cells

CONFIGS = {
    # name: (generator, n, n_lms, dim, ppe)
    "torus10k": ("torus", 10_000, 100, 3, 30),
    "torus1m": ("torus", 1_000_000, 1000, 3, 30),
    "torus1m2k": ("torus", 1_000_000, 2000, 3, 30),
    "cheese1m": ("cheese", 1_000_000, 1000, 3, 30),
    "gauss10m": ("gauss", 10_000_000, 5000, 3, 30),
    "uni5d": ("uniform", 2_000_000, 2000, 5, 6),
    "uni6d": ("uniform", 2_000_000, 2000, 6, 4),
    "uni4d": ("uniform", 1_000_000, 1000, 4, 10),
    "uni5d_small": ("uniform", 500_000, 600, 5, 6),
    "fig8_2d": ("fig8", 1_000_000, 2000, 2, 130),
    "torus_ppe20": ("torus", 1_000_000, 1000, 3, 20),
}
SLOTS = {2: 5, 3: 7, 4: 9, 5: 11, 6: 13}


def cloud(kind, n, d):
    torch.manual_seed(42)
    np.random.seed(42)
    if kind == "torus":
        return fb.generate_noisy_torus_points_3d(n)
    if kind == "cheese":
        return fb.generate_swiss_cheese_points(n, (0,) * d, (1,) * d, 6, (0.1, 0.2), device="cuda")[0].cpu()
    if kind == "gauss":
        return torch.randn(n, d)
    if kind == "uniform":
        return torch.rand(n, d)
    if kind == "fig8":
        return fb.generate_figure_eight_points_2d(n)
    raise ValueError(kind)


def ev_time(fn, reps=2):
    best, out = 1e30, None
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best, out


def main():
    names = [a for a in sys.argv[1:] if "=" not in a] or ["torus1m"]
    opts = dict(kv.split("=") for kv in sys.argv[1:] if "=" in kv)
    ext = _native.ext()
    for k, v in opts.items():
        ext.set_option(k, int(v))
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    for name in names:
        kind, n, n_lms, d, ppe = CONFIGS[name]
        pts = cloud(kind, n, d).cuda()
        t_fps, idx = ev_time(lambda: ext.fps(pts, n_lms, 0))
        t_cloud0, ws0 = ev_time(lambda: ext.cloud_build(pts, 0))
        t_fpsg, idxg = ev_time(lambda: ext.fps_grid(ws0, pts, n_lms, 0))
        print(f"   fps brute {t_fps:.2f} ms | fps grid {t_fpsg:.2f} ms (+ cloud {t_cloud0:.2f}) | equal {bool((idx == idxg).all())}")
        del ws0
        lms = pts[idx]
        t0 = time.perf_counter()
        cells = delaunay_cells(lms.cpu().numpy())
        t_del = time.perf_counter() - t0
        verts = lms[torch.as_tensor(cells, device="cuda")].contiguous()
        w = fb.core._grid_weights(ppe, d, "cuda")
        t_cloud, ws = ev_time(lambda: ext.cloud_build(pts, 0))
        c, r = ext.bounding_balls(verts)
        order = torch.argsort(r, descending=True)
        verts, c, r = verts[order].contiguous(), c[order].contiguous(), r[order].contiguous()
        t_cov, (md2, cnt, evals, executed) = ev_time(lambda: ext.covering_radius(ws, n, d, verts, w, None, c, r), reps=2)
        E = int(evals.item())
        peak = sms * 128 * 1.965e9 / SLOTS[d]
        print(f"[{name}] n={n} lms={n_lms} d={d} ppe={ppe} S={len(cells)} R={w.shape[0]} opts={opts}")
        print(f"   fps {t_fps:.2f} ms ({n * (4 * d + 8) * (n_lms - 1) / t_fps / 1e6:.0f} GB/s algorithmic) | "
              f"delaunay {t_del * 1e3:.0f} ms | cloud_build {t_cloud:.2f} ms | covering {t_cov:.2f} ms")
        ex = int(executed.item())
        print(f"   executed/E = {ex / max(E, 1):.3f}")
        print(f"   E={E:.4e} evals/s={E / (t_cov * 1e-3):.4e} frac_of_issue_roofline={E / (t_cov * 1e-3) / peak:.3f} "
              f"cand/simplex mean {cnt.float().mean().item():.0f} max {cnt.max().item()}", flush=True)
        if n_lms <= 2000 and d <= 3:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = fb.flood_complex(pts, lms, points_per_edge=ppe)
            torch.cuda.synchronize()
            print(f"   flood_complex wall {time.perf_counter() - t0:.3f} s, {len(res)} simplices", flush=True)
        del pts, md2, ws


if __name__ == "__main__":
    main()
