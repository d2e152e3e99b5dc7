"""Quick stage timings on one GPU (development helper, not the benchmark contract)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flooder_b200 as fb
from flooder_b200 import _native
from flooder_b200.simplex_tree import delaunay_cells


def ev_time(fn, reps=3):
    torch.cuda.synchronize()
    best = 1e30
    out = None
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best, out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    n_lms = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    opts = dict(kv.split("=") for kv in sys.argv[3:])
    ext = _native.ext()
    for k, v in opts.items():
        ext.set_option(k, int(v))
    torch.manual_seed(42)
    np.random.seed(42)
    pts = fb.generate_noisy_torus_points_3d(n).cuda()
    t_fps, idx = ev_time(lambda: ext.fps(pts, n_lms, 0))
    lms = pts[idx]
    t0 = time.perf_counter()
    cells = delaunay_cells(lms.cpu().numpy())
    t_del = time.perf_counter() - t0
    verts = lms[torch.as_tensor(cells, device="cuda")].contiguous()
    w, _, _ = fb.generate_grid(30, 3, "cuda")
    t_cloud, ws = ev_time(lambda: ext.cloud_build(pts, 0))
    c, r = ext.bounding_balls(verts)
    t_cov, (md2, cnt, evals, executed) = ev_time(lambda: ext.covering_radius(ws, n, 3, verts, w, None, c, r))
    E = int(evals.item())
    print(f"n={n} lms={n_lms} S={len(cells)} R={w.shape[0]} opts={opts}")
    print(f"fps {t_fps:.2f} ms | delaunay {t_del*1e3:.1f} ms | cloud_build {t_cloud:.2f} ms | covering {t_cov:.2f} ms")
    print(f"E={E:.4e}  evals/s={E/(t_cov*1e-3):.4e}  frac_of_5.32e12={E/(t_cov*1e-3)/5.32e12:.3f}")
    print("cand/simplex mean", cnt.float().mean().item(), "max", cnt.max().item())
    t0 = time.perf_counter()
    res = fb.flood_complex(pts, lms)
    torch.cuda.synchronize()
    print(f"flood_complex wall {time.perf_counter()-t0:.3f} s, {len(res)} simplices")
    t0 = time.perf_counter()
    res = fb.flood_complex(pts, lms)
    torch.cuda.synchronize()
    print(f"flood_complex wall (2nd) {time.perf_counter()-t0:.3f} s")


if __name__ == "__main__":
    main()
