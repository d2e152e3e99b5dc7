"""Round-2 kernel experiments on one GPU: kernel shapes, brick ordering, seed strides.

    python tools/sweep_r2.py torus1m [cheese1m ...] -- "prune=0" "prune=1,bricks=0" "prune=1,async_gather=0" ...

Every variant is a comma-separated list of library options (plus bricks=0/1 for the host-side
sample order).  Prints kernel times (CUDA events inside the library) and executed/E.
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flooder_b200 as fb
from flooder_b200 import _native, core
from flooder_b200.simplex_tree import delaunay_cells
from tools.config_bench import CONFIGS, SLOTS, cloud


def main():
    argv = sys.argv[1:]
    split = argv.index("--") if "--" in argv else len(argv)
    names = argv[:split] or ["torus1m"]
    variants = argv[split + 1:] or ["prune=1"]
    ext = _native.ext()
    ext.set_option("time_kernels", 1)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    for name in names:
        kind, n, n_lms, d, ppe = CONFIGS[name]
        pts = cloud(kind, n, d).cuda()
        idx = fb.fps_indices(pts, n_lms, 0)
        lms = pts[idx]
        cells = delaunay_cells(lms.cpu().numpy())
        verts = lms[torch.as_tensor(cells, device="cuda")].contiguous()
        _, radii = ext.bounding_balls(verts)
        verts = verts[torch.argsort(radii, descending=True)].contiguous()    # largest balls first, as core does
        w = core._grid_weights(ppe, d, "cuda")
        peak = sms * 128 * 1.965e9 / SLOTS[d]
        ref = None
        print(f"[{name}] n={n} lms={n_lms} d={d} ppe={ppe} S={len(cells)} R={w.shape[0]}", flush=True)
        for var in variants:
            opts = dict(kv.split("=") for kv in var.split(",") if kv)
            bricks = int(opts.pop("bricks", 1))
            core.USE_BRICKS = bool(bricks)
            prev = {k: ext.set_option(k, int(v)) for k, v in opts.items()}
            try:
                pc = core.PreparedCloud(pts)          # cloud options (points_per_cell, grid_axes) apply here
                best = None
                for rep in range(3):
                    ext.kernel_ms("cover_eval", True)
                    ext.kernel_ms("cover_seed", True)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    vals, det = core.covering_values(pc, verts, w, grid_mode=True, return_details=True)
                    torch.cuda.synchronize()
                    wall = (time.perf_counter() - t0) * 1e3
                    ev_ms, _ = ext.kernel_ms("cover_eval", True)
                    seed_ms, _ = ext.kernel_ms("cover_seed", True)
                    if best is None or ev_ms < best[0]:
                        best = (ev_ms, seed_ms, wall)
                E = int(det["evals"].item())
                ex = int(det["executed"].item())
                same = "ref"
                if ref is None:
                    ref = det["min_dist2"].clone()
                else:
                    same = "bit-identical" if torch.equal(ref, det["min_dist2"]) else "DIFFERENT"
                print(f"   {var:40s} eval {best[0]:8.2f} ms (seed {best[1]:6.2f}) wall {best[2]:8.2f} | executed/E {ex / max(E, 1):.4f} | "
                      f"E/s {E / best[0] * 1e3:.3e} frac {E / best[0] * 1e3 / peak:.3f} exec-frac {ex / best[0] * 1e3 / peak:.3f} | {same}",
                      flush=True)
            finally:
                for k, v in prev.items():
                    ext.set_option(k, v)
        del pts


if __name__ == "__main__":
    main()
