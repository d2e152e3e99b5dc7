"""End-to-end flood_complex on the larger BASELINE.json configurations (one GPU)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import flooder_b200 as fb
from flooder_b200 import core as fcore

def run(name, pts, n_lms, **kw):
    dev = pts.cuda()
    fb.flood_complex(dev[:20000], 50, points_per_edge=10)          # warm-up
    torch.cuda.synchronize()
    fcore.PROFILE_STAGES = True
    t0 = time.perf_counter()
    st = fb.flood_complex(dev, n_lms, return_simplex_tree=True, **kw)
    torch.cuda.synchronize()
    wall_prof = time.perf_counter() - t0
    stages = dict(fcore.last_stage_seconds)
    fcore.PROFILE_STAGES = False
    t0 = time.perf_counter()
    st = fb.flood_complex(dev, n_lms, return_simplex_tree=True, **kw)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    t0 = time.perf_counter()
    st.compute_persistence()
    ph = time.perf_counter() - t0
    vals = np.array([f for _, f in st.get_simplices()])
    print(f"[{name}] n={len(pts)} lms={n_lms} {kw}: wall {wall:.3f} s, {st.num_simplices()} simplices, "
          f"values in [{vals.min():.4g}, {vals.max():.4g}], PH (python) {ph:.2f} s, "
          f"H1 classes {len(st.persistence_intervals_in_dimension(1))}")
    print("    stages:", {k: round(v, 4) for k, v in stages.items()}, flush=True)

torch.manual_seed(42); np.random.seed(42)
which = sys.argv[1:] or ["cheese1m", "gauss10m"]
if "cheese1m" in which:
    run("cheese1m", fb.generate_swiss_cheese_points(1_000_000, (0, 0, 0), (1, 1, 1), 6, (0.1, 0.2))[0], 1000)
if "gauss10m" in which:
    run("gauss10m", torch.randn(10_000_000, 3), 5000)
if "torus2k" in which:
    run("torus1m_2k", fb.generate_noisy_torus_points_3d(1_000_000), 2000)
if "uni5d" in which:
    run("uni5d", torch.rand(2_000_000, 5), 2000, points_per_edge=6)
