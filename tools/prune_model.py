"""CPU model of the pruned sweep (development helper): which share of the algorithmic evaluations E
survives a box test when the sample boxes hold 8 / 4 / 2 / 1 groups of 32 samples?  Bounds are the
final minima (exact KD-tree), so the numbers are lower bounds for any order of evaluation.

    python tools/prune_model.py [n_points]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flooder_b200 as fb
from flooder_b200 import core, bricks
from flooder_b200.simplex_tree import delaunay_cells
from oracle import flood_oracle, native
from scipy.spatial import cKDTree

torch.manual_seed(0); np.random.seed(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
L = 1000
pts = fb.generate_noisy_torus_points_3d(N).numpy().astype(np.float32)
t=time.time()
lms_idx = native.fps(pts, L, 0) if hasattr(native, "fps") else None
print("fps", time.time()-t)
lms = pts[lms_idx]
cells = delaunay_cells(lms)
S = len(cells)
w = core._grid_weights(30, 3, "cpu").numpy()
R = w.shape[0]
print("S", S, "R", R)
tree = cKDTree(pts)
rng = np.random.default_rng(1)
sel = rng.choice(S, 40, replace=False)
groups = (R + 31)//32
def layout(gpb):  # bricks of gpb groups
    nbr = (groups + gpb - 1)//gpb
    base, rem = divmod(groups, nbr)
    return [base + (1 if b < rem else 0) for b in range(nbr)]
res = {}
for gpb in (8, 4, 2, 1):
    lay = layout(gpb)
    perm = bricks.brick_order(w, lay, len(lay))
    wp = w[perm]
    tot_surv = 0; tot_E = 0; tot_ideal = 0
    for s in sel:
        V = lms[cells[s]].astype(np.float64)
        X = (wp.astype(np.float64) @ V)
        # bounding ball: midpoint of longest edge? use oracle
        c, r = flood_oracle.bounding_balls(lms[cells[s]][None].astype(np.float32), 3)
        c = c[0].astype(np.float64); r = float(r[0])
        cand_idx = np.array(tree.query_ball_point(c, r), dtype=np.int64)
        P = pts[cand_idx].astype(np.float64)
        sub = cKDTree(P)
        d, _ = sub.query(X)
        off = 0
        for g in lay:
            n = min(g*32, R-off)
            xb = X[off:off+n]; db = d[off:off+n]; off += n
            lo = xb.min(0); hi = xb.max(0); u = db.max()
            e = np.maximum(np.maximum(lo - P, P - hi), 0)
            surv = ((e*e).sum(1) <= u*u).sum()
            tot_surv += surv * n
            # ideal: per-sample
        tot_E += len(P) * R
    print(f"groups/brick {gpb}: bricks {len(lay)} survivors-evals/E (final-u lower bound) = {tot_surv/tot_E:.4f}")
