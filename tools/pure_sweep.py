"""How fast is the eval kernel when every point is a candidate (no pruning, long streams)?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import flooder_b200 as fb
from flooder_b200 import _native
ext = _native.ext()
opts = dict(kv.split("=") for kv in sys.argv[1:] if "=" in kv)
for k, v in opts.items():
    if k != "R":
        ext.set_option(k, int(v))
torch.manual_seed(0)
n = 400_000
pts = torch.rand(n, 3, device="cuda")
S = 148 * 2
base = torch.tensor([[0.1, 0.1, 0.1], [0.9, 0.1, 0.1], [0.1, 0.9, 0.1], [0.1, 0.1, 0.9]], device="cuda")
verts = (base[None] + 0.001 * torch.rand(S, 4, 3, device="cuda")).contiguous()
R = int(opts.pop("R", 0)) if "R" in opts else 0
w = fb.core._grid_weights(30, 3, "cuda")
if R:
    w = torch.rand(R, 4, device="cuda"); w = (w / w.sum(1, keepdim=True)).contiguous()
ws = ext.cloud_build(pts, 0)
c, r = ext.bounding_balls(verts)
r = r * 3.0   # every ball swallows the unit cube
for rep in range(4):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    md2, cnt, ev, executed = ext.covering_radius(ws, n, 3, verts, w, None, c, r)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b); E = int(ev.item())
    print(f"opts={opts} {ms:.2f} ms E={E:.3e} evals/s={E/ms*1e3:.4e} frac={E/ms*1e3/5.3178e12:.3f} cand/simplex={cnt.float().mean().item():.0f}")
