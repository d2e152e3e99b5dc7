"""Launch the FPS kernels once each on a streaming-size cloud (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flooder_b200 import _native
ext = _native.ext()
n, n_lms = int(sys.argv[1]), int(sys.argv[2])
torch.manual_seed(42)
pts = torch.randn(n, 3, device="cuda")
a = ext.fps(pts, n_lms, 0)
ws = ext.cloud_build(pts, 0)
b = ext.fps_grid(ws, pts, n_lms, 0)
torch.cuda.synchronize()
print("equal", bool(torch.equal(a, b)))
