"""First-launch determinism of FPS in a fresh process (prints one checksum line)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import flooder_b200 as fb
from flooder_b200 import _native
ext = _native.ext()
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
ext.set_option("fps_barrier", mode)
torch.manual_seed(42); np.random.seed(42)
pts = fb.generate_noisy_torus_points_3d(1_000_000).cuda()
idx = ext.fps(pts, 1000, 0)
print("mode", mode, "first-launch idx checksum", int(idx.sum()), int((idx * torch.arange(1000, device='cuda')).sum()))
