import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import flooder_b200 as fb
from flooder_b200 import _native
ext = _native.ext()
torch.manual_seed(42); np.random.seed(42)
pts = fb.generate_noisy_torus_points_3d(1_000_000).cuda()
print("pts checksum", float(pts.double().sum()))
ref0 = None
for mode in (0, 1):
    ext.set_option("fps_barrier", mode)
    ref = ext.fps(pts, 1000, 0)
    if ref0 is None:
        ref0 = ref
    print("mode", mode, "checksum", int(ref.sum()), "equals mode-0 result:", bool(torch.equal(ref, ref0)))
    bad = 0
    for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 30):
        out = ext.fps(pts, 1000, 0)
        if not torch.equal(out, ref):
            bad += 1
            first = int((out != ref).nonzero()[0])
            print("mode", mode, "run", i, "first mismatch at", first)
    print("mode", mode, "mismatching runs:", bad)
