"""Where does a bench step spend its time outside the evaluation kernel?  (development helper)

    python tools/step_breakdown.py torus_1m_1k cheese_1m_1k

Uses bench.py's own Job / device_step; prints every step's event time next to the library's kernel
timer, exhaustive and default mode, with and without the L2 flush."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from flooder_b200 import _native, core

ext = _native.ext()
ext.set_option("time_kernels", 1)
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


for name in sys.argv[1:] or ["torus_1m_1k"]:
    torch.cuda.empty_cache()
    job = bench.Job(name, dev, 0, 1)
    for prune in (0, -1):
        ext.set_option("prune", prune)
        job.device_step()
        torch.cuda.synchronize()
        for rep in range(5):
            do_flush = rep % 2 == 0
            if do_flush:
                flush.fill_(1)
            ext.kernel_ms("cover_eval", True)
            e0 = ev()
            cloud = core.PreparedCloud(job.pts)
            e1 = ev()
            core.device_pass(cloud, job.verts, job.weights, True, None, {})
            e2 = ev()
            torch.cuda.synchronize()
            k, _ = ext.kernel_ms("cover_eval", True)
            print(f"[{name}] prune={prune} flush={do_flush}: cloud {e0.elapsed_time(e1):.2f} ms, pass {e1.elapsed_time(e2):.2f} ms, "
                  f"eval kernel {k:.2f} ms, step - kernel {e0.elapsed_time(e2) - k:.2f} ms", flush=True)
    ext.set_option("prune", -1)
    del job
