"""Host-side timeline of one device step (development helper): where does the host spend time
before the evaluation kernel is enqueued?  Prints per-call host milliseconds for a few steps and
flags steps whose GPU time exceeds the kernel time by more than 3 ms."""
import gc
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from flooder_b200 import _native, core

ext = _native.ext()
ext.set_option("time_kernels", 1)
dev = torch.device("cuda", 0)
job = bench.Job(sys.argv[1] if len(sys.argv) > 1 else "torus_1m_1k", dev, 0, 1)
mode = sys.argv[2] if len(sys.argv) > 2 else "plain"
if mode == "nogc":
    gc.disable()
w = job.weights
perm, wp = core._kernel_sample_order(w, 3, ("lattice", tuple(w.shape), w.data_ptr()))
support = core._support_masks(wp)
for prune in (0, -1):
    ext.set_option("prune", prune)
    for rep in range(12):
        torch.cuda.synchronize()
        ext.kernel_ms("cover_eval", True)
        t = [time.perf_counter()]
        e0 = torch.cuda.Event(enable_timing=True); e0.record()
        cloud = core.PreparedCloud(job.pts); t.append(time.perf_counter())
        c, r = ext.bounding_balls(job.verts); t.append(time.perf_counter())
        order = torch.argsort(r, descending=True)
        v2, c2, r2 = job.verts[order].contiguous(), c[order].contiguous(), r[order].contiguous(); t.append(time.perf_counter())
        md2, cnt, evals, executed = ext.covering_radius(cloud.workspace, cloud.n, cloud.d, v2, wp, None, c2, r2); t.append(time.perf_counter())
        vals = ext.face_max(md2, support, 4); t.append(time.perf_counter())
        e1 = torch.cuda.Event(enable_timing=True); e1.record()
        torch.cuda.synchronize(); t.append(time.perf_counter())
        k, _ = ext.kernel_ms("cover_eval", True)
        gpu = e0.elapsed_time(e1)
        host = [1e3 * (b - a) for a, b in zip(t[:-1], t[1:])]
        flag = "  <-- gap" if gpu - k > 3 else ""
        print(f"prune={prune} rep={rep} gpu {gpu:7.2f} kernel {k:7.2f} | host ms: cloud {host[0]:.2f} balls {host[1]:.2f} sort+gather {host[2]:.2f} "
              f"covering_radius {host[3]:.2f} face_max {host[4]:.2f} sync {host[5]:.2f}{flag}", flush=True)
        del md2, vals, cloud
