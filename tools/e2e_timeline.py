"""Host-side timeline of flood_complex calls issued back to back (development helper): per-stage host
seconds WITHOUT extra synchronisations, next to the wall time of the call.

    python tools/e2e_timeline.py [workload] [calls]
"""
import gc
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import flooder_b200 as fb
from flooder_b200 import core

dev = torch.device("cuda", 0)
job = bench.Job(sys.argv[1] if len(sys.argv) > 1 else "torus_1m_1k", dev, 0, 1)
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 6
for mode in ("plain", "host-stages", "host-stages-nogc"):
    core.PROFILE_STAGES = False if mode == "plain" else "host"
    if mode.endswith("nogc"):
        gc.disable()
    fb.flood_complex(job.host_pts.to(dev, non_blocking=True), job.n_lms, points_per_edge=job.ppe)
    torch.cuda.synchronize()
    for i in range(calls):
        t0 = time.perf_counter()
        dpts = job.host_pts.to(dev, non_blocking=True)
        t1 = time.perf_counter()
        res = fb.flood_complex(dpts, job.n_lms, points_per_edge=job.ppe)
        t2 = time.perf_counter()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        stages = " ".join(f"{k} {1e3 * v:.2f}" for k, v in core.last_stage_seconds.items()) if mode != "plain" else ""
        print(f"[{mode}] call {i}: wall {1e3 * (t3 - t0):6.2f} ms (h2d issue {1e3 * (t1 - t0):.2f}, flood_complex {1e3 * (t2 - t1):.2f}) {stages}",
              flush=True)
    gc.enable()
