#!/usr/bin/env python
"""Benchmark of the Flood-complex hot path (BASELINE.json: point-simplex distance evaluations/s
and flood_complex wall time; noisy torus 1 M points, 1 k landmarks, 3-D, 30 points per edge).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One JSON line on stdout (rank 0).  Definitions (DESIGN.md section "Measurement"):

* unit of work = one evaluation = (sample point on a simplex, cloud point inside the simplex's
  reference bounding ball) squared distance + running-min update; the algorithmic count
  E = sum_s R * |ball(s) & cloud| is produced by the kernel's own counter and does not depend
  on tiling or padding.
* step = one pass of the device path of flood_complex over the whole job, enqueued by the same
  function the product calls (flooder_b200.core.device_pass, after the cloud build): cell grid,
  bounding balls, covering-radius kernel, per-face maxima (and, when N > 1, cost plan, partition
  and the NCCL all-gather of the per-simplex values); the simplex list is sharded over the N
  ranks, the cloud is replicated (strong scaling: the job is fixed).
* value / ms_per_step / roofline: the step with EVERY evaluation executed (library option
  prune=0), i.e. the kernel whose executed work equals E; roofline = FP32 issue slots of the
  direct-difference form, 2D+1 = 7 per evaluation, peak = SMs x 128 lanes x
  max SM clock.
* default_path: the same step as the product runs it by default -- the sweep is pruned EXACTLY
  (bit-identical minima, tests/test_gpu_kernels.py::test_pruning_is_exact): executed_frac of E is
  evaluated; algorithmic_speedup = exhaustive step time / default step time.
* e2e = the metric through the public API with host buffers: pinned-host cloud -> device copy +
  flood_complex(points, n_landmarks) (FPS, host Delaunay, kernels, device->host read of the
  values, assembly of the complex), wall clock with synchronisation on both sides; default
  (pruned) path, the exhaustive figure beside it.
* cpu_baseline / --impl reference: the reference's CPU path (exact KD-tree nearest neighbour per
  sample point, scipy) restated in oracle/, timed on a bounded sample of the same job's simplices
  with all host threads; plus the UNMODIFIED reference (baseline/_ref, gudhi/fpsample stand-ins)
  on BASELINE configs[0] (torus 10 k / 100 landmarks), where the whole job fits.
* gpu_reference (N = 1): the unmodified reference's own Triton path on the same GPU and inputs.
* configs (N = 1; gauss_10m_5k also sharded at N > 1): the other BASELINE.json configurations.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "point-simplex distance evaluations/s (flood_complex, noisy torus 1M pts, 1k landmarks, 3D)"
UNIT = "evals/s"
WORKLOADS = {
    # name: (generator, n_points, n_landmarks, dim, points_per_edge, description)
    "torus_1m_1k": ("torus", 1_000_000, 1000, 3, 30, "noisy torus 1000000 points, 1000 landmarks, 3D (BASELINE.json configs[1])"),
    "torus_10k_100": ("torus", 10_000, 100, 3, 30, "noisy torus 10000 points, 100 landmarks, 3D (BASELINE.json configs[0])"),
    "cheese_1m_1k": ("cheese", 1_000_000, 1000, 3, 30, "swiss cheese 1000000 points, 1000 landmarks, 3D (BASELINE.json configs[2])"),
    "gauss_10m_5k": ("gauss", 10_000_000, 5000, 3, 30, "standard Gaussian 10000000 points, 5000 landmarks, 3D (BASELINE.json configs[3])"),
    "uniform5d_2m_2k_ppe6": ("uniform", 2_000_000, 2000, 5, 6, "uniform 5D 2000000 points, 2000 landmarks, 6 points per edge (BASELINE.json configs[4])"),
}
EXTRA_CONFIGS = ["cheese_1m_1k", "gauss_10m_5k", "uniform5d_2m_2k_ppe6"]
SLOTS_PER_EVAL = {1: 3, 2: 5, 3: 7, 4: 9, 5: 11, 6: 13, 7: 15, 8: 17}


def make_cloud(kind: str, n: int, dim: int):
    import torch

    import flooder_b200 as fb

    torch.manual_seed(42)
    np.random.seed(42)
    if kind == "torus":
        return fb.generate_noisy_torus_points_3d(n)
    if kind == "cheese":
        return fb.generate_swiss_cheese_points(n, (0,) * dim, (1,) * dim, 6, (0.1, 0.2))[0].cpu()
    if kind == "gauss":
        return torch.randn(n, dim)
    if kind == "uniform":
        return torch.rand(n, dim)
    raise ValueError(kind)


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.device_index = device_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.device_index)],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001  (nvidia-smi missing)
            self.proc = None

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        rows = []
        with open(self.path) as fh:
            for line in fh:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) >= 9:
                    rows.append(parts)
        os.unlink(self.path)
        if not rows:
            return None
        sm = sorted(float(r[1]) for r in rows if r[1].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [v for v in sm if v > 0.5 * max(sm)] or sm
        return {"sm_mhz": busy[len(busy) // 2] if busy else None,
                "sm_max_mhz": float(rows[0][2]) if rows[0][2].replace(".", "").isdigit() else None,
                "samples": len(rows), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
# CPU arms (oracle port of the reference's CPU path; the unmodified reference on configs[0])
# --------------------------------------------------------------------------------------------
def cpu_reference_sample(points: np.ndarray, simplex_vertices: np.ndarray, weights: np.ndarray,
                         cand_counts: np.ndarray, tree, workers: int):
    """Time the reference's CPU distance step (flooder/core.py:188, 197-199: sample points, one
    KDTree.query over all of them, float64) on the given simplices; returns (seconds, evaluations)."""
    from oracle import flood_oracle

    t0 = time.perf_counter()
    x = flood_oracle.sample_points(weights, simplex_vertices)             # core.py:188
    dist, _ = tree.query(x.reshape(-1, x.shape[-1]), workers=workers)     # core.py:197-199
    _ = dist.reshape(x.shape[0], x.shape[1]).max(axis=1)
    dt = time.perf_counter() - t0
    evals = int(cand_counts.sum()) * weights.shape[0]
    return dt, evals


def reference_config1():
    """The UNMODIFIED reference (baseline/_ref/flooder/core.py, CPU path) on BASELINE configs[0]:
    noisy torus 10 k points / 100 landmarks, whole job, with the gudhi/fpsample stand-ins."""
    from oracle import flood_oracle, ref_shims

    try:
        fl = ref_shims.import_reference(os.path.join(ROOT, "baseline", "_ref"))
    except ImportError as exc:
        return {"unavailable": str(exc)}
    import torch

    kind, n, n_lms, dim, ppe, desc = WORKLOADS["torus_10k_100"]
    pts = make_cloud(kind, n, dim)
    fl.flood_complex(pts[:2000], 20, points_per_edge=5)          # warm-up: lazy imports (sympy via torch)
    t0 = time.perf_counter()
    lms = fl.generate_landmarks(pts, n_lms, start_idx=0)
    t_fps = time.perf_counter() - t0
    t0 = time.perf_counter()
    res = fl.flood_complex(pts, lms, points_per_edge=ppe)
    wall = time.perf_counter() - t0
    E = flood_oracle.algorithmic_evals(pts.numpy(), lms.numpy(), points_per_edge=ppe)
    return {"kind": "_ref", "workload": desc, "value": E / wall, "unit": UNIT, "cores": 1,
            "flood_complex_wall_s": wall, "landmark_fps_s": t_fps, "evals": E, "simplices": len(res),
            "note": "unmodified reference core.py (CPU path: scipy KDTree.query, single thread as the reference "
                    "calls it) from baseline/_ref with the gudhi/fpsample stand-ins of oracle/ref_shims.py; whole job"}


def cpu_arm(workload: str, per_step: int, steps: int, warmup: int, threads: int = -1):
    """Oracle port of the reference CPU path on bounded samples of the workload's simplices."""
    from scipy.spatial import KDTree

    from oracle import flood_oracle, native
    from oracle.simplex_tree import delaunay_top_simplices

    kind, n, n_lms, dim, ppe, desc = WORKLOADS[workload]
    pts = make_cloud(kind, n, dim).numpy()
    t0 = time.perf_counter()
    lms = pts[native.fps(pts, n_lms, 0)]
    fps_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    cells = delaunay_top_simplices(lms)
    delaunay_s = time.perf_counter() - t0
    weights, _, _ = flood_oracle.generate_grid(ppe, dim)
    verts_all = lms[cells]
    centers, radii = flood_oracle.bounding_balls(verts_all, dim)
    t0 = time.perf_counter()
    tree = KDTree(pts)                                                    # core.py:128
    build_s = time.perf_counter() - t0
    order = np.random.default_rng(0).permutation(len(cells))
    times, evals = [], []
    for step in range(warmup + steps):
        sel = order[(step * per_step) % len(cells):][:per_step]
        counts = native.ball_counts(pts, centers[sel], radii[sel])
        dt, ev = cpu_reference_sample(pts, verts_all[sel], weights, counts, tree, threads)
        if step >= warmup:
            times.append(dt)
            evals.append(ev)
    sel = order[:per_step]
    counts = native.ball_counts(pts, centers[sel], radii[sel])
    dt1, ev1 = cpu_reference_sample(pts, verts_all[sel], weights, counts, tree, 1)
    total_t, total_e = float(np.sum(times)), float(np.sum(evals))
    frac = per_step / len(cells)
    one_off = build_s + fps_s + delaunay_s
    return {
        "value": total_e / total_t, "unit": UNIT, "cores": os.cpu_count() if threads < 0 else threads, "kind": "port",
        "sample": f"{per_step} of {len(cells)} simplices per step (random, seeded), {weights.shape[0]} samples each; "
                  f"{steps} timed steps after {warmup} warm-up; scipy KDTree.query(workers=-1) over a tree of the full "
                  f"{n}-point cloud",
        "host_cores": os.cpu_count(),
        "value_single_thread": ev1 / dt1,
        "note": "the reference calls KDTree.query without workers= (one thread, value_single_thread); value uses "
                "every host thread",
        "one_off_seconds": {"kdtree_build": build_s, "landmark_fps_scalar_c": fps_s, "delaunay_qhull": delaunay_s},
        "ms_per_step": 1e3 * total_t / max(1, steps),
        "e2e_value": total_e / (total_t + steps * frac * one_off),
        "projected_full_job_s": one_off + (total_t / max(1, steps)) / frac,
    }


def run_reference(args):
    """--impl reference: the reference's CPU path on the host cores, rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    arm = cpu_arm(args.workload, args.ref_simplices_per_step, args.steps, args.warmup)
    cpu = {k: arm[k] for k in ("value", "unit", "cores", "kind", "sample", "host_cores", "value_single_thread",
                                "note", "one_off_seconds", "projected_full_job_s")}
    if not args.no_ref_config1:
        cpu["reference_config1"] = reference_config1()
    line = {
        "impl": "reference", "metric": METRIC, "value": arm["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": arm["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config_dict(args.workload, args.gpus, None),
        "cpu_baseline": cpu,
        "e2e": {"value": arm["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "note": "distance step plus the sample's proportional share of the one-off host work (KD-tree build, "
                        "landmark FPS, Delaunay)"},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# the reference's own GPU path (Triton) on this GPU
# --------------------------------------------------------------------------------------------
def reference_triton(workload: str, runs: int = 3):
    """The UNMODIFIED reference package from baseline/_ref, its own Triton path on this GPU
    (flooder/core.py:193-226 with batch_size=64 as in examples/example_02_torus_3d.py), with the
    gudhi/fpsample stand-ins of oracle/ref_shims.py.  Returns a dict (or {"unavailable": why})."""
    import torch

    from oracle import ref_shims

    try:
        fl = ref_shims.import_reference(os.path.join(ROOT, "baseline", "_ref"))
    except ImportError as exc:
        return {"unavailable": str(exc)}
    kind, n, n_lms, dim, ppe, desc = WORKLOADS[workload]
    try:
        pts = make_cloud(kind, n, dim)
        dev = torch.device("cuda")
        lms = fl.generate_landmarks(pts, n_lms, start_idx=0).to(dev)
        dpts = pts.to(dev)
        core = fl.core
        stats = {"mask_ms": 0.0, "filt_ms": 0.0, "cand": 0}
        orig_mask, orig_filt = core.compute_mask, core.compute_filtration

        def timed(fn, key):
            def wrapper(*a, **k):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out = fn(*a, **k)
                e1.record()
                e1.synchronize()
                stats[key] += e0.elapsed_time(e1)
                return out
            return wrapper

        fl.flood_complex(dpts[:10000], lms, use_triton=True, points_per_edge=ppe)   # warm-up as in the examples
        torch.cuda.synchronize()
        walls = []
        for _ in range(runs):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = fl.flood_complex(dpts, lms, batch_size=64, use_triton=True, points_per_edge=ppe)
            torch.cuda.synchronize()
            walls.append(time.perf_counter() - t0)

        def mask_counting(points, centers, radii, *rest):     # instrumented pass: kernel split + work count
            out = timed(orig_mask, "mask_ms")(points, centers, radii, *rest)
            stats["cand"] += int(out[:, : points.shape[0]].sum().item())
            return out

        core.compute_mask = mask_counting
        core.compute_filtration = timed(orig_filt, "filt_ms")
        try:
            fl.flood_complex(dpts, lms, batch_size=64, use_triton=True, points_per_edge=ppe)
            torch.cuda.synchronize()
        finally:
            core.compute_mask, core.compute_filtration = orig_mask, orig_filt
    except Exception as exc:  # noqa: BLE001  (the reference arm must not take the benchmark down)
        return {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}
    from math import comb

    E = stats["cand"] * comb(ppe + dim - 1, dim)
    wall = float(np.median(walls))
    return {
        "impl": "reference Triton path (unmodified, baseline/_ref), batch_size=64", "workload": desc,
        "kernel_evals_per_s": E / (stats["filt_ms"] * 1e-3), "evals": E,
        "kernel_ms": {"compute_filtration": stats["filt_ms"], "compute_mask": stats["mask_ms"]},
        "flood_complex_wall_s": wall, "walls": walls, "e2e_evals_per_s": E / wall, "simplices": len(res),
        "note": "kernel_evals_per_s = E / time inside the reference's compute_filtration kernel (CUDA events, "
                "instrumented pass); flood_complex_wall_s = device-resident inputs, landmarks precomputed, median",
    }


def config_dict(workload, gpus, extra):
    kind, n, n_lms, dim, ppe, desc = WORKLOADS[workload]
    cfg = {"workload": desc, "n_points": n, "n_landmarks": n_lms, "dim": dim, "points_per_edge": ppe,
           "parallelism": f"simplices sharded over {gpus} GPU(s), cloud replicated",
           "l2": "flushed between timed steps (256 MiB write)"}
    if extra:
        cfg.update(extra)
    return cfg


# --------------------------------------------------------------------------------------------
# the CUDA arm
# --------------------------------------------------------------------------------------------
class Job:
    """One workload resident on this rank's GPU: cloud, landmarks (FPS), Delaunay cells, weights."""

    def __init__(self, workload, dev, rank, world):
        import torch

        import flooder_b200 as fb
        from flooder_b200 import core
        from flooder_b200 import distributed as fdist
        from flooder_b200.simplex_tree import delaunay_cells

        self.kind, self.n, self.n_lms, self.dim, self.ppe, self.desc = WORKLOADS[workload]
        self.workload, self.dev, self.rank, self.world = workload, dev, rank, world
        self.host_pts = make_cloud(self.kind, self.n, self.dim).pin_memory()
        self.pts = self.host_pts.to(dev, non_blocking=True)
        self.lms = self.pts[fb.fps_indices(self.pts, self.n_lms, 0)]
        t0 = time.perf_counter()
        self.cells = delaunay_cells(self.lms.cpu().numpy())
        self.delaunay_s = time.perf_counter() - t0
        self.S = len(self.cells)
        self.verts = self.lms[torch.as_tensor(self.cells, device=dev)].contiguous()
        self.weights = core._grid_weights(self.ppe, self.dim, dev)
        self.R, self.K = self.weights.shape
        self.shard = fdist.Shard(rank, world) if world > 1 else None

    def device_step(self, stats=None):
        """The device path of one flood_complex call, through the product's own functions."""
        from flooder_b200 import core

        cloud = core.PreparedCloud(self.pts)
        return core.device_pass(cloud, self.verts, self.weights, True, self.shard, stats)


def timed_steps(job, steps, flush, barrier):
    """K timed device steps (CUDA events, L2 flushed before each); returns (per-step ms list, the
    last step's evals / executed counters of this rank)."""
    import torch

    out, stats = [], {}
    for _ in range(steps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        job.device_step(stats)
        b.record()
        b.synchronize()
        out.append(a.elapsed_time(b))
    barrier()
    return out, int(stats["evals"].item()), int(stats["executed"].item())


def measure_job(job, ext, steps, warmup, flush, barrier, reduce_max, reduce_sum, sm_count, sm_max_mhz):
    """Exhaustive and default (pruned) timed regions for one job; returns the result dict."""
    slots = SLOTS_PER_EVAL[job.dim]
    peak_slots = sm_count * 128 * sm_max_mhz * 1e6
    res = {}
    for mode, prune in (("exhaustive", 0), ("default", -1)):
        ext.set_option("prune", prune)
        for _ in range(warmup):
            job.device_step()
            flush.fill_(1)
        barrier()
        ext.kernel_ms("cover_eval", True)
        ext.kernel_ms("cover_seed", True)
        ext.launch_count(True)
        ms, evals_local, executed_local = timed_steps(job, steps, flush, barrier)
        launches = ext.launch_count(True)
        eval_total, eval_n = ext.kernel_ms("cover_eval", True)
        seed_total, seed_n = ext.kernel_ms("cover_seed", True)
        total_ms = reduce_max(float(np.sum(ms)))
        E = reduce_sum(float(evals_local))
        executed = reduce_sum(float(executed_local))
        kernel_ms = reduce_max(eval_total / max(1, eval_n))
        kernel_ms_mean = reduce_sum(eval_total / max(1, eval_n)) / job.world
        res[mode] = {"ms_per_step": total_ms / steps, "evals_per_step": E, "executed_per_step": executed,
                     "kernel_ms_mean_over_ranks": kernel_ms_mean,
                     "kernel_ms_per_launch": kernel_ms, "seed_ms_per_launch": seed_total / max(1, seed_n),
                     "launches": int(reduce_sum(float(launches))), "evals_per_s": E / (total_ms / steps * 1e-3),
                     "kernel_evals_per_s": (E / job.world) / (kernel_ms * 1e-3),
                     "kernel_frac": (E / job.world) / (kernel_ms * 1e-3) * slots / peak_slots}
    ext.set_option("prune", -1)
    res["slots_per_eval"] = slots
    res["peak_slots"] = peak_slots
    return res


def run_cuda(args):
    import torch
    import torch.distributed as dist

    import flooder_b200 as fb
    from flooder_b200 import _native
    from flooder_b200 import core as fcore

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback exists)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ext = _native.ext()
    ext.set_option("time_kernels", 1)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def _reduce(value, op):
        t = torch.tensor([value], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=op)
        return t.item()

    reduce_max = lambda v: _reduce(v, dist.ReduceOp.MAX)   # noqa: E731
    reduce_sum = lambda v: _reduce(v, dist.ReduceOp.SUM)   # noqa: E731

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except OSError:
        pass
    sm_max_mhz = float(peaks.get("sm_max_mhz", 1965.0))
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    job = Job(args.workload, dev, rank, world)
    n, n_lms, dim, ppe = job.n, job.n_lms, job.dim, job.ppe

    # landmark FPS on its own (device time, rank 0 reports): HBM-side evidence for the second kernel
    ext.kernel_ms("fps", True)
    for _ in range(3):
        fb.fps_indices(job.pts, n_lms, 0)
    torch.cuda.synchronize()
    fps_ms_total, fps_launches = ext.kernel_ms("fps", True)

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    wall0 = time.perf_counter()
    m = measure_job(job, ext, args.steps, args.warmup, flush, barrier, reduce_max, reduce_sum, sm_count, sm_max_mhz)
    wall_s = time.perf_counter() - wall0
    clock_info = clocks.stop() if rank == 0 else None
    exh, dflt = m["exhaustive"], m["default"]
    E = exh["evals_per_step"]

    # ---- end to end through the public API --------------------------------------------------
    e2e_call_walls = []      # per-call walls of every e2e_wall() run (diagnostic, rank-local)

    def e2e_wall(steps, **env):
        for k, v in env.items():
            os.environ[k] = v
        try:
            for _ in range(2):                                                    # warm-up
                fb.flood_complex(job.host_pts.to(dev, non_blocking=True), n_lms, points_per_edge=ppe)
            gc.collect()     # a full collection of the benchmark's own garbage must not land in a timed call
            barrier()
            t0 = time.perf_counter()
            marks = [t0]
            for _ in range(steps):
                dpts = job.host_pts.to(dev, non_blocking=True)                   # H2D from pinned memory
                res = fb.flood_complex(dpts, n_lms, points_per_edge=ppe)        # includes D2H of the values
                marks.append(time.perf_counter())
            barrier()
            e2e_call_walls.append([round(b - a, 5) for a, b in zip(marks[:-1], marks[1:])])
            return reduce_max((time.perf_counter() - t0) / steps), len(res)
        finally:
            for k in env:
                del os.environ[k]

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    e2e_s, n_simplices = e2e_wall(e2e_steps)
    ext.set_option("prune", 0)
    e2e_exh_s, _ = e2e_wall(e2e_steps)
    ext.set_option("prune", -1)
    e2e_single_s = e2e_wall(e2e_steps, FLOODER_B200_NO_SHARD="1")[0] if world > 1 else e2e_s
    # untimed diagnostic pass: per-stage seconds with a device sync after every stage
    fcore.PROFILE_STAGES = True
    fb.flood_complex(job.host_pts.to(dev, non_blocking=True), n_lms, points_per_edge=ppe)
    fcore.PROFILE_STAGES = False
    stage_seconds = {k: round(v, 5) for k, v in fcore.last_stage_seconds.items()}

    # ---- the other BASELINE configurations -----------------------------------------------------
    configs = {}
    if not args.no_configs:
        names = EXTRA_CONFIGS if world == 1 else ["gauss_10m_5k"]
        del job.verts, job.pts
        for name in names:
            torch.cuda.empty_cache()
            cj = Job(name, dev, rank, world)
            cm = measure_job(cj, ext, max(1, min(args.steps, args.config_steps)), 1, flush, barrier, reduce_max,
                             reduce_sum, sm_count, sm_max_mhz)
            entry = {"workload": cj.desc, "simplices": cj.S, "samples_per_simplex": int(cj.R),
                     "evals_per_step": cm["exhaustive"]["evals_per_step"],
                     "exhaustive": {k: cm["exhaustive"][k] for k in ("ms_per_step", "evals_per_s", "kernel_ms_per_launch")},
                     "frac": cm["exhaustive"]["kernel_frac"],
                     "default_path": {"ms_per_step": cm["default"]["ms_per_step"], "evals_per_s": cm["default"]["evals_per_s"],
                                      "executed_frac": cm["default"]["executed_per_step"] / max(1.0, cm["default"]["evals_per_step"])},
                     "host_delaunay_s": cj.delaunay_s}
            if cj.dim <= 3:         # 5-D end to end is dominated by the host triangulation / face table
                fb.flood_complex(cj.pts[:20000], 50, points_per_edge=cj.ppe)
                barrier()
                t0 = time.perf_counter()
                dpts = cj.host_pts.to(dev, non_blocking=True)
                r = fb.flood_complex(dpts, cj.n_lms, points_per_edge=cj.ppe)
                barrier()
                entry["e2e_wall_s"] = reduce_max(time.perf_counter() - t0)
                entry["returned_simplices"] = len(r)
                del dpts, r
            configs[name] = entry
            del cj

    if rank == 0:
        slots, peak_slots = m["slots_per_eval"], m["peak_slots"]
        traffic = traffic_pruned = None
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fh:
                entry = json.load(fh).get(args.workload, {})
                traffic = entry.get("dram_bytes_per_launch")
                traffic_pruned = entry.get("pruned_dram_bytes_per_step")
        except OSError:
            pass
        executed_frac = dflt["executed_per_step"] / max(1.0, dflt["evals_per_step"])
        roofline = {
            "bound": "fp32-issue (CUDA-core FP32 lane slots; neither HBM nor tensor: contraction length is D=3)",
            "kernel": "cover_eval_kernel<D, PRUNE=false> (every in-ball candidate evaluated; option prune=0)",
            "achieved": exh["kernel_evals_per_s"] * slots / 1e12, "peak": peak_slots / 1e12, "unit": "Tslot/s",
            "frac": exh["kernel_frac"], "traffic": traffic,
            "slots_per_eval": slots, "evals_per_launch": E / world, "kernel_ms_per_launch": exh["kernel_ms_per_launch"],
            "kernel_ms_mean_over_ranks": exh["kernel_ms_mean_over_ranks"],
            "kernel_evals_per_s": exh["kernel_evals_per_s"], "peak_evals_per_s": peak_slots / slots,
            "peak_source": f"{sm_count} SMs x 128 FP32 lanes x {sm_max_mhz:.0f} MHz (MEASURED_PEAKS.json sm_max_mhz)",
            "flop_view": {"achieved_tflops": exh["kernel_evals_per_s"] * (3 * dim - 1) / 1e12,
                          "peak_tflops": 2 * peak_slots / 1e12},
            "hbm_view": {"algorithmic_bytes_per_launch": 16.0 * n + 4.0 * (job.S / world) * job.R,
                         "peak_gbs": peaks.get("hbm_gbs")},
        }
        default_path = {
            "kernel": "cover_eval_kernel<D, PRUNE=true>, seed pass + full pass (what flood_complex runs by default)",
            "ms_per_step": dflt["ms_per_step"], "evals_per_s": dflt["evals_per_s"],
            "kernel_ms_per_launch": dflt["kernel_ms_per_launch"], "seed_ms_per_launch": dflt["seed_ms_per_launch"],
            "kernel_ms_mean_over_ranks": dflt["kernel_ms_mean_over_ranks"],
            "executed_frac": executed_frac,
            "frac_of_issue_peak_on_executed_work": dflt["kernel_evals_per_s"] * executed_frac * slots / peak_slots,
            "algorithmic_speedup": exh["ms_per_step"] / dflt["ms_per_step"],
            "traffic": traffic_pruned, "gpu_launches": dflt["launches"],
            "note": "exact pruning (SURVEY 8(f2)), two levels: a warp skips candidates that are at least as far "
                    "from the box of its sample brick (256 samples) as the brick's largest running minimum, and the "
                    "survivors are re-tested against the boxes and bounds of the brick's pairs of 32-sample groups; "
                    "minima are bit-identical to the exhaustive sweep, executed_frac of the algorithmic evaluations "
                    "E is performed",
        }
        fps_ms = fps_ms_total / max(1, fps_launches)
        fps_bytes = float(n) * (4 * dim + 8) * (n_lms - 1)      # N (4D + 8) per iteration
        fps_info = {
            "kernel": "fps_kernel (register-resident)" if n <= 148 * 1024 * 8 else "fps_grid_kernel (bucketed)",
            "ms": fps_ms, "us_per_landmark": 1e3 * fps_ms / max(1, n_lms - 1),
            "bound": "hbm", "achieved": fps_bytes / (fps_ms * 1e-3) / 1e9, "peak": peaks.get("hbm_gbs"),
            "unit": "GB/s (algorithmic bytes N(4D+8) per iteration)",
            "frac": (fps_bytes / (fps_ms * 1e-3) / 1e9) / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None,
            "note": "clouds up to 1.2M points stay in registers: no HBM traffic, the iteration is bound by the "
                    "grid-wide argmax + sync (not by HBM); larger clouds use the bucketed kernel",
        }
        line = {
            "metric": METRIC, "value": exh["evals_per_s"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": exh["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args.workload, world, {"simplices": job.S, "samples_per_simplex": int(job.R),
                                                         "evals_per_step": E, "returned_simplices": n_simplices}),
            "value_note": "every algorithmic evaluation executed (library option prune=0): the step the roofline "
                          "describes. flood_complex's default path prunes exactly and is faster: see default_path "
                          "(value_default_path = E / its step time) and e2e",
            "value_default_path": dflt["evals_per_s"],
            "e2e": {"value": E / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": int(n * dim * 4 * world),
                    "d2h_bytes_per_step": int((job.S * (2 ** job.K - 1) * 4 + n_lms * dim * 4) * world),
                    "flood_complex_wall_s": e2e_s, "steps": e2e_steps, "call_walls_s": e2e_call_walls[0],
                    "path": "public API, default (exactly pruned) sweep",
                    "exhaustive": {"value": E / e2e_exh_s, "flood_complex_wall_s": e2e_exh_s},
                    "single_gpu_wall_s": e2e_single_s,
                    "scaling_efficiency": e2e_single_s / (world * e2e_s),
                    "stage_seconds_serialised": stage_seconds,
                    "includes": "H2D of the cloud, landmark FPS, host Delaunay, kernels, D2H, complex assembly"},
            "gpu_launches": exh["launches"],
            "roofline": roofline,
            "default_path": default_path,
            "roofline_fps": fps_info,
            "clocks": clock_info,
            "timed_region_wall_s": wall_s,
        }
        if configs:
            line["configs"] = configs
        if world == 1 and not args.no_gpu_reference:
            line["gpu_reference"] = reference_triton(args.workload)
        if world == 1 and not args.no_cpu_baseline:
            arm = cpu_arm(args.workload, args.cpu_sample_simplices, 1, 0)
            line["cpu_baseline"] = {k: arm[k] for k in ("value", "unit", "cores", "kind", "sample", "host_cores",
                                                        "value_single_thread", "note", "one_off_seconds",
                                                        "projected_full_job_s")}
            if not args.no_ref_config1:
                line["cpu_baseline"]["reference_config1"] = reference_config1()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference", "reference_triton"])
    ap.add_argument("--workload", default="torus_1m_1k", choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--config-steps", type=int, default=2)
    ap.add_argument("--cpu-sample-simplices", type=int, default=400)
    ap.add_argument("--ref-simplices-per-step", type=int, default=200)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--no-ref-config1", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "reference_triton":
        if int(os.environ.get("RANK", "0")) == 0:
            print(json.dumps({"impl": "reference_triton", **reference_triton(args.workload)}), flush=True)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
