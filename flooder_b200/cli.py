"""Command line front-end: ``python -m flooder_b200.cli --input-file cloud.npy ...``

Same options, defaults, output payload (pickle with ``diagrams`` + ``meta``) and step statistics
as the reference's ``flooder`` console script (``flooder/cli.py:186-293, 427-500``), on top of the
sm_100a path.  Differences: ``--device cpu`` is rejected (there is no CPU path), ``--no-triton``,
``--fpsh`` and ``--batch-size`` are accepted and have no effect, and the output is plain text
(``rich_argparse`` is not required).
"""
from __future__ import annotations

import argparse
import json
import pickle
import time
from dataclasses import asdict, dataclass
from pathlib import Path
from typing import List, Optional, Tuple

import numpy as np
import torch


@dataclass
class StepStats:
    name: str
    wall_s: float
    cpu_s: float
    cuda_ms: Optional[float]
    peak_device_mb: Optional[float]


@dataclass
class RunMeta:
    input_file: str
    output_file: Optional[str]
    num_landmarks: int
    max_dimension: int
    fps_height: int
    batch_size: int
    device: str
    points_per_edge: Optional[int]
    num_rand: Optional[int]
    seed: Optional[int]
    use_triton: bool
    n_points: int
    ambient_dim: int


class StepTimer:
    """Wall / process-CPU time of a step, optional CUDA-event time and peak device memory."""

    def __init__(self, name: str, device: torch.device, use_cuda_events: bool = False):
        self.name, self.device, self.use_events = name, device, use_cuda_events and device.type == "cuda"
        self.stats: Optional[StepStats] = None

    def __enter__(self):
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)
            torch.cuda.reset_peak_memory_stats(self.device)
        if self.use_events:
            self._e0, self._e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self._e0.record()
        self._w0, self._c0 = time.perf_counter(), time.process_time()
        return self

    def __exit__(self, *exc):
        cuda_ms = None
        if self.use_events:
            self._e1.record()
            self._e1.synchronize()
            cuda_ms = self._e0.elapsed_time(self._e1)
        peak = None
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)
            peak = torch.cuda.max_memory_allocated(self.device) / 2 ** 20
        self.stats = StepStats(self.name, time.perf_counter() - self._w0, time.process_time() - self._c0, cuda_ms, peak)
        return False


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(prog="flooder_b200", description="Flood complex persistent homology (B200)")
    g0 = p.add_argument_group("Flooder options")
    g0.add_argument("--num-landmarks", metavar="INT", type=int, default=2000)
    g0.add_argument("--max-dimension", metavar="INT", type=int, default=None,
                    help="compute PH up to this dimension (exclusive); default: ambient dimension")
    g0.add_argument("--fpsh", dest="fps_height", metavar="INT", type=int, default=9, help="accepted, no effect")
    g0.add_argument("--batch-size", metavar="INT", type=int, default=64, help="accepted, no effect")
    g0.add_argument("--device", type=str, default="cuda:0", help='"cuda" or "cuda:N"')
    g0.add_argument("--seed", metavar="INT", type=int, default=None, help="only used with --num-rand")
    g0.add_argument("--no-triton", action="store_true", help="accepted, no effect (there is no Triton path)")
    mex = g0.add_mutually_exclusive_group(required=False)
    mex.add_argument("--points-per-edge", metavar="INT", type=int, default=None)
    mex.add_argument("--num-rand", metavar="INT", type=int, default=None)
    g1 = p.add_argument_group("Input/Output options")
    g1.add_argument("--input-file", metavar="FILE", type=str, required=True, help=".npy file with an (N, D) cloud")
    g1.add_argument("--output-file", metavar="FILE", type=str, default=None, help="pickle with diagrams + metadata")
    g1.add_argument("-v", "--verbose", action="store_true")
    g1.add_argument("--stats-json", metavar="FILE", type=str, default=None)
    g1.add_argument("--cuda-events", action="store_true")
    return p


def validate_device(device_str: str) -> torch.device:
    if device_str == "cuda":
        device_str = "cuda:0"
    device = torch.device(device_str)
    if device.type != "cuda":
        raise SystemExit(f"device '{device_str}' is not supported: flooder_b200 runs on CUDA devices only")
    if not torch.cuda.is_available() or (device.index or 0) >= torch.cuda.device_count():
        raise SystemExit(f"CUDA device '{device_str}' is not available")
    torch.cuda.set_device(device)
    return device


def load_point_cloud(path: Path) -> Tuple[torch.Tensor, int, int]:
    if path.suffix != ".npy":
        raise SystemExit(f"input file must be a .npy array, got '{path}'")
    arr = np.load(path, mmap_mode="r")
    if arr.ndim != 2:
        raise SystemExit(f"expected an (N, D) array, got shape {arr.shape}")
    pts = torch.from_numpy(np.array(arr, dtype=np.float32, copy=True))
    return pts, int(pts.shape[0]), int(pts.shape[1])


def resolve_simplex_representation(points_per_edge: Optional[int], num_rand: Optional[int]):
    """30 points per edge unless one of the two options is given (reference default)."""
    if points_per_edge is None and num_rand is None:
        return 30, None
    return points_per_edge, num_rand


def save_output(path: Path, diagrams, meta: RunMeta) -> Path:
    if path.suffix == "":
        path = path.with_suffix(".pkl")
    path.parent.mkdir(parents=True, exist_ok=True)
    tmp = path.with_suffix(path.suffix + ".tmp")
    with tmp.open("wb") as fh:
        pickle.dump({"diagrams": diagrams, "meta": asdict(meta)}, fh, protocol=pickle.HIGHEST_PROTOCOL)
    tmp.replace(path)
    return path


def print_stats_table(stats: List[StepStats]) -> None:
    print(f"{'step':<16}{'wall [s]':>10}{'cpu [s]':>10}{'cuda [ms]':>12}{'peak dev [MB]':>15}")
    for s in stats:
        cuda = f"{s.cuda_ms:.2f}" if s.cuda_ms is not None else "-"
        peak = f"{s.peak_device_mb:.1f}" if s.peak_device_mb is not None else "-"
        print(f"{s.name:<16}{s.wall_s:>10.3f}{s.cpu_s:>10.3f}{cuda:>12}{peak:>15}")


def main(argv: Optional[List[str]] = None) -> int:
    from . import flood_complex

    args = build_parser().parse_args(argv)
    if args.verbose:
        print(vars(args))
    device = validate_device(args.device)
    stats: List[StepStats] = []

    with StepTimer("Loading", device, args.cuda_events) as t:
        pc_cpu, n_pts, dim = load_point_cloud(Path(args.input_file))
    stats.append(t.stats)
    print(f"Loaded point cloud ({n_pts},{dim})")

    max_dim = dim if args.max_dimension is None else args.max_dimension
    points_per_edge, num_rand = resolve_simplex_representation(args.points_per_edge, args.num_rand)
    if num_rand is not None and args.seed is not None:
        np.random.seed(args.seed)
        torch.manual_seed(args.seed)

    with StepTimer("Flood complex", device, args.cuda_events) as t:
        pc = pc_cpu.to(device, non_blocking=True)
        st = flood_complex(pc, args.num_landmarks, max_dimension=max_dim, points_per_edge=points_per_edge,
                           batch_size=args.batch_size, fps_h=args.fps_height, use_triton=not args.no_triton,
                           return_simplex_tree=True, num_rand=num_rand)
    stats.append(t.stats)
    print(f"Built Flood complex with {st.num_simplices()} simplices")

    with StepTimer("Persistence", device, args.cuda_events) as t:
        st.compute_persistence()
        diagrams = [st.persistence_intervals_in_dimension(i) for i in range(max_dim)]
    stats.append(t.stats)
    print(f"Computed persistence up to max. dim {max_dim}\n")

    if args.output_file:
        meta = RunMeta(args.input_file, args.output_file, args.num_landmarks, max_dim, args.fps_height,
                       args.batch_size, str(device), points_per_edge, num_rand,
                       args.seed if num_rand is not None else None, not args.no_triton, n_pts, dim)
        save_output(Path(args.output_file), diagrams, meta)
    print_stats_table(stats)
    if args.stats_json:
        Path(args.stats_json).write_text(json.dumps([asdict(s) for s in stats], indent=2))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
