"""Multi-GPU sharding of the covering-radius pass (one process per GPU, torch.distributed).

The reference is single-process / single-GPU, so this is new: the value
of a simplex depends only on its own vertices and on the point cloud, hence the simplex list
shards with no data-path exchange.  Every rank holds the whole cloud on its own GPU, evaluates
its share of the simplices, and the per-simplex values (S x (2^K - 1) floats, a few hundred KB)
are all-gathered -- NCCL over NVLink when the process group is NCCL, gloo in the CPU tests.
Landmark FPS and the host Delaunay step are replicated (deterministic), not sharded.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Callable, List, Optional

import torch


@dataclass
class Shard:
    rank: int
    world: int
    group: Optional[object] = None


def current_shard() -> Optional[Shard]:
    """The active shard when torch.distributed is initialised with >1 ranks (and sharding is not
    disabled with FLOODER_B200_NO_SHARD=1), else None."""
    if os.environ.get("FLOODER_B200_NO_SHARD", "0") == "1":
        return None
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return None
    world = dist.get_world_size()
    if world <= 1:
        return None
    return Shard(rank=dist.get_rank(), world=world)


def partition(cost: torch.Tensor, world: int) -> List[torch.Tensor]:
    """Deal simplices to ranks in order of decreasing cost, in serpentine order
    (0..w-1, w-1..0, ...), so that the heavy tail (a few simplices whose ball swallows most of the
    cloud) is spread first and the per-rank cost sums stay within a fraction of the smallest
    items.  Every rank's share is itself sorted by decreasing cost (big work items first, small
    ones fill the tail).  Deterministic: ties are broken by index.
    Returns one int64 index tensor per rank (on ``cost.device``)."""
    n = int(cost.numel())
    order = torch.argsort(cost, descending=True, stable=True)
    pos = torch.arange(n, device=order.device)
    lap, col = pos // world, pos % world
    owner = torch.where(lap % 2 == 0, col, world - 1 - col)
    # share sizes are known on the host, so no device->host synchronisation is needed
    laps, rem = divmod(n, world)
    sizes = [laps] * world
    for c in range(rem):
        sizes[c if laps % 2 == 0 else world - 1 - c] += 1
    grouped = order[torch.argsort(owner, stable=True)]
    return [p.contiguous() for p in torch.split(grouped, sizes)]


def cost_proxy(simplex_vertices: torch.Tensor) -> torch.Tensor:
    """Longest edge of each simplex (the bounding-ball radius is proportional to it)."""
    v = simplex_vertices
    diff = v[:, :, None, :] - v[:, None, :, :]
    return diff.square().sum(dim=3).flatten(1).amax(dim=1)


def gather_rows(local: torch.Tensor, parts: List[torch.Tensor], shard: Shard) -> torch.Tensor:
    """All-gather per-rank row blocks (``local`` holds the rows ``parts[rank]``) into the full
    (S, W) tensor on every rank."""
    import torch.distributed as dist

    world = shard.world
    width = local.shape[1]
    rows_max = max(int(p.numel()) for p in parts)
    padded = torch.zeros((rows_max, width), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    gathered = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(gathered, padded, group=shard.group)
    total = sum(int(p.numel()) for p in parts)
    full = torch.empty((total, width), dtype=local.dtype, device=local.device)
    for r in range(world):
        full[parts[r].to(local.device)] = gathered[r][: parts[r].numel()]
    return full


def sharded_covering_values(
    shard: Shard,
    simplex_vertices: torch.Tensor,
    compute: Callable[[torch.Tensor], torch.Tensor],
    cost: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """Evaluate ``compute`` on this rank's share of the simplices and all-gather the rows.
    ``cost`` (one value per simplex, identical on every rank) drives the partition; without it
    the longest edge is used."""
    parts = partition(cost_proxy(simplex_vertices) if cost is None else cost, shard.world)
    mine = parts[shard.rank]
    local = compute(simplex_vertices[mine].contiguous())
    return gather_rows(local, parts, shard)


def broadcast_int(shard: Shard, value: int, device) -> int:
    """Rank 0's ``value`` on every rank (used for the random FPS start index)."""
    import torch.distributed as dist

    dev = device if dist.get_backend(shard.group) == "nccl" else "cpu"
    t = torch.tensor([int(value)], dtype=torch.int64, device=dev)
    dist.broadcast(t, src=0, group=shard.group)
    return int(t.item())


def broadcast_tensor(shard: Shard, tensor: torch.Tensor) -> torch.Tensor:
    """In place: rank 0's contents on every rank."""
    import torch.distributed as dist

    if dist.get_backend(shard.group) != "nccl" and tensor.device.type != "cpu":
        host = tensor.cpu()
        dist.broadcast(host, src=0, group=shard.group)
        tensor.copy_(host)
    else:
        dist.broadcast(tensor, src=0, group=shard.group)
    return tensor


def assert_same_on_all_ranks(shard: Shard, array, device, what: str) -> None:
    """Raise on every rank if ``array`` (numpy, integer) differs between ranks: the sharded pass
    all-gathers rows by position, so all ranks must hold the same simplex list."""
    import numpy as np
    import torch.distributed as dist

    a = np.ascontiguousarray(array, dtype=np.int64).reshape(-1)
    # position-weighted checksum + length; int64 wrap-around is fine for an equality check
    with np.errstate(over="ignore"):
        digest = int((a * (np.arange(a.size, dtype=np.int64) * 2654435761 + 1)).sum()) & 0x3FFFFFFFFFFFFFFF
    dev = device if dist.get_backend(shard.group) == "nccl" else "cpu"
    t = torch.tensor([digest, -digest, a.size, -a.size], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=shard.group)
    lo_hi = t.tolist()
    if lo_hi[0] != -lo_hi[1] or lo_hi[2] != -lo_hi[3]:
        raise RuntimeError(f"flooder_b200: {what} differ between ranks; the sharded pass needs identical inputs "
                           "on every rank (same points, same landmarks or the same integer landmark count)")
