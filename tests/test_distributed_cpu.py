"""World-size-2 gloo tests (CPU) of the sharding logic used for multi-GPU runs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flooder_b200 import distributed as fdist


def test_partition_is_a_balanced_permutation():
    g = torch.Generator().manual_seed(0)
    cost = torch.rand(1001, generator=g) ** 8          # heavy tail
    for world in (2, 3, 8):
        parts = fdist.partition(cost, world)
        allidx = torch.cat(parts).sort().values
        assert torch.equal(allidx, torch.arange(1001))
        sums = torch.stack([cost[p].sum() for p in parts])
        assert (sums.max() - sums.min()) / sums.mean() < 0.25
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    # deterministic under ties
    tie = torch.ones(10)
    assert all(torch.equal(a, b) for a, b in zip(fdist.partition(tie, 2), fdist.partition(tie, 2)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_rows, width, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shard = fdist.current_shard()
        assert shard is not None and shard.rank == rank and shard.world == world
        g = torch.Generator().manual_seed(123)               # same "landmarks" on every rank
        verts = torch.rand(n_rows, 4, 3, generator=g)
        calls = []

        def compute(v):                                       # stand-in for the CUDA pass
            calls.append(v.shape[0])
            return torch.stack([v.sum(dim=(1, 2)) * (j + 1) for j in range(width)], dim=1)

        full = fdist.sharded_covering_values(shard, verts, compute)
        want = compute(verts)
        assert torch.equal(full, want)
        assert calls[0] in (n_rows // world, n_rows // world + 1) or n_rows < world
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), full.numpy())
        # sharding can be switched off
        os.environ["FLOODER_B200_NO_SHARD"] = "1"
        assert fdist.current_shard() is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_rows", [1, 7, 250])
def test_sharded_values_gloo_world2(tmp_path, n_rows):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_rows, 15, str(tmp_path)), nprocs=world, join=True)
    a = np.load(tmp_path / "rank0.npy")
    b = np.load(tmp_path / "rank1.npy")
    np.testing.assert_array_equal(a, b)
    assert a.shape == (n_rows, 15)


def test_no_process_group_means_no_shard():
    assert fdist.current_shard() is None


def _consistency_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shard = fdist.current_shard()
        # rank 0's draw wins (random FPS start index, random weights)
        assert fdist.broadcast_int(shard, 100 + rank, "cpu") == 100
        w = torch.full((5, 3), float(rank))
        fdist.broadcast_tensor(shard, w)
        assert torch.equal(w, torch.zeros(5, 3))
        same = np.arange(30).reshape(10, 3)
        fdist.assert_same_on_all_ranks(shard, same, "cpu", "cells")          # identical: passes
        different = same + (rank == 1)
        try:
            fdist.assert_same_on_all_ranks(shard, different, "cpu", "cells")
            raised = False
        except RuntimeError as exc:
            raised = "differ between ranks" in str(exc)
        shorter = same[: 10 - rank]
        try:
            fdist.assert_same_on_all_ranks(shard, shorter, "cpu", "cells")
            raised_len = False
        except RuntimeError:
            raised_len = True
        np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([raised, raised_len]))
    finally:
        dist.destroy_process_group()


def test_rank_consistency_helpers_gloo_world2(tmp_path):
    """Every rank must derive the same landmarks / cells / weights: rank 0's random
    draws are broadcast, and differing simplex lists raise on EVERY rank instead of hanging or
    mis-scattering rows in the all-gather."""
    mp.spawn(_consistency_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for rank in range(2):
        assert np.load(tmp_path / f"ok{rank}.npy").all()
