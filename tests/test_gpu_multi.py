"""Multi-GPU check of the sharded path (NCCL, one process per GPU): skipped unless at least two CUDA
devices are visible; uses every visible device (2 under `gpurun --gpus 2`, 8 under `--gpus 8`)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist

    import flooder_b200 as fb

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        torch.manual_seed(42)
        np.random.seed(42)
        pts = fb.generate_noisy_torus_points_3d(60_000).cuda()
        sharded = fb.flood_complex(pts, 200, points_per_edge=12)
        os.environ["FLOODER_B200_NO_SHARD"] = "1"
        single = fb.flood_complex(pts, 200, points_per_edge=12)
        del os.environ["FLOODER_B200_NO_SHARD"]
        assert sharded == single, "sharded result differs from the single-GPU result"

        # random mode: every rank has its OWN CPU generator state; rank 0's weights are used everywhere
        torch.manual_seed(1)
        expect = None
        if rank == 0:
            os.environ["FLOODER_B200_NO_SHARD"] = "1"
            expect = fb.flood_complex(pts, 100, points_per_edge=None, num_rand=64)
            del os.environ["FLOODER_B200_NO_SHARD"]
            torch.manual_seed(1)
        else:
            torch.manual_seed(1000 + rank)
        rand_sharded = fb.flood_complex(pts, 100, points_per_edge=None, num_rand=64)
        if rank == 0:
            assert rand_sharded == expect, "sharded random mode differs from rank 0's single-GPU result"

        # random FPS start index: drawn by rank 0, shared; all ranks must return the same complex
        np.random.seed(7 + rank)
        any_start = fb.flood_complex(pts, 80, points_per_edge=8, start_idx=None)

        # a larger job: more simplices than ranks x warps, heavy-tailed costs (same cloud on every rank)
        torch.manual_seed(43)
        np.random.seed(43)
        big = fb.generate_noisy_torus_points_3d(300_000).cuda()
        big_sharded = fb.flood_complex(big, 400, points_per_edge=20)
        os.environ["FLOODER_B200_NO_SHARD"] = "1"
        big_single = fb.flood_complex(big, 400, points_per_edge=20)
        del os.environ["FLOODER_B200_NO_SHARD"]
        assert big_sharded == big_single

        def flat(d):
            keys = sorted(d)
            return np.array([d[k] for k in keys]), np.array([hash(k) & 0xFFFFFFFF for k in keys])

        np.savez(os.path.join(out_dir, f"n{rank}.npz"), grid=flat(sharded)[0], rand=flat(rand_sharded)[0],
                 start_vals=flat(any_start)[0], start_keys=flat(any_start)[1], big=flat(big_sharded)[0])

        # ranks that disagree about the landmarks must fail loudly, not hang or mis-scatter
        lms = pts[:50].clone()
        if rank == 1:
            lms[0, 0] += 0.25
        try:
            fb.flood_complex(pts, lms, points_per_edge=6)
            raised = False
        except RuntimeError as exc:
            raised = "differ between ranks" in str(exc)
        assert raised, "differing landmarks were not detected"
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least two GPUs")
def test_sharded_equals_single_gpu(tmp_path):
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 8)
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    first = np.load(tmp_path / "n0.npz")
    for rank in range(1, world):
        other = np.load(tmp_path / f"n{rank}.npz")
        for key in first.files:
            np.testing.assert_array_equal(first[key], other[key], err_msg=f"rank {rank}: {key}")
    print(f"sharded == single GPU on {world} GPUs: {len(first['grid'])} + {len(first['big'])} simplices")
