"""Two-GPU check of the sharded path (NCCL): skipped unless two CUDA devices are visible."""
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
        torch.manual_seed(1)
        rand_sharded = fb.flood_complex(pts, 100, points_per_edge=None, num_rand=64)
        os.environ["FLOODER_B200_NO_SHARD"] = "1"
        torch.manual_seed(1)
        rand_single = fb.flood_complex(pts, 100, points_per_edge=None, num_rand=64)
        assert rand_sharded == rand_single
        np.save(os.path.join(out_dir, f"n{rank}.npy"), np.array(sorted(sharded.values())))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_equals_single_gpu(tmp_path):
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "n0.npy"), np.load(tmp_path / "n1.npy")
    np.testing.assert_array_equal(a, b)
