"""Command line front-end (reference: flooder/cli.py)."""
import json
import pickle

import numpy as np
import pytest
import torch

from flooder_b200 import cli


def test_parser_defaults_match_reference():
    args = cli.build_parser().parse_args(["--input-file", "x.npy"])
    assert (args.num_landmarks, args.fps_height, args.batch_size, args.device) == (2000, 9, 64, "cuda:0")
    assert args.points_per_edge is None and args.num_rand is None and args.max_dimension is None
    assert cli.resolve_simplex_representation(None, None) == (30, None)
    assert cli.resolve_simplex_representation(None, 100) == (None, 100)
    with pytest.raises(SystemExit):
        cli.build_parser().parse_args(["--input-file", "x.npy", "--points-per-edge", "5", "--num-rand", "9"])
    with pytest.raises(SystemExit):
        cli.build_parser().parse_args([])          # --input-file is required


def test_cpu_device_is_rejected(tmp_path):
    np.save(tmp_path / "c.npy", np.random.rand(50, 3).astype(np.float32))
    with pytest.raises(SystemExit, match="not supported"):
        cli.main(["--input-file", str(tmp_path / "c.npy"), "--device", "cpu"])


def test_save_output_roundtrip(tmp_path):
    meta = cli.RunMeta("in.npy", "out", 10, 2, 9, 64, "cuda:0", 30, None, None, True, 100, 2)
    path = cli.save_output(tmp_path / "sub" / "out", [np.zeros((2, 2))], meta)
    assert path.suffix == ".pkl"
    payload = pickle.loads(path.read_bytes())
    assert payload["meta"]["num_landmarks"] == 10 and len(payload["diagrams"]) == 1


@pytest.mark.gpu
def test_cli_end_to_end(tmp_path):
    import flooder_b200 as fb

    torch.manual_seed(0)
    np.save(tmp_path / "torus.npy", fb.generate_noisy_torus_points_3d(20_000).numpy())
    rc = cli.main(["--input-file", str(tmp_path / "torus.npy"), "--num-landmarks", "80", "--points-per-edge", "10",
                   "--output-file", str(tmp_path / "out.pkl"), "--stats-json", str(tmp_path / "stats.json"),
                   "--cuda-events"])
    assert rc == 0
    payload = pickle.loads((tmp_path / "out.pkl").read_bytes())
    assert len(payload["diagrams"]) == 3 and payload["meta"]["n_points"] == 20_000
    h0 = payload["diagrams"][0]
    assert np.isinf(h0[:, 1]).sum() == 1                      # one connected component
    h1 = payload["diagrams"][1]
    assert (np.sort(h1[:, 1] - h1[:, 0])[-2:] > 0.5).all()    # the two generating loops of the torus
    stats = json.loads((tmp_path / "stats.json").read_text())
    assert [s["name"] for s in stats] == ["Loading", "Flood complex", "Persistence"]
