"""world_size-2 (gloo) test of the multi-GPU plumbing: sample-range partition + one framebuffer sum.
The per-rank render is stood in for by the CPU oracle (this is a test: the product has no CPU path)."""
import importlib
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, out_path):
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rtb = importlib.import_module("ray-tracing-v06_b200"); orc = importlib.import_module("pyoracle")
    scene = rtb.Scene.named("book2_checker"); cam = scene.info.camera
    spp, W, H, D = 10, 48, 27, 8
    s0, s1 = rtb.sample_range(spp, rank, world)
    acc, _, rays = orc.OracleScene(scene.serialize()).render(cam, W, H, s0, s1, D, seed=1984, threads=1)
    t = torch.from_numpy(acc)
    dist.reduce(t, dst=0)                                      # the only collective of the path
    r = torch.tensor([rays], dtype=torch.int64); dist.all_reduce(r)
    if rank == 0:
        np.save(out_path, t.numpy()); np.save(out_path + ".rays.npy", r.numpy())
    dist.destroy_process_group()


def test_sample_range_partition(rtb):
    for spp in (1, 7, 10, 10000):
        for world in (1, 2, 3, 8):
            ranges = [rtb.sample_range(spp, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == spp
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
    rows = [rtb.row_range(675, r, 4) for r in range(4)]
    assert rows[0][0] == 0 and rows[-1][1] == 675 and all(rows[i][1] == rows[i + 1][0] for i in range(3))


def test_two_ranks_sum_to_the_single_rank_image(rtb, orc, tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "sum.npy")
    mp.start_processes(_worker, args=(2, 29641, out), nprocs=2, join=True, start_method="spawn")
    got = np.load(out); rays = int(np.load(out + ".rays.npy")[0])
    scene = rtb.Scene.named("book2_checker")
    whole, _, wrays = orc.OracleScene(scene.serialize()).render(scene.info.camera, 48, 27, 0, 10, 8, seed=1984, threads=1)
    assert np.array_equal(got[..., 3], whole[..., 3]) and rays == wrays
    np.testing.assert_allclose(got, whole, rtol=1e-6, atol=1e-6)   # counter-based RNG: same paths, different summation grouping
