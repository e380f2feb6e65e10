"""Two-GPU test of the real multi-GPU path: sample-range partition, render on each device through the C ABI,
one NCCL reduce of the accumulator, compared with the single-GPU render of the same samples.  Skipped on
boxes with fewer than two GPUs (the gloo version in test_cpu_distributed.py always runs)."""
import importlib
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_path):
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    rtb = importlib.import_module("ray-tracing-v06_b200")
    scene = rtb.Scene.named("book2_final"); cam = scene.info.camera
    W, H, SPP, D = 96, 96, 12, 20
    r = rtb.Renderer(rank); r.set_scene(scene); r.set_camera(cam)
    s0, s1 = rtb.sample_range(SPP, rank, world)
    r.render(W, H, s0, s1, D, seed=1984, stream=torch.cuda.current_stream().cuda_stream)
    acc = r.accum_tensor()
    dist.reduce(acc, dst=0)
    torch.cuda.synchronize()
    if rank == 0:
        np.save(out_path, acc.cpu().numpy())
        r.render(W, H, 0, SPP, D, seed=1984); np.save(out_path + ".single.npy", r.download_accum())
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpus_sum_to_the_single_gpu_image(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    out = str(tmp_path / "sum.npy")
    mp.start_processes(_worker, args=(2, 29651, out), nprocs=2, join=True, start_method="spawn")
    got, single = np.load(out), np.load(out + ".single.npy")
    assert np.array_equal(got[..., 3], single[..., 3])
    np.testing.assert_allclose(got, single, rtol=1e-5, atol=1e-5)
