"""N > 1 host logic on CPU: world_size-2 gloo.  Each rank produces the pixels of its ray-plane slice
(with the oracle standing in for the kernels), the product's compositing collective sums them on
rank 0, and the result must equal the single-process frame bit for bit."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    R = importlib.import_module("rle-based-voxel-raycasting_b200")
    MG = importlib.import_module("rle-based-voxel-raycasting_b200.multigpu")
    from oracle import refbind as rb
    from util import few_cameras, oracle_raymap
    scene = R.RLE4.synth(0, 64, 64, 64, seed=1)
    cfg = R.FrameConfig.default(256, 192)
    pos, rot = few_cameras(-40.0)[0]
    rm = R.RayMap(cfg).get_ray_map(pos, rot)
    orm = oracle_raymap(rb, rm, scene)
    count = rm.map_line_count
    # contiguous slices = interleaved slices with one block per rank
    block = (count + world - 1) // world
    b, e = rank * block, min(count, (rank + 1) * block)
    assert MG.owned_count(count, block, world, rank) == e - b
    warp, _, _ = rb.orc_render(orm, cfg.render_size, cfg.rays_casted, cfg.mip_distance, cfg.z_far, ray_begin=b, ray_end=e)
    part = rb.orc_unwarp(orm, cfg.width, cfg.height, cfg.render_size, cfg.rays_casted, cfg.rays_casted_res, warp, ray_begin=b, ray_end=e)
    img = MG.composite(torch.from_numpy(part.copy()), dist)
    dist.barrier()
    if rank == 0:
        full_warp, _, _ = rb.orc_render(orm, cfg.render_size, cfg.rays_casted, cfg.mip_distance, cfg.z_far)
        full = rb.orc_unwarp(orm, cfg.width, cfg.height, cfg.render_size, cfg.rays_casted, cfg.rays_casted_res, full_warp)
        ok = bool(np.array_equal(img.numpy(), full))
        nz = int((part.reshape(-1, 4).max(axis=1) > 0).sum())
        open(out_path, "w").write("%d %d %d" % (ok, nz, full.shape[0] * full.shape[1]))
    dist.destroy_process_group()


def test_two_rank_composite_equals_single(tmp_path):
    out = str(tmp_path / "result.txt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    ok, nz, total = (int(v) for v in open(out).read().split())
    assert ok == 1
    assert 0 < nz < total          # rank 0 really only had part of the frame


def test_interleaved_partition_arithmetic():
    MG = importlib.import_module("rle-based-voxel-raycasting_b200.multigpu")
    for count in (0, 1, 31, 32, 33, 1000, 7680):
        for block in (1, 32, 128):
            for n in (1, 2, 3, 8):
                masks = [MG.owned_mask(count, block, n, r) for r in range(n)]
                # a partition: every ray plane has exactly one owner
                assert all(sum(m[i] for m in masks) == 1 for i in range(count))
                for r in range(n):
                    assert MG.owned_count(count, block, n, r) == sum(masks[r])
