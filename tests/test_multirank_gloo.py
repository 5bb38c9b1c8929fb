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


class _OracleRenderer:
    """Stand-in for Renderer in the CPU test of FrameFarm: `render` + `unwarp` produce the oracle's frame for the ray
    map and copy it to the destination pointer (the product renderer launches the CUDA kernels there)."""
    device = 0

    def __init__(self, rb, scene, cfg):
        self.rb, self.scene, self.cfg, self.frames = rb, scene, cfg, 0

    def set_stream(self, s):
        pass

    def render(self, rm, cfg):
        from util import oracle_raymap
        self.orm = oracle_raymap(self.rb, rm, self.scene)
        self.warp, _, _ = self.rb.orc_render(self.orm, cfg.render_size, cfg.rays_casted, cfg.mip_distance, cfg.z_far)

    def unwarp(self, rm, cfg, d_rgba=None):
        import ctypes as C
        img = self.rb.orc_unwarp(self.orm, cfg.width, cfg.height, cfg.render_size, cfg.rays_casted, cfg.rays_casted_res, self.warp)
        C.memmove(d_rgba, img.ctypes.data, img.nbytes)
        self.frames += 1


def _farm_worker(rank, world, port, out_path, nframes):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    R = importlib.import_module("rle-based-voxel-raycasting_b200")
    MG = importlib.import_module("rle-based-voxel-raycasting_b200.multigpu")
    from oracle import refbind as rb
    scene = R.RLE4.synth(0, 64, 64, 64, seed=1)
    cfg = R.FrameConfig.default(128, 96)
    poses = [((10000.0 + 3.0 * i, -40.0, 10000.0), (0.4, 1.9 + 0.2 * i, 0.0)) for i in range(nframes)]
    rend = _OracleRenderer(rb, scene, cfg)
    farm = MG.FrameFarm(rend, cfg, torch, rank, world, dist, device=torch.device("cpu"))
    rounds = (nframes + world - 1) // world
    got = {}
    for rd in range(rounds + 1):                                 # the loop of bench.py: gather of round rd-1 overlaps round rd
        if rd < rounds:
            i = rd * world + rank
            farm.render_round(rd, R.RayMap(cfg).get_ray_map(*poses[i]) if i < nframes else None)
        if rd > 0 and rank == 0:
            done = farm.wait_round(rd - 1)
            for j in range(world):
                if (rd - 1) * world + j < nframes:
                    got[(rd - 1) * world + j] = done[j].numpy().copy()
    farm.finish()
    dist.barrier()
    assert rend.frames == len(range(rank, nframes, world))      # every rank rendered exactly its own frames
    if rank == 0:
        solo = _OracleRenderer(rb, scene, cfg)
        ok = len(got) == nframes
        for i in range(nframes):
            buf = np.zeros((cfg.height, cfg.width, 4), np.uint8)
            solo.render(R.RayMap(cfg).get_ray_map(*poses[i]), cfg)
            solo.unwarp(None, cfg, d_rgba=buf.ctypes.data)
            ok = ok and np.array_equal(got[i], buf)
        open(out_path, "w").write("%d" % ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("nframes", [4, 5])
def test_frame_farm_deals_and_gathers_whole_frames(tmp_path, nframes):
    """Frames mode (the default for N > 1): frames dealt round-robin, double-buffered gather on rank 0, in order, also
    when the last round is incomplete."""
    out = str(tmp_path / "farm.txt")
    port = 31500 + (os.getpid() % 2000) + nframes
    mp.spawn(_farm_worker, args=(2, port, out, nframes), nprocs=2, join=True)
    assert open(out).read() == "1"


def _pull_worker(rank, world, port, out_path, block):
    """The round-2 compositor (csrc/group.cu) with the oracle standing in for the kernels and gloo for NVLink: every rank
    traverses its interleaved blocks of ray planes into ITS warped buffer, then produces its band of window rows, taking
    every texel from the buffer of the rank that owns the texel's ray plane."""
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
    cfg = R.FrameConfig.default(256, 190)                       # a height the band size does not divide
    pos, rot = few_cameras(-40.0)[2]
    rm = R.RayMap(cfg).get_ray_map(pos, rot)
    orm = oracle_raymap(rb, rm, scene)
    count = min(rm.map_line_count, cfg.rays_casted)
    # 1. traversal of the owned ray planes only (rows of the others stay zero in this rank's buffer)
    mine = np.zeros((cfg.rays_casted, cfg.render_size), np.uint32)
    own = np.array(MG.owned_mask(count, block, world, rank), bool)
    b = 0
    while b < count:
        if own[b]:
            e = b
            while e < count and own[e]:
                e += 1
            w, _, _ = rb.orc_render(orm, cfg.render_size, cfg.rays_casted, cfg.mip_distance, cfg.z_far, ray_begin=b, ray_end=e)
            mine[b:e] = w[b:e]
            b = e
        else:
            b += 1
    # 2. "peer memory": every rank can read every rank's buffer
    bufs = [torch.zeros(mine.shape, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(bufs, torch.from_numpy(mine.view(np.int32).copy()))
    bufs = [t.numpy().view(np.uint32) for t in bufs]
    # 3. this rank's band of rows: texel (iy, ix) of every pixel from the owner of ray plane iy
    rows = MG.band_rows(cfg.height, world)
    r0, r1 = min(cfg.height, rank * rows), min(cfg.height, (rank + 1) * rows)
    _, tex = rb.orc_unwarp(orm, cfg.width, cfg.height, cfg.render_size, cfg.rays_casted, cfg.rays_casted_res, mine, want_texels=True)
    iy, ix = tex[r0:r1, :, 0], tex[r0:r1, :, 1]
    owner = (iy // block) % world
    pulled = np.zeros((cfg.rays_casted, cfg.render_size), np.uint32)          # the texels this band samples, each from its owner
    for p in range(world):
        sel = owner == p
        pulled[iy[sel], ix[sel]] = bufs[p][iy[sel], ix[sel]]
    band = rb.orc_unwarp(orm, cfg.width, cfg.height, cfg.render_size, cfg.rays_casted, cfg.rays_casted_res, pulled)[r0:r1]
    # 4. the bands assembled on rank 0 (the product pushes them there over NVLink)
    padded = np.zeros((rows, cfg.width, 4), np.uint8)
    padded[:r1 - r0] = band
    parts = [torch.zeros(padded.shape, dtype=torch.uint8) for _ in range(world)] if rank == 0 else None
    dist.gather(torch.from_numpy(padded), parts, dst=0)
    if rank == 0:
        got = np.concatenate([t.numpy() for t in parts], axis=0)[:cfg.height]
        full_warp, _, _ = rb.orc_render(orm, cfg.render_size, cfg.rays_casted, cfg.mip_distance, cfg.z_far)
        full = rb.orc_unwarp(orm, cfg.width, cfg.height, cfg.render_size, cfg.rays_casted, cfg.rays_casted_res, full_warp)
        open(out_path, "w").write("%d %d" % (bool(np.array_equal(got, full)), int((owner != rank).sum())))
    dist.destroy_process_group()


@pytest.mark.parametrize("block", [32, 7])
def test_two_rank_pull_compositor_equals_single(tmp_path, block):
    """csrc/group.cu's scheme on CPU: slices traversed per rank, bands unwarped per rank with texels pulled from the owning
    rank, bands assembled on rank 0 == the single-process frame, byte for byte."""
    out = str(tmp_path / "pull.txt")
    port = 33500 + (os.getpid() % 2000) + block
    mp.spawn(_pull_worker, args=(2, port, out, block), nprocs=2, join=True)
    ok, remote = (int(v) for v in open(out).read().split())
    assert ok == 1
    assert remote > 0              # rank 0's band really needed texels of the other rank
