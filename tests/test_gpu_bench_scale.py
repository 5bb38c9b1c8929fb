"""Parity AT BENCH SCALE: the very scenes, windows and fly-through frames bench.py times (BASELINE configs 1-5), rendered by
the production path through the C ABI and by the reference's own render_line compiled for the host (oracle/_ref, OpenMP
over ray planes; the oracle port when oracle/_ref is absent): warped ray buffer bit-exact over every ray plane of the
frame, final RGBA <= 1 LSB per channel and >= 99.9 % of the pixels identical (north star)."""
import math
import os
import subprocess
import sys

import numpy as np
import pytest

from util import oracle_raymap, rgb_parity

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    import bench
    return bench


@pytest.fixture(scope="module")
def gpu(R):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box")
    r = R.Renderer(0)
    yield r
    r.close()


def _cpu_frame(R, rb, scene, cfg, rm):
    orm = oracle_raymap(rb, rm, scene)
    threads = os.cpu_count() or 1
    if rb.have_ref():
        warp, _ = rb.ref_render_frame(orm, cfg.render_size, mip_distance=cfg.mip_distance, z_far=cfg.z_far, rays=cfg.rays_casted, threads=threads)
    else:
        warp, _, _ = rb.orc_render(orm, cfg.render_size, cfg.rays_casted, cfg.mip_distance, cfg.z_far, threads=threads)
    return orm, warp


def _check_frames(R, rb, gpu, workload, frames, K, lanes=(0,), rows=None, rgba=True):
    """frames: step indices of the K-step fly-through bench.py runs for this workload.  rows = (chunks, size): compare
    `chunks` runs of `size` consecutive ray planes spread over the frame only."""
    import torch
    bench = _bench()
    R.lib().rlerc_set_host_threads(os.cpu_count() or 1)
    scene, name, sy = bench.build_scene(R, workload, lambda m: None)
    W, H = bench.WORKLOADS[workload][3]
    cfg = R.FrameConfig.default(W, H)
    gpu.all_to_gpu(scene)
    host = torch.zeros((H, W, 4), dtype=torch.uint8).pin_memory()
    for i in frames:
        pos, rot = bench.path_pose(R, i, K, sy, name == "Imrodh.rle4")
        for ln in lanes:
            gpu.set_lanes_per_ray(ln)
            rm = gpu.render_frame(pos, rot, cfg, host.numpy())           # the call a host makes: pose in, pixels out
            n = min(rm.map_line_count, cfg.rays_casted)
            got = gpu.read_warp(cfg)
            if ln == lanes[0]:
                if rows is None:
                    orm, want = _cpu_frame(R, rb, scene, cfg, rm)
                else:
                    orm = oracle_raymap(rb, rm, scene)
                    want = None
            if rows is None:
                assert np.array_equal(got[:n], want[:n]), (workload, i, ln, int((got[:n] != want[:n]).sum()))
            else:
                chunks, size = rows
                for x in np.unique(np.linspace(0, n - size, chunks).astype(int)):
                    w, _, _ = rb.orc_render(orm, cfg.render_size, cfg.rays_casted, cfg.mip_distance, cfg.z_far, ray_begin=int(x), ray_end=int(x) + size,
                                            threads=os.cpu_count() or 1)
                    assert np.array_equal(got[x:x + size], w[x:x + size]), (workload, i, ln, int(x))
                    del w
            if rgba and rows is None:
                want_rgba = rb.orc_unwarp(orm, W, H, cfg.render_size, cfg.rays_casted, cfg.rays_casted_res, want)
                mx, same = rgb_parity(host.numpy(), want_rgba)
                assert mx <= 1 and same >= 0.999, (workload, i, ln, mx, same)      # tolerance stated by the north star
    gpu.set_lanes_per_ray(0)
    del scene


def test_config2_imrodh1080p_all_driver_frames(R, rb, gpu):
    """BASELINE config 2 (1024^3 scene, 1920x1080): ALL 20 frames of the driver's `--steps 20` run."""
    _check_frames(R, rb, gpu, "imrodh1080p", range(20), 20)


def test_config2_imrodh1080p_both_production_kernels(R, rb, gpu):
    _check_frames(R, rb, gpu, "imrodh1080p", (0, 7, 13), 20, lanes=(65, 68, 69), rgba=False)


def test_config3_tiled4k(R, rb, gpu):
    """BASELINE config 3 (16x16 physical tiling, 3.8 GB of RLE, 3840x2160): frames of the north-star measurement."""
    _check_frames(R, rb, gpu, "tiled4k", (0, 9, 17), 24)


def test_config4_shortrun16k(R, rb, gpu):
    """BASELINE config 4 at full size (16384 x 1024 x 16384, 7.9 GB of RLE, worst-case short-run band)."""
    _check_frames(R, rb, gpu, "shortrun16k", (3,), 20)


def test_config5_view8k(R, rb, gpu):
    """BASELINE config 5 window (7680x4320 view of the tiled scene): 8 x 16 ray planes spread over the frame against the
    oracle port (its buffers for a whole 8K frame are ~1 GB per call); the texel selection of the unwarp is covered at
    8K by test_gpu_parity.test_unwarp_texel_selection_equals_oracle."""
    _check_frames(R, rb, gpu, "view8k", (5,), 20, rows=(8, 16), rgba=False)


def test_config1_headless_cpp_host(R, rb):
    """BASELINE config 1 at full size through the C++ host program (examples/headless: RLE4::load -> all_to_gpu ->
    get_ray_map -> traversal -> unwarp -> PPM) against the reference compiled for the host: tools/config1_golden.py
    asserts an identical warped buffer and <= 1 LSB RGB."""
    out = os.path.join(ROOT, "gpurun_out", "config1_test")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "config1_golden.py"), out], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, (p.stdout[-1500:], p.stderr[-1500:])
    assert '"warped_buffer_identical": true' in p.stdout
