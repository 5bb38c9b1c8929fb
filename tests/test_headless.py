"""The host C++ side above the C ABI: include/rlerc.hpp (the reference's RLE4 / RayMap / Map4 names, R/src/Rle4.h:25-52,
R/src/RayMap.h:57-418) and examples/headless.cpp (the frame loop of R/src/main.cpp without GL, BASELINE config 1:
one headless 1024x768 frame, PPM + warped-buffer dump for golden comparison)."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

from util import oracle_raymap, rgb_parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "rle-based-voxel-raycasting_b200")
EXE = os.path.join(ROOT, "examples", "headless")
POS, ROT = (10000.0, -818.0, 10000.0), (0.40, 0.30 + math.pi / 2, 0.0)        # main.cpp:316-320,344-347


@pytest.fixture(scope="module")
def host_check(tmp_path_factory):
    d = tmp_path_factory.mktemp("hpp")
    exe = str(d / "hpp_host_check")
    subprocess.run(["g++", "-std=gnu++11", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "hpp_host_check.cpp"), "-o", exe, "-L", PKG, "-lrlerc", "-Wl,-rpath," + PKG], check=True)
    r = subprocess.run([exe, str(d)], capture_output=True, text=True)
    return d, r


def test_hpp_mirror_scene_round_trip_and_ray_map(R, host_check):
    d, r = host_check
    assert r.returncode == 0 and r.stdout.startswith("ok"), (r.returncode, r.stdout, r.stderr)
    # the file the C++ RLE4 wrote loads through the Python mirror and equals the same scene built there
    a, b = R.RLE4.load(str(d / "scene.rle4")), R.RLE4.synth(0, 64, 64, 64, seed=1)
    assert a.nummaps == b.nummaps == int(r.stdout.split()[2])
    for m in range(a.nummaps):
        la, lb = a.level(m), b.level(m)
        assert la[:3] == lb[:3] and np.array_equal(la[3], lb[3]) and np.array_equal(la[4], lb[4])
    # RayMap::get_ray_map through the C++ class == through the Python mirror (itself pinned byte-exact to the
    # compiled reference, tests/test_oracle_vs_ref.py), apart from the nummaps field the caller filled in
    cfg = R.FrameConfig.default(1024, 768)
    rm = R.RayMap(cfg)
    rm.set_border(0.125)
    rm.set_ray_limit(4096)
    g = rm.get_ray_map(POS, ROT)
    g.nummaps = a.nummaps
    mine = bytes(C.string_at(C.byref(g), 896))
    theirs = open(str(d / "raymap.bin"), "rb").read()
    assert len(theirs) == 896 and theirs == mine
    assert int(r.stdout.split()[4]) == g.map_line_count


def test_headless_example_is_built_and_fails_loudly_without_a_gpu():
    assert os.access(EXE, os.X_OK), "examples/headless is built by csrc/Makefile (__graft_entry__.build)"
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" not in r.stdout
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the no-device error path is covered on the CPU box")
    r = subprocess.run([EXE, "--synth", "64", "--out", "/tmp/rlerc_headless_nogpu"], capture_output=True, text=True)
    assert r.returncode == 3 and "no CUDA device" in r.stderr, (r.returncode, r.stderr)      # no CPU fallback
    assert not os.path.exists("/tmp/rlerc_headless_nogpu.ppm")


@pytest.mark.gpu
@pytest.mark.parametrize("size", [(1024, 768), (640, 480)])
def test_headless_frame_against_the_oracle(R, rb, scene_mid, tmp_path, size):
    """BASELINE config 1 through the C++ host program: .rle4 on disk -> RLE4::load -> all_to_gpu -> get_ray_map ->
    traversal -> unwarp -> PPM; the dumps are compared with the oracle on the same scene and camera."""
    W, H = size
    path = str(tmp_path / "scene.rle4")
    scene_mid.save(path)
    pos = (POS[0], -40.0, POS[2])
    out = str(tmp_path / "f")
    r = subprocess.run([EXE, "--scene", path, "--size", str(W), str(H), "--pos"] + [repr(v) for v in pos] + ["--rot"] + [repr(v) for v in ROT]
                       + ["--out", out, "--frames", "8"], capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert "fly-through: 8 frames" in r.stdout
    info = dict(l.split() for l in open(out + ".txt").read().splitlines())
    cfg = R.FrameConfig.default(W, H)
    rm = R.RayMap(cfg).get_ray_map(pos, ROT)
    assert int(info["map_line_count"]) == rm.map_line_count
    lines = min(rm.map_line_count, cfg.rays_casted)
    warp = np.fromfile(out + ".warp.raw", dtype=np.uint32).reshape(lines, cfg.render_size)
    orm = oracle_raymap(rb, rm, scene_mid)
    owarp, _, cnt = rb.orc_render(orm, cfg.render_size, cfg.rays_casted, cfg.mip_distance, cfg.z_far)
    assert cnt["pixels"] > 10000
    assert np.array_equal(warp, owarp[:lines]), "warped ray buffer of the C++ host program differs from the oracle"
    with open(out + ".ppm", "rb") as f:
        assert f.readline() == b"P6\n" and f.readline() == b"%d %d\n" % (W, H) and f.readline() == b"255\n"
        rgb = np.frombuffer(f.read(), dtype=np.uint8).reshape(H, W, 3)
    orgba = rb.orc_unwarp(orm, W, H, cfg.render_size, cfg.rays_casted, cfg.rays_casted_res, owarp)
    dmax, same = rgb_parity(rgb, orgba[:, :, :3])
    assert dmax <= 1 and same >= 0.999, (dmax, same)         # north star: <= 1 LSB per channel, >= 99.9 % identical


@pytest.mark.gpu
def test_headless_multi_gpu_cpp_host(R, scene_mid, tmp_path):
    """The multi-GPU frame from a C++ host (rlerc_create_multi, include/rlerc.h): examples/headless --gpus N renders the
    same frame on every GPU of the box and exits non-zero unless it is byte-identical to its single-GPU frame."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    path = str(tmp_path / "scene.rle4")
    scene_mid.save(path)
    r = subprocess.run([EXE, "--scene", path, "--size", "1024", "768", "--pos", "10000", "-40", "10000", "--out", str(tmp_path / "m"),
                        "--frames", "64", "--gpus", str(n)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert "is byte-identical to the single-GPU frame" in r.stdout and "fly-through on %d GPUs" % n in r.stdout
