"""The parity tests proper: the CUDA path, called through the C ABI, against the oracle on the same
seeded inputs.  Bit-exact for the warped ray buffer, hit identity (column, mip, voxel index) and work
counters; <= 1 LSB per channel and >= 99.9 % identical pixels for the final RGBA (north star)."""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest

from util import camera_grid, edge_cameras, edge_scenes, few_cameras, oracle_raymap, rgb_parity, sha

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gpu(R):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box")
    r = R.Renderer(0)
    yield r
    r.close()


def _variants(R):
    """The measured-and-rejected kernels of round 1 (lanes 64, 66, 67) are only in a library built with VARIANTS=1."""
    return bool(R.lib().rlerc_has_variants())


def _lanes_list(R, lanes):
    return [l for l in lanes if l not in (64, 66, 67) or _variants(R)]


def _fresh_warp(r, cfg):
    wp = r.warp_buffer(cfg)
    r.upload(wp, np.zeros((cfg.rays_casted, cfg.render_size), np.uint32))
    return wp


def _oracle(rb, rm, scene, cfg, ids=False):
    orm = oracle_raymap(rb, rm, scene)
    warp, oid, cnt = rb.orc_render(orm, cfg.render_size, cfg.rays_casted, cfg.mip_distance, cfg.z_far, want_ids=ids)
    return orm, warp, oid, cnt


@pytest.mark.parametrize("lanes", [0, 69, 68, 66, 67, 65, 64, 32, 8, 1])
def test_warp_buffer_bit_exact_camera_grid(R, rb, gpu, scene_mid, lanes):
    if lanes in (64, 66, 67) and not _variants(R):
        pytest.skip("variant kernels are not in this build (make VARIANTS=1)")
    gpu.all_to_gpu(scene_mid)
    gpu.set_lanes_per_ray(lanes)
    cfg = R.FrameConfig.default(640, 480)
    cams = list(camera_grid(-100.0))
    if lanes < 32:
        cams = cams[::3]
    else:
        # exactly on a lattice point looking along a grid axis (NaN / inf DDA tracks: merge path -> serial
        # fallback), and negative coordinates (first intersections opposite in sign to their gradients)
        cams = cams + [((10000.0, -100.0, 10000.0), (0.3, math.pi / 2, 0.0)), ((10000.0, -100.0, 10000.0), (0.3, 0.0, 0.0)),
                       ((-3000.25, -100.0, -70.5), (0.3, 0.7, 0.0)), ((-0.5, -60.0, 2000.75), (0.5, 3.9, 0.0))]
    for pos, rot in cams:
        rm = R.RayMap(cfg).get_ray_map(pos, rot)
        _, want, _, _ = _oracle(rb, rm, scene_mid, cfg)
        _fresh_warp(gpu, cfg)
        gpu.render(rm, cfg)
        got = gpu.read_warp(cfg)
        assert np.array_equal(got, want), (lanes, rot, int((got != want).sum()))
    gpu.set_lanes_per_ray(0)


@pytest.mark.parametrize("variant", ["closed", "serial", "merge", "producer"])
def test_dda_variants_bit_exact(R, rb, gpu, scene_mid, variant):
    """k_traverse_w can advance the DDA four ways (closed form evaluated lane-parallel, serial recurrence in
    every warp, merge path over the two tracks, dedicated producer blocks + ring in global memory): same warped
    buffer, bit for bit."""
    if not _variants(R):
        pytest.skip("variant kernels are not in this build (make VARIANTS=1)")
    gpu.all_to_gpu(scene_mid)
    gpu.set_lanes_per_ray(64)
    gpu.set_dda_mode({"merge": 2, "closed": 3}.get(variant, 0))
    gpu.set_dda_producer(variant == "producer")
    try:
        for wh in ((640, 480), (1920, 1080)):
            cfg = R.FrameConfig.default(*wh)
            cams = list(camera_grid(-100.0))[::2] if wh[0] == 640 else few_cameras(-100.0)
            # a camera exactly on a lattice point looking along a grid axis: NaN/inf tracks (merge path falls back)
            cams = cams + [((10000.0, -100.0, 10000.0), (0.3, math.pi / 2, 0.0)), ((10000.0, -100.0, 10000.0), (0.3, 0.0, 0.0))]
            # negative coordinates: the first intersections have the opposite sign of their gradients
            cams = cams + [((-3000.25, -100.0, -70.5), (0.3, 0.7, 0.0)), ((-0.5, -60.0, 2000.75), (0.5, 3.9, 0.0))]
            for pos, rot in cams:
                rm = R.RayMap(cfg).get_ray_map(pos, rot)
                _, want, _, _ = _oracle(rb, rm, scene_mid, cfg)
                _fresh_warp(gpu, cfg)
                gpu.render(rm, cfg)
                got = gpu.read_warp(cfg)
                assert np.array_equal(got, want), (variant, wh, rot, int((got != want).sum()))
    finally:
        gpu.set_lanes_per_ray(0)
        gpu.set_dda_mode(0)
        gpu.set_dda_producer(False)


@pytest.mark.parametrize("lanes", [0, 32])
def test_warp_buffer_bit_exact_short_run_scene(R, rb, gpu, scene_runs, scene_rle, lanes):
    """Columns with dozens of runs: the lane<->run path.  Both worst-case scenes: the compressed bit volume and the
    one written straight into RLE (the small sibling of BASELINE config 4's full-size scene)."""
    gpu.set_lanes_per_ray(lanes)
    cfg = R.FrameConfig.default(512, 384)
    for scene, h in ((scene_runs, -90.0), (scene_rle, -30.0)):
        gpu.all_to_gpu(scene)
        for pos, rot in few_cameras(h):
            rm = R.RayMap(cfg).get_ray_map(pos, rot)
            _, want, _, _ = _oracle(rb, rm, scene, cfg)
            _fresh_warp(gpu, cfg)
            gpu.render(rm, cfg)
            assert np.array_equal(gpu.read_warp(cfg), want), (lanes, h, rot)
    gpu.set_lanes_per_ray(0)


def test_against_compiled_reference(R, rb, gpu, have_ref, scene_mid):
    """Same check against the reference's own render_line compiled for the host (oracle/_ref)."""
    if not have_ref:
        pytest.skip("oracle/_ref was not shipped with this snapshot")
    gpu.all_to_gpu(scene_mid)
    cfg = R.FrameConfig.default(1024, 768)
    for pos, rot in few_cameras(-100.0):
        rm = R.RayMap(cfg).get_ray_map(pos, rot)
        orm = oracle_raymap(rb, rm, scene_mid)
        want, _ = rb.ref_render_frame(orm, cfg.render_size, mip_distance=cfg.mip_distance, z_far=cfg.z_far, rays=cfg.rays_casted)
        _fresh_warp(gpu, cfg)
        gpu.render(rm, cfg)
        assert np.array_equal(gpu.read_warp(cfg), want), rot


def test_golden_hashes(R, gpu):
    gold = json.load(open(os.path.join(HERE, "golden", "golden.json")))
    for name, case in gold["cases"].items():
        n = case["size"]
        scene = R.RLE4.synth(case["kind"], n, n, n, seed=case["seed"])
        gpu.all_to_gpu(scene)
        cfg = R.FrameConfig.default(*case["window"])
        for fr in case["frames"]:
            rm = R.RayMap(cfg).get_ray_map(fr["pos"], fr["rot"])
            _fresh_warp(gpu, cfg)
            gpu.render(rm, cfg)
            assert sha(gpu.read_warp(cfg)) == fr["warp_sha"], (name, fr["rot"])


@pytest.mark.parametrize("lanes", [0, 32])
def test_hit_identity_and_counters(R, rb, gpu, scene_mid, scene_runs, lanes):
    """Per-pixel voxel id, mip level (and depth, inside the warp word) bit-exact; work counters equal."""
    import torch
    gpu.set_lanes_per_ray(lanes)
    for scene, h in ((scene_mid, -100.0), (scene_runs, -90.0)):
        gpu.all_to_gpu(scene)
        cfg = R.FrameConfig.default(512, 384)
        for pos, rot in few_cameras(h)[:4]:
            rm = R.RayMap(cfg).get_ray_map(pos, rot)
            _, want, want_ids, cnt = _oracle(rb, rm, scene, cfg, ids=True)
            ids = torch.full((cfg.rays_casted, cfg.render_size, 2), -1, dtype=torch.int32, device="cuda")
            _fresh_warp(gpu, cfg)
            gpu.render_ids(rm, cfg, ids.data_ptr())
            gpu.sync()
            assert np.array_equal(gpu.read_warp(cfg), want)
            got_ids = ids.cpu().numpy().view(np.uint32)
            assert np.array_equal(got_ids, want_ids), rot
            c = dict(zip(rb.COUNTER_NAMES, gpu.counters()))
            for k in ("pixels", "elems_rendered", "cols_fetched", "cols_nonempty", "elems_total", "run_iters",
                      "elems_processed", "voxels_processed", "cleared"):
                assert c[k] == cnt[k], (k, c[k], cnt[k], rot)
            assert c["dda_steps"] >= cnt["dda_steps"]      # batches of 32 crossings may overshoot the last one
    gpu.set_lanes_per_ray(0)


def test_degenerate_scenes_bit_exact(R, rb, gpu):
    """Nothing, everything, one voxel, white noise (irregular columns with many runs, top-attached spans), floating
    layers with a shaft, thin walls, a non-cubic grid, a one-voxel comb; cameras outside, inside matter, straight down
    and up.  Production kernel: warped buffer, hit identity and counters; every other kernel variant: warped buffer;
    final RGBA against the oracle's unwarp."""
    import torch
    cfg = R.FrameConfig.default(400, 300)
    rgba = torch.empty((cfg.height, cfg.width, 4), dtype=torch.uint8, device="cuda")
    for name, scene in edge_scenes(R).items():
        gpu.all_to_gpu(scene)
        for pos, rot in edge_cameras():
            rm = R.RayMap(cfg).get_ray_map(pos, rot)
            orm, want, want_ids, cnt = _oracle(rb, rm, scene, cfg, ids=True)
            gpu.set_lanes_per_ray(0)
            ids = torch.full((cfg.rays_casted, cfg.render_size, 2), -1, dtype=torch.int32, device="cuda")
            _fresh_warp(gpu, cfg)
            gpu.render_ids(rm, cfg, ids.data_ptr())
            gpu.sync()
            assert np.array_equal(gpu.read_warp(cfg), want), (name, pos, rot)
            assert np.array_equal(ids.cpu().numpy().view(np.uint32), want_ids), (name, pos, rot)
            c = dict(zip(rb.COUNTER_NAMES, gpu.counters()))
            for k in ("pixels", "elems_rendered", "cols_fetched", "cols_nonempty", "elems_total", "cleared"):
                assert c[k] == cnt[k], (name, k, c[k], cnt[k], rot)
            # Dead iterations: a column that closes its ray plane with crossed bounds keeps
            # iterating in the reference (clamped, empty spans; y_clip_min moves down, Cuda_Render.h:564-577); the
            # warp kernels stop it at the first run that breaks at projection time.  Seen on the white-noise scene
            # from just above its top only: 9492 vs 9512 iterations, picture and every other counter exact.
            for k in ("run_iters", "elems_processed", "voxels_processed"):
                assert abs(c[k] - cnt[k]) <= 0.005 * cnt[k], (name, k, c[k], cnt[k], rot)
                if name != "noise50":
                    assert c[k] == cnt[k], (name, k, c[k], cnt[k], rot)
            for lanes in _lanes_list(R, (0, 65, 68, 69, 64, 66, 32, 1)):
                gpu.set_lanes_per_ray(lanes)
                _fresh_warp(gpu, cfg)
                gpu.render(rm, cfg)
                assert np.array_equal(gpu.read_warp(cfg), want), (name, lanes, pos, rot)
            gpu.set_lanes_per_ray(0)
            gpu.unwarp(rm, cfg, d_rgba=rgba.data_ptr())
            gpu.sync()
            orgba = rb.orc_unwarp(orm, cfg.width, cfg.height, cfg.render_size, cfg.rays_casted, cfg.rays_casted_res, want)
            dmax, same = rgb_parity(rgba.cpu().numpy(), orgba)
            assert dmax <= 1 and same >= 0.999, (name, rot, dmax, same)
    gpu.set_lanes_per_ray(0)


def test_unwarp_rgba_parity(R, rb, gpu, scene_mid):
    gpu.all_to_gpu(scene_mid)
    for wh in ((640, 480), (1024, 768), (1920, 1080)):
        cfg = R.FrameConfig.default(*wh)
        for pos, rot in few_cameras(-100.0):
            rm = R.RayMap(cfg).get_ray_map(pos, rot)
            orm, want_warp, _, _ = _oracle(rb, rm, scene_mid, cfg)
            want = rb.orc_unwarp(orm, cfg.width, cfg.height, cfg.render_size, cfg.rays_casted, cfg.rays_casted_res, want_warp)
            host = np.zeros((cfg.height, cfg.width, 4), np.uint8)
            _fresh_warp(gpu, cfg)
            gpu.render(rm, cfg)
            gpu.unwarp(rm, cfg)
            gpu.sync()
            got = gpu.download(_rgba_ptr(gpu, cfg, rm), (cfg.height, cfg.width, 4), np.uint8)
            mx, same = rgb_parity(got, want)
            assert mx <= 1 and same >= 0.999, (wh, rot, mx, same)      # tolerance stated by the north star


def _rgba_ptr(gpu, cfg, rm):
    """Device RGBA of the last unwarp: re-run it into a torch buffer we own."""
    import torch
    buf = torch.zeros((cfg.height, cfg.width, 4), dtype=torch.uint8, device="cuda")
    gpu.unwarp(rm, cfg, d_rgba=buf.data_ptr())
    gpu.sync()
    _rgba_ptr.keep = buf
    return buf.data_ptr()


def test_soft_pass_parity(R, rb, gpu, scene_mid):
    """Depth-aware smoothing (GLSL pass 2, soft.frag) against its oracle restatement: <= 1 LSB, >= 99.9 % identical."""
    import torch
    gpu.all_to_gpu(scene_mid)
    for wh in ((640, 480), (1920, 1080), (2560, 1440)):          # the last one needs the 4096^2 FBO
        cfg = R.FrameConfig.default(*wh)
        for pos, rot in few_cameras(-100.0)[:3]:
            rm = R.RayMap(cfg).get_ray_map(pos, rot)
            a = torch.zeros((cfg.height, cfg.width, 4), dtype=torch.uint8, device="cuda")
            b = torch.zeros_like(a)
            gpu.render(rm, cfg)
            gpu.unwarp(rm, cfg, d_rgba=a.data_ptr())
            gpu.soft(cfg, a.data_ptr(), b.data_ptr())
            gpu.sync()
            want = rb.orc_soft(a.cpu().numpy())
            mx, same = rgb_parity(b.cpu().numpy(), want)
            assert mx <= 1 and same >= 0.999, (wh, rot, mx, same)
    with pytest.raises(R.RlercError):
        gpu.soft(cfg, a.data_ptr(), a.data_ptr())                # in place is refused


def test_render_frame_host_buffers_and_pipeline(R, rb, gpu, scene_mid):
    """The all-in-one C-ABI call with HOST buffers, synchronous and pipelined, equals the staged calls."""
    gpu.all_to_gpu(scene_mid)
    cfg = R.FrameConfig.default(800, 600)
    cams = few_cameras(-100.0)
    frames = []
    for pos, rot in cams:
        host = np.zeros((cfg.height, cfg.width, 4), np.uint8)
        rm = gpu.render_frame(pos, rot, cfg, host)
        orm, want_warp, _, _ = _oracle(rb, rm, scene_mid, cfg)
        want = rb.orc_unwarp(orm, cfg.width, cfg.height, cfg.render_size, cfg.rays_casted, cfg.rays_casted_res, want_warp)
        mx, same = rgb_parity(host, want)
        assert mx <= 1 and same >= 0.999
        frames.append(host)
    pins = [R.PinnedBuffer((cfg.height, cfg.width, 4)) for _ in cams]
    tickets = [gpu.frame_submit(pos, rot, cfg, pin.array) for (pos, rot), pin in zip(cams, pins)]
    for t, pin, want in zip(tickets, pins, frames):
        gpu.frame_wait(t)
        assert np.array_equal(pin.array, want)
    for pin in pins:
        pin.free()


def test_ray_slices_compose_to_the_full_frame(R, gpu, scene_mid):
    """Multi-GPU contract on one GPU: slices of the ray range, contiguous or interleaved, give the same
    warped buffer, and the per-slice images sum to the single image (disjoint support)."""
    import torch
    gpu.all_to_gpu(scene_mid)
    cfg = R.FrameConfig.default(1024, 768)
    pos, rot = few_cameras(-100.0)[0]
    rm = R.RayMap(cfg).get_ray_map(pos, rot)
    _fresh_warp(gpu, cfg)
    gpu.render(rm, cfg)
    full = gpu.read_warp(cfg)
    full_img = torch.zeros((cfg.height, cfg.width, 4), dtype=torch.uint8, device="cuda")
    gpu.unwarp(rm, cfg, d_rgba=full_img.data_ptr())
    gpu.sync()
    n = rm.map_line_count
    for cuts in ([0, n // 2, n], [0, 7, n // 3, n - 1, n]):
        _fresh_warp(gpu, cfg)
        acc = torch.zeros((cfg.height, cfg.width, 4), dtype=torch.int32, device="cuda")
        for b, e in zip(cuts[:-1], cuts[1:]):
            gpu.render(rm, cfg, ray_begin=b, ray_end=e)
            part = torch.zeros((cfg.height, cfg.width, 4), dtype=torch.uint8, device="cuda")
            gpu.unwarp_slice(rm, cfg, b, e, d_rgba=part.data_ptr())
            gpu.sync()
            acc += part
        assert np.array_equal(gpu.read_warp(cfg), full)
        assert torch.equal(acc.to(torch.uint8), full_img)
    for block, world in ((32, 2), (32, 8), (128, 3)):
        _fresh_warp(gpu, cfg)
        acc = torch.zeros((cfg.height, cfg.width, 4), dtype=torch.int32, device="cuda")
        for rank in range(world):
            gpu.render_interleaved(rm, cfg, block, world, rank)
            part = torch.zeros((cfg.height, cfg.width, 4), dtype=torch.uint8, device="cuda")
            gpu.unwarp_interleaved(rm, cfg, block, world, rank, d_rgba=part.data_ptr())
            gpu.sync()
            acc += part
        assert np.array_equal(gpu.read_warp(cfg), full), (block, world)
        assert torch.equal(acc.to(torch.uint8), full_img), (block, world)


def test_frame_does_not_depend_on_history(R, rb, gpu, scene_mid):
    """Texels the reference leaves stale (outside a ray's clip range) are defined as 0 here: a frame
    rendered over garbage equals the frame rendered over zeros; rows of unused ray planes stay untouched."""
    gpu.all_to_gpu(scene_mid)
    cfg = R.FrameConfig.default(640, 480)
    for lanes in (0, 32):
        gpu.set_lanes_per_ray(lanes)
        for pos, rot in few_cameras(-100.0)[:3]:
            rm = R.RayMap(cfg).get_ray_map(pos, rot)
            _, want, _, _ = _oracle(rb, rm, scene_mid, cfg)
            wp = gpu.warp_buffer(cfg)
            gpu.upload(wp, np.full((cfg.rays_casted, cfg.render_size), 0xdeadbeef, np.uint32))
            gpu.render(rm, cfg)
            got = gpu.read_warp(cfg)
            n = rm.map_line_count
            assert np.array_equal(got[:n], want[:n])
            assert np.all(got[n:] == 0xdeadbeef)
    gpu.set_lanes_per_ray(0)


def test_physical_tiling_equals_coordinate_wrap(R, gpu, scene_small):
    """Size-independent property: the traversal wraps coordinates with & (grid-1) (Cuda_Render.h:441-442),
    so a physically 4x4-tiled scene must render the identical frame."""
    cfg = R.FrameConfig.default(1024, 768)
    tiled = scene_small.tile(4, 4)
    outs = []
    for scene in (scene_small, tiled):
        gpu.all_to_gpu(scene)
        frames = []
        for pos, rot in few_cameras(-40.0)[:3]:
            rm = R.RayMap(cfg).get_ray_map(pos, rot)
            _fresh_warp(gpu, cfg)
            gpu.render(rm, cfg)
            frames.append(gpu.read_warp(cfg))
        outs.append(frames)
    # the tiled pyramid has the same number of levels; beyond the last level both wrap identically
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


def test_full_size_properties_1080p(R, gpu):
    """BASELINE config-2 sizes (256^3 stand-in scene for speed): idempotence, conservation, lane-count
    independence, stale-row behaviour."""
    scene = R.RLE4.synth(0, 256, 256, 256, seed=1)
    gpu.all_to_gpu(scene)
    cfg = R.FrameConfig.default(1920, 1080)
    import torch
    for t in (0, 250, 600):
        pos, rot = R.flythrough_pose(t)
        pos = (pos[0], pos[1] * 0.25, pos[2])
        rm = R.RayMap(cfg).get_ray_map(pos, rot)
        outs = []
        for lanes in (0, 32, 0):
            gpu.set_lanes_per_ray(lanes)
            _fresh_warp(gpu, cfg)
            ids = torch.full((cfg.rays_casted, cfg.render_size, 2), -1, dtype=torch.int32, device="cuda")
            gpu.render_ids(rm, cfg, ids.data_ptr())
            gpu.sync()
            outs.append(gpu.read_warp(cfg))
            c = dict(zip(["elems_total", "elems_processed", "voxels_processed", "elems_rendered", "pixels",
                          "cols_fetched", "run_iters", "cols_nonempty", "cleared", "dda_steps"], gpu.counters()))
            w = outs[-1]
            has_id = ids.cpu().numpy()[..., 0] != -1
            sky = int((w == 0xff8844).sum())
            hit = int(has_id.sum())
            assert hit == c["pixels"] and sky + hit == c["cleared"]      # every cleared pixel is sky or one hit
            assert not np.any(w[has_id] == 0xff8844)
            assert np.all((w[has_id] >> 16) % 2 == 0)                     # depth is forced even (Cuda_Render.h:708)
        assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
        # rows beyond map_line_count are never touched (Cuda_Main.cu:163)
        assert not outs[0][rm.map_line_count:].any()
    gpu.set_lanes_per_ray(0)


def test_legacy_cuda_main_render2(R, rb, gpu, scene_small):
    """The reference's own entry point, same signature (Cuda_Main.cu:124,183)."""
    import torch
    cfg = R.FrameConfig.default(512, 384)
    lib = R.lib()
    R._check(lib.rlerc_legacy_init(0, scene_small._h, C.byref(cfg)))
    buf = torch.zeros((cfg.rays_casted, cfg.render_size), dtype=torch.int32, device="cuda")
    lib.pboRegister(7)
    lib.rlerc_pbo_bind(7, buf.data_ptr())
    pos, rot = few_cameras(-40.0)[0]
    rm = R.RayMap(cfg).get_ray_map(pos, rot)
    lib.cuda_main_render2(7, cfg.render_size, cfg.render_size, C.byref(rm))
    _, want, _, _ = _oracle(rb, rm, scene_small, cfg)
    assert np.array_equal(buf.cpu().numpy().view(np.uint32), want)
    lib.cuda_main_render2(0, cfg.render_size, cfg.render_size, C.byref(rm))     # pbo 0: no-op (Cuda_Main.cu:187)
    lib.pboUnregister(7)
    # gpu_malloc / gpu_memcpy / cpu_memcpy (core.h:144-147)
    p = lib.gpu_malloc(4096)
    assert p
    src = np.arange(1024, dtype=np.uint32)
    dst = np.zeros_like(src)
    lib.gpu_memcpy(p, src.ctypes.data, 4096)
    lib.cpu_memcpy(dst.ctypes.data, p, 4096)
    assert np.array_equal(src, dst)


def test_error_paths_on_device(R, gpu, scene_small):
    cfg = R.FrameConfig.default(512, 384)
    r2 = R.Renderer(0)
    rm = R.RayMap(cfg).get_ray_map((0, -40, 0), (0.3, 1.0, 0))
    with pytest.raises(R.RlercError):
        r2.render(rm, cfg)                        # no scene uploaded
    with pytest.raises(R.RlercError):
        r2.set_lanes_per_ray(3)
    bad = R.FrameConfig.default(512, 384)
    bad.render_size = 7
    r2.all_to_gpu(scene_small)
    with pytest.raises(R.RlercError):
        r2.render(rm, bad)
    with pytest.raises(R.RlercError):
        R.Renderer(99)
    # a non power-of-two grid is refused at upload (the kernel wraps with & (grid-1))
    sx, sy, sz, mp, sl = scene_small.level(0)
    m4 = R.Map4(); m4.sx, m4.sy, m4.sz, m4.slabs_size = 48, sy, 64, len(sl)
    m4.map, m4.slabs = mp.ctypes.data, sl.ctypes.data
    with pytest.raises(R.RlercError):
        r2.all_to_gpu(R.RLE4.from_maps([m4]))
    r2.close()


def test_production_build_equals_instrumented_build_up_to_8k(R, rb, gpu):
    """BASELINE config 2-5 sizes (1080p, 4K, 8K): the production build (column filter, B0 / B1 batch paths) and the
    instrumented build of the same kernel (every column through, owner-lane event loop) give the same warped buffer;
    a slice of the ray planes is also checked against the oracle directly."""
    import torch
    scenes = ((R.RLE4.synth(0, 256, 256, 256, seed=1), 0.25), (R.RLE4.synth(1, 128, 64, 128, seed=42), 0.03))
    for scene, hscale in scenes:
        gpu.all_to_gpu(scene)
        for wh in ((1920, 1080), (3840, 2160), (7680, 4320)):
            cfg = R.FrameConfig.default(*wh)
            for t in ((0, 600) if wh[0] < 7680 else (250,)):
                pos, rot = R.flythrough_pose(t)
                pos = (pos[0], pos[1] * hscale, pos[2])
                rm = R.RayMap(cfg).get_ray_map(pos, rot)
                gpu.set_lanes_per_ray(0)
                _fresh_warp(gpu, cfg)
                gpu.render(rm, cfg)
                gpu.sync()
                prod = gpu.read_warp(cfg)
                ids = torch.empty((cfg.rays_casted, cfg.render_size, 2), dtype=torch.int32, device="cuda")
                _fresh_warp(gpu, cfg)
                gpu.render_ids(rm, cfg, ids.data_ptr())
                gpu.sync()
                inst = gpu.read_warp(cfg)
                del ids
                assert np.array_equal(prod, inst), (wh, t, int((prod != inst).sum()))
                # oracle on 192 ray planes around the middle of the frame
                r0 = max(0, rm.map_line_count // 2 - 96)
                r1 = min(rm.map_line_count, r0 + 192)
                orm = oracle_raymap(rb, rm, scene)
                want, _, _ = rb.orc_render(orm, cfg.render_size, cfg.rays_casted, cfg.mip_distance, cfg.z_far, ray_begin=r0, ray_end=r1)
                assert np.array_equal(prod[r0:r1], want[r0:r1]), (wh, t)
                del prod, inst, want
    gpu.set_lanes_per_ray(0)


@pytest.mark.parametrize("flags", [1, 2, 3])
def test_core_h_options_clipregion_height_color(R, rb, gpu, scene_small, scene_mid, flags):
    """The reference's CLIPREGION / HEIGHT_COLOR compile-time options (R/src/core.h:18,22) as run-time flags of the
    frame config: production kernel == oracle, bit for bit; other kernel variants refuse them."""
    cams = [((64.5, -60.0, 64.5), (0.5, 0.8, 0.0)), ((3.25, -40.0, 120.5), (0.3, 2.4, 0.0)),
            ((-50.0, -80.0, -50.0), (0.45, 0.785 + 1.5708, 0.0)), ((300.0, -120.0, 64.0), (0.4, 4.71, 0.0)),
            ((10000.0, -100.0, 10000.0), (0.4, 1.9, 0.0))]
    gpu.set_lanes_per_ray(0)
    for scene in (scene_small, scene_mid):
        gpu.all_to_gpu(scene)
        for wh in ((512, 384), (1920, 1080)):
            cfg = R.FrameConfig.default(*wh)
            cfg.flags = flags
            for pos, rot in cams:
                rm = R.RayMap(cfg).get_ray_map(pos, rot)
                orm = oracle_raymap(rb, rm, scene)
                want, _, _ = rb.orc_render(orm, cfg.render_size, cfg.rays_casted, cfg.mip_distance, cfg.z_far, flags=flags)
                for lanes in (0, 65, 68, 69):      # automatic choice, k_traverse_f, k_traverse_p, k_traverse_q
                    gpu.set_lanes_per_ray(lanes)
                    _fresh_warp(gpu, cfg)
                    gpu.render(rm, cfg)
                    got = gpu.read_warp(cfg)
                    assert np.array_equal(got, want), (flags, lanes, wh, pos, int((got != want).sum()))
    cfg = R.FrameConfig.default(512, 384)
    cfg.flags = flags
    gpu.set_lanes_per_ray(32)                      # the simple lane <-> run kernels do not implement the flags
    with pytest.raises(R.RlercError):
        gpu.render(R.RayMap(cfg).get_ray_map(*cams[0]), cfg)
    gpu.set_lanes_per_ray(0)


def test_two_devices_in_one_process(R, rb, scene_small):
    """One process, one context per GPU: kernels, shared-memory opt-in and buffers are
    per device.  Skipped on a single-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cfg = R.FrameConfig.default(1920, 1080)          # > 48 KB of dynamic shared memory per block
    pos, rot = few_cameras(-40.0)[0]
    rm = R.RayMap(cfg).get_ray_map(pos, rot)
    orm = oracle_raymap(rb, rm, scene_small)
    want, _, _ = rb.orc_render(orm, cfg.render_size, cfg.rays_casted, cfg.mip_distance, cfg.z_far)
    outs = []
    for dev in (0, 1):
        r = R.Renderer(dev)
        r.all_to_gpu(scene_small)
        r.render(rm, cfg)
        r.unwarp(rm, cfg)
        outs.append(r.read_warp(cfg))
        r.close()
    assert np.array_equal(outs[0], want) and np.array_equal(outs[1], want)


def test_more_ray_planes_than_buffer_rows(R, rb, gpu, scene_small):
    """map_line_count > RAYS_CASTED (square window, steep pitch): the kernel renders rays_casted ray planes
    (R/src/Cuda_Main.cu:196) and touches nothing beyond the buffer."""
    gpu.all_to_gpu(scene_small)
    gpu.set_lanes_per_ray(0)
    cfg = R.FrameConfig.default(512, 512)
    pos, rot = (812.2, -40.0, 18034.6), (-1.20644718889272, 1.5707963267948966, 0.0)
    rm = R.RayMap(cfg).get_ray_map(pos, rot)
    assert rm.map_line_count > cfg.rays_casted
    _, want, _, _ = _oracle(rb, rm, scene_small, cfg)
    _fresh_warp(gpu, cfg)
    gpu.render(rm, cfg)
    assert np.array_equal(gpu.read_warp(cfg), want)



def _texels(R, gpu, rm, cfg):
    import torch
    out = torch.zeros((cfg.height, cfg.width), dtype=torch.int32, device="cuda")
    f = R.lib().rlerc_debug_unwarp_texels
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    assert f(gpu._c, C.byref(rm), C.byref(cfg), out.data_ptr()) == 0
    gpu.sync()
    t = out.cpu().numpy().view(np.uint32)
    return np.stack([t >> 16, t & 0xffff], axis=-1).astype(np.int32)


@pytest.mark.parametrize("wh", [(1024, 768), (1920, 1080), (3840, 2160), (7680, 4320), (641, 479)])
def test_unwarp_texel_selection_equals_oracle(R, rb, gpu, scene_small, wh):
    """k_unwarp evaluates the shader's texel arithmetic with row / column invariants hoisted and only the selected
    branch of every step() blend (csrc/kernels.cu): the (ray plane, texel) it samples must equal the oracle's
    statement-by-statement evaluation of colorize_buddha_soft.frag:19-85 at EVERY pixel of every BASELINE window,
    for cameras in all quadrant configurations, incl. a level camera (vanishing point at infinity: generic path)."""
    gpu.all_to_gpu(scene_small)
    cfg = R.FrameConfig.default(*wh)
    cams = list(camera_grid(-40.0)) if wh[0] <= 1920 else few_cameras(-40.0)
    cams += [((10000.0, -40.0, 10000.0), (0.0, 0.3, 0.0)), ((10000.0, -40.0, 10000.0), (1e-7, 2.0, 0.0))]
    warp = np.zeros((cfg.rays_casted, cfg.render_size), np.uint32)
    for pos, rot in cams:
        rm = R.RayMap(cfg).get_ray_map(pos, rot)
        orm = oracle_raymap(rb, rm, scene_small)
        _, want = rb.orc_unwarp(orm, cfg.width, cfg.height, cfg.render_size, cfg.rays_casted, cfg.rays_casted_res, warp, want_texels=True)
        got = _texels(R, gpu, rm, cfg)
        assert np.array_equal(got, want), (wh, rot, int((got != want).any(axis=-1).sum()))


def test_unwarp_texel_selection_independent_restatement(R, gpu, scene_small):
    """A second, independent restatement of the unwarp geometry (written from the shader's MEANING:
    boolean segment selection, no step() arithmetic; numpy float32) picks the same texel as k_unwarp wherever its own
    arithmetic is well away from a texel boundary."""
    cfg = R.FrameConfig.default(1024, 768)
    W, H, RS, RC, RCR = cfg.width, cfg.height, cfg.render_size, cfg.rays_casted, cfg.rays_casted_res
    f = np.float32
    for pos, rot in list(camera_grid(-40.0))[::3]:
        rm = R.RayMap(cfg).get_ray_map(pos, rot)
        got = _texels(R, gpu, rm, cfg)
        vp = rm.vanishing_point_2d
        vx, vy = f(1) - f(vp.x), (f(1) - f(vp.y) - f(rm.border)) * f(W) / f(H)
        o1 = f(4) * f(rm.res[0]) / f(RCR); o2_ = o1 + f(4) * f(rm.res[1]) / f(RCR); o3 = o2_ + f(4) * f(rm.res[2]) / f(RCR)
        ofs = [-f(rm.p_ofs_min[0]), -f(rm.p_ofs_min[1]) + o1, -f(rm.p_ofs_min[2]) + o2_, -f(rm.p_ofs_min[3]) + o3]
        px, py = np.meshgrid(np.arange(W, dtype=f) + f(0.5), (H - 1 - np.arange(H, dtype=f)) + f(0.5))
        with np.errstate(all="ignore"):
            dx, dy = px / f(W) - vx, py / f(H) - vy
            b = f(W - H) / f(2 * W)
            upper, left = dy <= 0, dx <= 0
            horiz = np.abs(dy) - np.abs(dx) * f(W) / f(H) <= 0
            o2 = np.where(horiz, px, py) / f(W)
            ang2 = dx * np.abs(f(1) - upper.astype(f) - vy) / dy + np.where(upper, f(1) - vx, vx)
            ang3 = (dy * np.abs(f(1) - left.astype(f) - vx) / dx + np.where(left, f(1) - vy, vy)) * f(H) / f(W) + b
            xp = np.where(horiz, ang3, ang2)
            ty = np.where(~horiz, np.where(upper, ofs[1] + xp, ofs[0] + f(1) - xp), np.where(left, ofs[3] + xp, ofs[2] + f(1) - xp))
            ty = ty * (f(RCR) / f(RC)) * f(0.25)
            flip = not (rm.rotation.x > 0)
            tx = np.where(~horiz, np.where(upper != flip, f(1) - (o2 + b), o2 + b), np.where(left != flip, f(1) - o2, o2))
            fy, fx = ty * f(RC), tx * f(RS)
        sure = np.isfinite(fy) & np.isfinite(fx) & (np.abs(fy - np.round(fy)) > 1e-2) & (np.abs(fx - np.round(fx)) > 1e-2)
        iy = np.clip(np.floor(np.where(sure, fy, 0)), 0, RC - 1).astype(np.int32)
        ix = np.clip(np.floor(np.where(sure, fx, 0)), 0, RS - 1).astype(np.int32)
        assert sure.mean() > 0.9
        assert np.array_equal(got[..., 0][sure], iy[sure]) and np.array_equal(got[..., 1][sure], ix[sure]), rot


def _solo_frame(R, r, rm, cfg):
    import torch
    solo = torch.zeros((cfg.height, cfg.width, 4), dtype=torch.uint8, device="cuda:%d" % r.device)
    r.set_lanes_per_ray(0)
    r.frame_device(rm, cfg, 1, 1, 0, solo.data_ptr())
    r.sync()
    return solo.cpu().numpy()


def test_group_of_one_and_multi_renderer(R, rb, gpu, scene_mid):
    """The multi-GPU entry points with ONE member are the plain frame (runs on the single-GPU box): rlerc_group_*
    with nranks = 1 and rlerc_create_multi([0]); frames in flight keep their order."""
    gpu.all_to_gpu(scene_mid)
    cfg = R.FrameConfig.default(640, 480)
    cams = few_cameras(-100.0)
    maps = []
    for pos, rot in cams:
        rm = R.RayMapGPU()
        C.memmove(C.byref(rm), C.byref(R.RayMap(cfg).get_ray_map(pos, rot)), 896)
        maps.append(rm)
    want = [_solo_frame(R, gpu, rm, cfg) for rm in maps]
    g = R.Group(gpu, cfg, 0, 1, depth=3)
    hosts = [R.PinnedBuffer((cfg.height, cfg.width, 4)) for _ in maps]
    tickets = [g.submit(rm, 0, hosts[i].array) for i, rm in enumerate(maps)]
    for t in tickets:
        g.wait(t)
    g.sync()
    for i in range(len(maps)):
        assert np.array_equal(hosts[i].array, want[i]), i
    g.close()
    m = R.MultiRenderer([0], depth=2)
    m.all_to_gpu(scene_mid)
    for i, (pos, rot) in enumerate(cams):
        hosts[i].array[:] = 0
        m.render_frame(pos, rot, cfg, hosts[i].array)
        assert np.array_equal(hosts[i].array, want[i]), i
    m.close()
    for h in hosts:
        h.free()


def test_multi_gpu_frame_in_one_process(R, rb, scene_mid):
    """rlerc_create_multi on every GPU of the box (>= 2): slices traversed per GPU, bands unwarped per GPU with texels
    pulled over NVLink, every GPU copies its band to the host: the frame equals the single-GPU frame bit for bit,
    synchronous and with frames in flight."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    r0 = R.Renderer(0)
    r0.all_to_gpu(scene_mid)
    for wh in ((1024, 768), (1920, 1080)):
        cfg = R.FrameConfig.default(*wh)
        cams = few_cameras(-100.0)
        want = [_solo_frame(R, r0, R.RayMap(cfg).get_ray_map(p, q), cfg) for p, q in cams]
        for nd in sorted({2, n}):
            m = R.MultiRenderer(list(range(nd)), depth=3)
            m.all_to_gpu(scene_mid)
            hosts = [R.PinnedBuffer((cfg.height, cfg.width, 4)) for _ in cams]
            for i, (p, q) in enumerate(cams):
                m.render_frame(p, q, cfg, hosts[i].array)
                assert np.array_equal(hosts[i].array, want[i]), (wh, nd, i)
                hosts[i].array[:] = 0
            tickets = [m.frame_submit(p, q, cfg, hosts[i].array) for i, (p, q) in enumerate(cams)]
            for t in tickets:
                m.frame_wait(t)
            for i in range(len(cams)):
                assert np.array_equal(hosts[i].array, want[i]), (wh, nd, i, "pipelined")
            m.close()
            for h in hosts:
                h.free()
    r0.close()


def test_multi_gpu_frame_one_process_per_gpu(R):
    """The same under torchrun, one process per GPU (CUDA IPC mappings): tools/group_probe.py compares the frame
    assembled on rank 0 and every rank's band with single-GPU frames and exits non-zero on any difference."""
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(HERE)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(root, "tools", "group_probe.py"), "small", "12", "3"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, (p.stdout[-3000:], p.stderr[-1500:])
    assert "mismatches 0" in p.stdout


def test_unwarp_2xaa_shader_parity(R, rb, gpu, scene_mid):
    """RLERC_FLAG_SHADER_2XAA: pass 1 of the reference's -DANTIALIAS build (colorize_buddha_soft_2xAA.frag, selected at
    R/src/main.cpp:510-522) against the oracle's restatement of that shader: <= 1 LSB, >= 99.9 % identical (parity
    unpinned like every GLSL restatement: there is no GL here)."""
    import torch
    gpu.all_to_gpu(scene_mid)
    gpu.set_lanes_per_ray(0)
    for wh in ((640, 480), (1920, 1080)):
        cfg = R.FrameConfig.default(*wh)
        aa = R.FrameConfig.default(*wh)
        aa.flags = R.FrameConfig.SHADER_2XAA
        for pos, rot in few_cameras(-100.0)[:3]:
            rm = R.RayMap(cfg).get_ray_map(pos, rot)
            orm, want_warp, _, _ = _oracle(rb, rm, scene_mid, cfg)
            want = rb.orc_unwarp(orm, cfg.width, cfg.height, cfg.render_size, cfg.rays_casted, cfg.rays_casted_res, want_warp, shader=1)
            plain = rb.orc_unwarp(orm, cfg.width, cfg.height, cfg.render_size, cfg.rays_casted, cfg.rays_casted_res, want_warp)
            assert not np.array_equal(want[..., :3], plain[..., :3])          # a different picture ...
            assert np.array_equal(want[..., 3], plain[..., 3])                # ... with the same smoothing weights
            _fresh_warp(gpu, cfg)
            gpu.render(rm, cfg)
            buf = torch.zeros((cfg.height, cfg.width, 4), dtype=torch.uint8, device="cuda")
            gpu.unwarp(rm, aa, d_rgba=buf.data_ptr())
            gpu.sync()
            mx, same = rgb_parity(buf.cpu().numpy(), want)
            assert mx <= 1 and same >= 0.999, (wh, rot, mx, same)


def test_lod_streaming_frames_equal_full_replica(R, rb, gpu):
    """LOD streaming (rlerc_scene_upload_streamed + rlerc_stream_prepare, csrc/stream.cu): of every mip level only the
    z-rows a frame can reach are resident (virtual address ranges, physical chunks mapped on demand).  Along a camera
    path — incl. across the wrap-around of the infinite tiling and a jump — every frame is bit-identical to the full
    replica's, far less than the whole scene is resident, and a moving camera uploads only the newly reached rows.  A
    frame rendered WITHOUT the prepare call for its camera touches unmapped memory: a CUDA error, not a wrong picture."""
    import torch
    scene = R.RLE4.synth_rle(4096, 256, 4096, seed=7, band_every=16)
    cfg = R.FrameConfig.default(640, 480)
    full = R.Renderer(0)
    full.all_to_gpu(scene)
    st = R.Renderer(0)
    st.all_to_gpu_streamed(scene)
    path = [((2000.0 + 37.0 * i, -60.0, 2000.0 + 61.0 * i), (0.35, 0.3 + 0.02 * i, 0.0)) for i in range(12)]
    path += [((100.0, -60.0, 4090.0 + 3.0 * i), (0.2, 1.6, 0.0)) for i in range(4)]          # crosses z = 4096: the rows wrap
    path += [((1.0e6 + 3000.0, -500.0, -2.5e5), (0.6, 4.0, 0.0))]                            # a jump, far outside the base tile
    uploaded = []
    for i, (pos, rot) in enumerate(path):
        s = st.stream_prepare(pos, rot, cfg, margin_voxels=128)
        uploaded.append(s.uploaded_bytes)
        assert s.resident_bytes <= s.total_bytes + (64 << 20)
        rm = R.RayMap(cfg).get_ray_map(pos, rot)
        full.render(rm, cfg)
        st.render(rm, cfg)
        a, b = full.read_warp(cfg), st.read_warp(cfg)
        assert np.array_equal(a, b), (i, pos, rot)
        if i == 0:
            first = s.resident_bytes
            assert first < 0.75 * s.total_bytes, (first, s.total_bytes)      # level 0 dominates the scene and is only partly reachable
    # steady motion: the frames after the first upload little (only the rows that came into reach)
    assert max(uploaded[1:12]) < 0.25 * uploaded[0], uploaded
    # the oracle agrees with the streamed replica too
    pos, rot = path[5]
    st.stream_prepare(pos, rot, cfg, margin_voxels=128)
    rm = R.RayMap(cfg).get_ray_map(pos, rot)
    _fresh_warp(st, cfg)                    # rows beyond this frame's ray planes still hold earlier frames
    st.render(rm, cfg)
    _, want, _, _ = _oracle(rb, rm, scene, cfg)
    assert np.array_equal(st.read_warp(cfg), want)
    st.close()
    full.close()


def test_view_batch_two_gpus_in_one_process(R, scene_mid):
    """BASELINE config 5 in small: every GPU renders the whole frame of ITS OWN camera and delivers it into GPU 0's view
    array over NVLink (rlerc_group_enable_views / rlerc_group_submit_view); each view equals the single-GPU frame of
    that camera, also with frames in flight and after the slots have been recycled."""
    import torch
    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("needs two GPUs")
    cfg = R.FrameConfig.default(800, 600)
    rs = [R.Renderer(d) for d in range(n)]
    for r in rs:
        r.all_to_gpu(scene_mid)
    gs = [R.Group(rs[d], cfg, d, n, depth=2, views=True) for d in range(n)]
    blobs = b"".join(g.export() for g in gs)
    for g in gs:
        g.connect(blobs)
    cams = few_cameras(-100.0) + list(camera_grid(-100.0))[:7]
    rounds = len(cams) // n
    for rd in range(rounds):
        maps = []
        for d in range(n):
            rm = R.RayMapGPU()
            C.memmove(C.byref(rm), C.byref(R.RayMap(cfg).get_ray_map(*cams[rd * n + d])), 896)
            maps.append(rm)
        tickets = [gs[d].submit_view(maps[d], 0) for d in range(n)]
        for d in range(n):
            gs[d].wait(tickets[d])
        ptr, stride = gs[0].views(tickets[0])
        img0, _, _ = gs[0].image(tickets[0])
        for d in range(n):
            got = rs[0].download(img0 if d == 0 else ptr + d * stride, (cfg.height, cfg.width, 4), np.uint8)
            want = _solo_frame(R, rs[0], maps[d], cfg)
            assert np.array_equal(got, want), (rd, d)
    for g in gs:
        g.close()
    for r in rs:
        r.close()
