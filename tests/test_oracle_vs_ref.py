"""Pins the oracle port (oracle/rlerc_oracle.cpp) to the reference itself, compiled as host C++
from /root/reference (oracle/_ref).  Runs wherever oracle/_ref has been built."""
import ctypes as C

import numpy as np
import pytest

from util import camera_grid, edge_cameras, edge_scenes, few_cameras, levels_of, oracle_raymap


@pytest.fixture(scope="module")
def need_ref(have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref not built (needs /root/reference)")


def _both(R, rb, scene, cfg, pos, rot):
    rm = R.RayMap(cfg).get_ray_map(pos, rot)
    orm = oracle_raymap(rb, rm, scene)
    ref, _ = rb.ref_render_frame(orm, cfg.render_size, mip_distance=cfg.mip_distance, z_far=cfg.z_far, rays=cfg.rays_casted)
    port, ids, cnt = rb.orc_render(orm, cfg.render_size, cfg.rays_casted, cfg.mip_distance, cfg.z_far, want_ids=True)
    return ref, port, ids, cnt


def test_render_line_port_equals_reference_on_camera_grid(R, rb, need_ref, scene_mid):
    cfg = R.FrameConfig.default(640, 480)
    for pos, rot in camera_grid(-100.0):
        ref, port, ids, cnt = _both(R, rb, scene_mid, cfg, pos, rot)
        assert np.array_equal(ref, port), rot
        # every written pixel has an identity, and nothing else does
        has_id = ids[..., 0] != 0xffffffff
        assert int(has_id.sum()) == cnt["pixels"]
        assert not np.any(port[has_id] == 0xff8844)               # a hit never looks like sky (depth is even)
        assert np.all(has_id[(port != 0) & (port != 0xff8844)])


def test_render_line_port_equals_reference_other_scenes(R, rb, need_ref, scene_small, scene_runs, scene_rle):
    for scene, h in ((scene_small, -40.0), (scene_runs, -90.0), (scene_rle, -30.0)):
        cfg = R.FrameConfig.default(512, 384)
        for pos, rot in few_cameras(h):
            ref, port, ids, cnt = _both(R, rb, scene, cfg, pos, rot)
            assert np.array_equal(ref, port), (h, rot)


def test_render_line_port_equals_reference_on_degenerate_scenes(R, rb, need_ref):
    """Nothing, everything, one voxel, white noise, floating layers, thin walls, a non-cubic grid, a one-voxel comb:
    cameras outside, inside solid matter, inside the shaft, looking straight down / up."""
    cfg = R.FrameConfig.default(400, 300)
    for name, scene in edge_scenes(R).items():
        for pos, rot in edge_cameras():
            ref, port, ids, cnt = _both(R, rb, scene, cfg, pos, rot)
            assert np.array_equal(ref, port), (name, pos, rot)
            if name == "air":
                assert cnt["pixels"] == 0 and cnt["run_iters"] == 0


@pytest.mark.parametrize("mode", ["centerseg", "normalclip"])
def test_reference_culling_modes_render_the_same_picture(R, rb, need_ref, scene_mid, scene_runs, mode):
    """the reference built with -DCENTERSEG / -DNORMALCLIP (alternative culling, R/src/core.h:
    27-30) writes the SAME warped ray buffer as the shipped configuration, also on the adversarial scenes: these
    modes are ablations of the culling work, not of the picture, so the B200 path is a drop-in for those builds too.
    (-DPERPIXELFORWARD is not: with SHAREMEMCLIP on its pixel loop never terminates, Cuda_Render.h:682-731, and with
    the mask off it draws different pictures)"""
    cfg = R.FrameConfig.default(400, 300)
    scenes = [(scene_mid, list(camera_grid(-100.0))[::4] + few_cameras(-100.0)), (scene_runs, few_cameras(-90.0))]
    scenes += [(s, edge_cameras()) for k, s in edge_scenes(R).items() if k in ("noise50", "layers", "walls", "comb")]
    n = 0
    for scene, cams in scenes:
        for pos, rot in cams:
            rm = R.RayMap(cfg).get_ray_map(pos, rot)
            orm = oracle_raymap(rb, rm, scene)
            a, _ = rb.ref_render_frame(orm, cfg.render_size, mip_distance=cfg.mip_distance, z_far=cfg.z_far, rays=cfg.rays_casted)
            b, _ = rb.ref_render_frame(orm, cfg.render_size, mip_distance=cfg.mip_distance, z_far=cfg.z_far, rays=cfg.rays_casted, flags=mode)
            assert np.array_equal(a, b), (mode, pos, rot)
            n += 1
    assert n >= 40


def test_render_line_port_non_default_config(R, rb, need_ref, scene_small):
    """z_far, mip_distance and a non power-of-two render size are run-time here."""
    cfg = R.FrameConfig.default(600, 400)
    cfg.z_far = 5000
    cfg.mip_distance = 300
    for pos, rot in few_cameras(-40.0)[:3]:
        ref, port, ids, cnt = _both(R, rb, scene_small, cfg, pos, rot)
        assert np.array_equal(ref, port), rot


def test_detail_bench_counters(R, rb, need_ref, scene_mid):
    """The reference's own DETAIL_BENCH counters (Cuda_Render.h:14-22).  That build changes the control
    flow (Cuda_Render.h:369-372,505-508: no early return, and columns are skipped once
    y_clip_min>>1 >= y_clip_max>>1), so it stops a ray up to one pixel early and keeps counting
    elems_total to z_far: its counters bound ours, they do not equal them."""
    cfg = R.FrameConfig.default(640, 480)
    pos, rot = few_cameras(-100.0)[0]
    rm = R.RayMap(cfg).get_ray_map(pos, rot)
    orm = oracle_raymap(rb, rm, scene_mid)
    _, perf = rb.ref_render_frame(orm, cfg.render_size, mip_distance=cfg.mip_distance, z_far=cfg.z_far, bench=True)
    _, _, cnt = rb.orc_render(orm, cfg.render_size, cfg.rays_casted, cfg.mip_distance, cfg.z_far)
    assert 0.995 * cnt["pixels"] <= perf[4] <= cnt["pixels"]
    assert 0.99 * cnt["elems_rendered"] <= perf[3] <= cnt["elems_rendered"]
    assert perf[0] >= cnt["elems_total"] > 0


def test_compressor_equals_reference(R, rb, need_ref):
    """rlerc_scene_compress vs RLE4::compress_all on a CSG volume built with the reference's Tree."""
    lib = rb.ref_host()
    N = 64
    for color in (1, 0):
        t = lib.ref_tree_new(N, N, N, color)
        lib.ref_tree_cube(t, 0, 40, 0, N, 50, N)
        lib.ref_tree_sphere(t, 30, 25, 30, 12, 0)
        lib.ref_tree_set_color(t, 3)
        lib.ref_tree_sphere(t, 10, 38, 12, 6, 0)
        lib.ref_tree_set_color(t, 0)
        lib.ref_tree_sphere(t, 50, 38, 40, 7, 0)
        lib.ref_tree_sphere(t, 20, 45, 50, 9, 1)      # carve
        nb = N * N * N // 8
        grab = lambda p: np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nb,)).copy()
        vox = grab(lib.ref_tree_voxel(t))
        c1 = grab(lib.ref_tree_col1(t)) if color else None
        c2 = grab(lib.ref_tree_col2(t)) if color else None
        ref = rb.RefScene(lib.ref_compress_all(t))
        mine = R.RLE4.compress_all(vox, N, N, N, c1, c2)
        assert ref.nummaps == mine.nummaps == 6
        for m in range(ref.nummaps):
            a, b = ref.level(m, 1), mine.level(m)
            assert a[:3] == b[:3]
            if a[0] < 8:
                # levels narrower than 8 voxels: the reference's Tree::init allocates (sx/8)*sy*sz = 0
                # bytes and writes past it (tree.h:132-137) — undefined contents, not compared
                continue
            assert np.array_equal(a[4], b[4]), m
            assert np.array_equal(a[3], b[3][0::2]), m     # compress() emits offsets only (Rle4.cpp:78)


def test_loader_equals_reference(R, rb, need_ref, scene_mid, tmp_path):
    f = str(tmp_path / "s.rle4")
    scene_mid.save(f)
    ref = rb.RefScene.load(f)
    assert ref.nummaps == scene_mid.nummaps
    for m in range(ref.nummaps):
        a, b = ref.level(m), scene_mid.level(m)
        assert a[:3] == b[:3] and np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4])
    # and the reference's writer produces the bytes ours does
    g = str(tmp_path / "r.rle4")
    ref.save(g)
    assert open(f, "rb").read() == open(g, "rb").read()


# cameras for the finite-scene option: inside the grid, at its edge, outside looking in, far outside
CLIP_CAMERAS = [((64.5, -60.0, 64.5), (0.5, 0.8, 0.0)), ((3.25, -40.0, 120.5), (0.3, 2.4, 0.0)),
                ((-50.0, -80.0, -50.0), (0.45, 0.785 + 1.5708, 0.0)), ((300.0, -120.0, 64.0), (0.4, 4.71, 0.0)),
                ((10000.0, -100.0, 10000.0), (0.4, 1.9, 0.0))]


@pytest.mark.parametrize("flags", [1, 2, 3])
def test_render_line_port_equals_reference_with_core_h_options(R, rb, need_ref, scene_small, flags):
    """R/src/core.h:18,22: CLIPREGION (finite scene) and HEIGHT_COLOR, compiled into the reference with -D and
    switched at run time in the port (flags bit 0 / bit 1)."""
    cfg = R.FrameConfig.default(512, 384)
    drew = 0
    for pos, rot in CLIP_CAMERAS + few_cameras(-40.0)[:2]:
        rm = R.RayMap(cfg).get_ray_map(pos, rot)
        orm = oracle_raymap(rb, rm, scene_small)
        ref, _ = rb.ref_render_frame(orm, cfg.render_size, mip_distance=cfg.mip_distance, z_far=cfg.z_far, rays=cfg.rays_casted, flags=flags)
        port, _, cnt = rb.orc_render(orm, cfg.render_size, cfg.rays_casted, cfg.mip_distance, cfg.z_far, flags=flags)
        assert np.array_equal(ref, port), (flags, pos, rot)
        plain, _, _ = rb.orc_render(orm, cfg.render_size, cfg.rays_casted, cfg.mip_distance, cfg.z_far)
        drew += int(cnt["pixels"] > 0 and not np.array_equal(plain, port))
    assert drew >= 3          # the options change the picture on most of these cameras


def test_more_ray_planes_than_buffer_rows(R, rb, need_ref, scene_small):
    """A square window looking steeply down makes map_line_count exceed RAYS_CASTED (2061 > 2048 here); cudaRender
    renders min(count, RAYS_CASTED) ray planes (R/src/Cuda_Main.cu:196) and so do the oracle wrappers."""
    cfg = R.FrameConfig.default(512, 512)
    pos, rot = (812.2, -40.0, 18034.6), (-1.20644718889272, 1.5707963267948966, 0.0)
    rm = R.RayMap(cfg).get_ray_map(pos, rot)
    assert rm.map_line_count > cfg.rays_casted
    ref, port, ids, cnt = _both(R, rb, scene_small, cfg, pos, rot)
    assert ref.shape == port.shape == (cfg.rays_casted, cfg.render_size)
    assert np.array_equal(ref, port) and cnt["pixels"] > 0


def test_reference_cuda_kernel_comparator_builds(rb, have_ref):
    """oracle/_ref/libref_cuda*.so: the reference's own cudaRender compiled for sm_100a from the sources in place
    (oracle/Makefile refcuda); here only that it was built and exports its entry points (it needs a GPU to run:
    tools/ref_cuda_kernel.py, bench.py roofline.reference_kernel_on_this_gpu)."""
    import ctypes as C
    import os
    if not have_ref:
        pytest.skip("oracle/_ref not built (no /root/reference)")
    assert rb.have_ref_cuda()
    for name in ("libref_cuda.so", "libref_cuda_nofma.so"):
        lib = C.CDLL(os.path.join(rb.REF_DIR, name))
        assert lib.refcuda_render_size() == 1024 and lib.refcuda_rays_casted() == 4096      # R/src/core.h:5-6 as shipped
        assert lib.refcuda_sizeof_render() > 896 + 4096 * 20                               # RayMap_GPU + perf[RAYS_CASTED]
        assert hasattr(lib, "refcuda_frame")
