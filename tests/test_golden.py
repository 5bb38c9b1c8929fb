"""The oracle port against the committed golden vectors (generated from the compiled reference by
tests/golden/make_golden.py).  No GPU, no /root/reference."""
import json
import os

import numpy as np
import pytest

from util import oracle_raymap, sha

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden.json")))


@pytest.mark.parametrize("name", sorted(GOLD["cases"]))
def test_port_reproduces_golden(R, rb, name):
    case = GOLD["cases"][name]
    n = case["size"]
    scene = R.RLE4.synth(case["kind"], n, n, n, seed=case["seed"])
    assert [sha(scene.level(m)[4]) for m in range(scene.nummaps)] == case["scene_sha"]
    W, H = case["window"]
    cfg = R.FrameConfig.default(W, H)
    for fr in case["frames"]:
        rm = R.RayMap(cfg).get_ray_map(fr["pos"], fr["rot"])
        assert sha(np.frombuffer(bytes(rm), np.uint8)) == fr["raymap_sha"]
        assert rm.map_line_count == fr["rays"]
        orm = oracle_raymap(rb, rm, scene)
        warp, ids, cnt = rb.orc_render(orm, cfg.render_size, cfg.rays_casted, cfg.mip_distance, cfg.z_far, want_ids=True)
        assert sha(warp) == fr["warp_sha"]
        assert sha(ids) == fr["ids_sha"]
        assert cnt["pixels"] == fr["pixels"]
        rgba = rb.orc_unwarp(orm, W, H, cfg.render_size, cfg.rays_casted, cfg.rays_casted_res, warp)
        assert sha(rgba) == fr["rgba_sha"]
        aa = rb.orc_unwarp(orm, W, H, cfg.render_size, cfg.rays_casted, cfg.rays_casted_res, warp, shader=1)
        assert sha(aa) == fr["rgba_2xaa_sha"]            # colorize_buddha_soft_2xAA.frag restatement


def test_raw_golden_rows(R, rb):
    gold = np.load(os.path.join(HERE, "golden", "terrain64_frame0_rays0_64.npy"))
    case = GOLD["cases"]["terrain64"]
    scene = R.RLE4.synth(case["kind"], 64, 64, 64, seed=case["seed"])
    cfg = R.FrameConfig.default(*case["window"])
    fr = case["frames"][0]
    rm = R.RayMap(cfg).get_ray_map(fr["pos"], fr["rot"])
    warp, _, _ = rb.orc_render(oracle_raymap(rb, rm, scene), cfg.render_size, cfg.rays_casted, cfg.mip_distance, cfg.z_far, ray_end=64)
    assert np.array_equal(warp[:64], gold)
    # layout of a texel: attr16 | depth16<<16, depth even, sky sentinel 0xff8844 (Cuda_Render.h:257,708,723)
    hit = gold[(gold != 0) & (gold != 0xff8844)]
    assert hit.size > 0 and np.all(((hit >> 16) & 1) == 0)


def test_soft_pass_oracle_properties(rb):
    """soft.frag restatement: sky (alpha 0 -> rad 0) is left alone; a flat opaque image stays flat away from the
    top/right window edges (beyond them the reference's FBO was never rendered: read as 0)."""
    sky = np.zeros((96, 128, 4), np.uint8)
    sky[..., :3] = (178, 204, 255)
    assert np.array_equal(rb.orc_soft(sky), sky)
    flat = np.zeros((300, 400, 4), np.uint8)
    flat[...] = (90, 120, 150, 40)
    out = rb.orc_soft(flat)
    assert np.all(np.abs(out[40:, :360].astype(int) - flat[40:, :360].astype(int)) <= 1)
    assert not np.array_equal(out[:8], flat[:8])              # top edge mixes in the unrendered FBO


def test_unwarp_shading_independent_restatement(R, rb, scene_small):
    """A second, independent restatement of the Buddha shading (colorize_buddha_soft.frag:89-136, vectorised numpy
    float32, written from the shader text) agrees with the oracle's orc_unwarp on every pixel within 1 LSB (the
    tolerance the north star states for the final RGB; the GLSL pass has no executable reference here)."""
    from util import few_cameras
    f = np.float32
    cfg = R.FrameConfig.default(320, 240)
    for pos, rot in few_cameras(-40.0)[:3]:
        rm = R.RayMap(cfg).get_ray_map(pos, rot)
        orm = oracle_raymap(rb, rm, scene_small)
        warp, _, _ = rb.orc_render(orm, cfg.render_size, cfg.rays_casted, cfg.mip_distance, cfg.z_far)
        rgba, tex = rb.orc_unwarp(orm, cfg.width, cfg.height, cfg.render_size, cfg.rays_casted, cfg.rays_casted_res, warp, want_texels=True)
        t = warp[tex[..., 0], tex[..., 1]]
        c = [((t >> s) & 255).astype(f) / f(255) for s in (0, 8, 16, 24)]           # texture2D of an RGBA8 texel: r g b a
        sky = c[2] == f(1)                                                          # frag:89 c.z != 1.0
        with np.errstate(all="ignore"):
            z = c[2] * f(1.0 / 256.0) + c[3]                                        # frag:90
            fragz = f(0.001) / z                                                    # frag:135
            light = (f(1) - c[1]) * f(1) + (f(0) + c[0]) * f(0.3) - f(0.5)          # frag:113
            pw = f(1.2) * np.power(np.maximum(light, f(0)), f(4))                   # frag:121
            rgb = [light * k + pw * f(1.2) for k in (f(1.3), f(0.9), f(0.7))]
        out = np.zeros(rgba.shape, np.int32)
        for k in range(3):
            v = np.where(sky, f((178, 204, 255)[k]) / f(255), rgb[k])
            out[..., k] = (np.clip(v, 0, 1) * f(255) + f(0.5)).astype(np.int32)
        out[..., 3] = (np.clip(np.where(sky, f(0), fragz), 0, 1) * f(255) + f(0.5)).astype(np.int32)
        d = np.abs(out - rgba.astype(np.int32))
        assert d.max() <= 1, (rot, int(d.max()))
        assert (d.reshape(-1, 4).max(axis=1) == 0).mean() >= 0.999
        assert (~sky).any()                                                         # the frame has shaded (non-sky) pixels
