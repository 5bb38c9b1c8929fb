"""Generates tests/golden/golden.json.  Run HERE (the container with /root/reference mounted):

    python tests/golden/make_golden.py

Warped-buffer and ray-map hashes come from the REFERENCE compiled as host C++ (oracle/_ref);
the scene bytes from the product compressor after it has been checked against the reference's
(tests/test_oracle_vs_ref.py); the unwarp hashes from the oracle port (the GLSL pass has no
executable reference).  The fixture travels to the GPU box, /root/reference does not.
"""
import ctypes as C
import importlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
from oracle import refbind as rb  # noqa: E402
from util import few_cameras, levels_of, oracle_raymap, sha  # noqa: E402

R = importlib.import_module("rle-based-voxel-raycasting_b200")

CASES = [("terrain64", 0, 64, 1, -40.0, (256, 192)),
         ("terrain128", 0, 128, 1, -100.0, (640, 480)),
         ("shortruns128", 1, 128, 42, -90.0, (512, 384))]


def main():
    assert rb.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    out = {"generator": "tests/golden/make_golden.py", "source": "oracle/_ref (reference compiled as host C++)", "cases": {}}
    for name, kind, n, seed, h, (W, H) in CASES:
        scene = R.RLE4.synth(kind, n, n, n, seed=seed)
        cfg = R.FrameConfig.default(W, H)
        case = {"kind": kind, "size": n, "seed": seed, "window": [W, H],
                "scene_sha": [sha(scene.level(m)[4]) for m in range(scene.nummaps)], "frames": []}
        for pos, rot in few_cameras(h):
            ref_rm = rb.ref_get_ray_map(pos, rot, cfg.border, cfg.rays_casted_res)
            rm = R.RayMap(cfg).get_ray_map(pos, rot)
            assert bytes(ref_rm) == bytes(rm)
            orm = oracle_raymap(rb, rm, scene)
            warp, _ = rb.ref_render_frame(orm, cfg.render_size, mip_distance=cfg.mip_distance, z_far=cfg.z_far, rays=cfg.rays_casted)
            port, ids, cnt = rb.orc_render(orm, cfg.render_size, cfg.rays_casted, cfg.mip_distance, cfg.z_far, want_ids=True)
            assert np.array_equal(warp, port)
            rgba = rb.orc_unwarp(orm, W, H, cfg.render_size, cfg.rays_casted, cfg.rays_casted_res, warp)
            rgba_2xaa = rb.orc_unwarp(orm, W, H, cfg.render_size, cfg.rays_casted, cfg.rays_casted_res, warp, shader=1)
            case["frames"].append({"pos": list(pos), "rot": list(rot), "raymap_sha": sha(np.frombuffer(bytes(ref_rm), np.uint8)),
                                   "rays": ref_rm.map_line_count, "warp_sha": sha(warp), "ids_sha": sha(ids),
                                   "rgba_sha": sha(rgba), "rgba_2xaa_sha": sha(rgba_2xaa), "pixels": cnt["pixels"]})
        out["cases"][name] = case
    # one tiny raw vector: the first frame of the smallest case, rays 0..3
    name, kind, n, seed, h, (W, H) = CASES[0]
    scene = R.RLE4.synth(kind, n, n, n, seed=seed)
    cfg = R.FrameConfig.default(W, H)
    pos, rot = few_cameras(h)[0]
    rm = R.RayMap(cfg).get_ray_map(pos, rot)
    warp, _ = rb.ref_render_frame(oracle_raymap(rb, rm, scene), cfg.render_size, mip_distance=cfg.mip_distance, z_far=cfg.z_far, rays=cfg.rays_casted)
    np.save(os.path.join(HERE, "terrain64_frame0_rays0_64.npy"), warp[:64])
    json.dump(out, open(os.path.join(HERE, "golden.json"), "w"), indent=1)
    print("wrote golden.json with", sum(len(c["frames"]) for c in out["cases"].values()), "frames")


if __name__ == "__main__":
    main()
