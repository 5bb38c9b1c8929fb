"""Render one camera a few times (for ncu / timing). Usage:
python tools/profile_case.py SIZE W H CAM LANES [REPS]"""
import importlib
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
R = importlib.import_module("rle-based-voxel-raycasting_b200")

CAMS = [(-0.8, (0.40, 0.30)), (-0.3, (0.05, 2.0)), (-1.5, (1.2, 4.5)), (-0.5, (-0.6, 0.8))]


def main():
    size, W, H, cam, lanes = (int(a) for a in sys.argv[1:6])
    reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
    scene = R.RLE4.synth(0, size, size, size, seed=1)
    cfg = R.FrameConfig.default(W, H)
    r = R.Renderer(0)
    r.all_to_gpu(scene)
    r.set_timing(True)
    r.set_lanes_per_ray(lanes)
    hy, (pitch, yaw) = CAMS[cam]
    pos, rot = (10000., hy * size, 10000.), (pitch, yaw + math.pi / 2, 0.)
    rm = R.RayMap(cfg).get_ray_map(pos, rot)
    ts = []
    for _ in range(reps):
        r.render(rm, cfg)
        r.sync()
        ts.append(r.last_kernel_ms()[0])
    print("rays", rm.map_line_count, "traverse ms", ["%.3f" % t for t in ts])
    import torch
    ids = torch.full((cfg.rays_casted, cfg.render_size, 2), 0xffffffff, dtype=torch.int64, device="cuda").to(torch.int32)
    r.render_ids(rm, cfg, ids.data_ptr())
    r.sync()
    c = r.counters()
    names = ["elems_total", "elems_processed", "voxels_processed", "elems_rendered", "pixels", "cols_fetched", "run_iters", "cols_nonempty", "cleared", "dda_steps"]
    print({k: v for k, v in zip(names, c)})


if __name__ == "__main__":
    main()
