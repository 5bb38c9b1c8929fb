"""Small frames through every traversal variant, for compute-sanitizer (memcheck / racecheck / initcheck).
usage: compute-sanitizer --tool memcheck python tools/sanitize_case.py [KERNEL_CODE ...]"""
import importlib, math, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
R = importlib.import_module("rle-based-voxel-raycasting_b200")
scene = R.RLE4.synth(0, 128, 128, 128, seed=3)
runs = R.RLE4.synth(1, 64, 64, 64, seed=42)
r = R.Renderer(0)
for sc, h in ((scene, -70.0), (runs, -50.0)):
    r.all_to_gpu(sc)
    for wh in ((320, 240), (1920, 1080)):
        cfg = R.FrameConfig.default(*wh)
        for code in ([int(a) for a in sys.argv[1:]] or (0, 64, 67, 32)):
            r.set_lanes_per_ray(code)
            for rot in ((0.4, 0.3 + math.pi / 2, 0.0), (-0.6, 2.0, 0.0), (1.2, 4.5, 0.0)):
                rm = R.RayMap(cfg).get_ray_map((10000.0, h, 10000.0), rot)
                host = np.zeros((cfg.height, cfg.width, 4), np.uint8)
                r.render(rm, cfg); r.unwarp(rm, cfg); r.sync()
        print("ok", wh, flush=True)
r.set_lanes_per_ray(0)
r.close()
