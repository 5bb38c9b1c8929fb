"""k_traverse_p (variant 68: filter warp + consume warp per ray plane) against the production kernel: identical
warped buffers on small scenes and on the bench frames, single-frame traversal time, time with 1/k of the ray planes
(uncontended chain).  usage: python tools/pair_probe.py [WORKLOAD]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from util import few_cameras, camera_grid, edge_scenes, edge_cameras
R = importlib.import_module("rle-based-voxel-raycasting_b200")
r = R.Renderer(0); r.set_timing(True)
bad = 0
def same(scene, cfg, pos, rot, tag):
    global bad
    rm = R.RayMap(cfg).get_ray_map(pos, rot)
    out = []
    for code in (65, 68, 69):
        r.set_lanes_per_ray(code)
        r.upload(r.warp_buffer(cfg), np.zeros((cfg.rays_casted, cfg.render_size), np.uint32))
        r.render(rm, cfg); r.sync()
        out.append(r.read_warp(cfg).copy())
    if not np.array_equal(out[0], out[1]):
        bad += 1
        print("MISMATCH", tag, pos, rot, int((out[0] != out[1]).sum()), flush=True)
scene = R.RLE4.synth(0, 128, 128, 128, seed=1); r.all_to_gpu(scene)
cfg = R.FrameConfig.default(640, 480)
n = 0
for pos, rot in list(camera_grid(-100.0)) + few_cameras(-100.0):
    same(scene, cfg, pos, rot, "terrain128"); n += 1
for name, sc in edge_scenes(R).items():
    r.all_to_gpu(sc)
    for pos, rot in edge_cameras():
        same(sc, R.FrameConfig.default(400, 300), pos, rot, name); n += 1
print("parity cases", n, "mismatches", bad, flush=True)
workload = sys.argv[1] if len(sys.argv) > 1 else "imrodh1080p"
scene, name, sy = bench.build_scene(R, workload, lambda m: None)
W, H = bench.WORKLOADS[workload][3]
cfg = R.FrameConfig.default(W, H)
r.all_to_gpu(scene)
for t in (0, 250, 750):
    pos, rot = bench.path_pose(R, t, 1000, sy, False)
    same(scene, cfg, pos, rot, workload)
    rm = R.RayMap(cfg).get_ray_map(pos, rot)
    line = "t %3d rays %5d |" % (t, rm.map_line_count)
    for code in [int(c) for c in os.environ.get("PAIR_CODES", "65,68,69").split(",")]:
        r.set_lanes_per_ray(code)
        for k in (1, 2, 4, 8, 16):
            best = 1e9
            for _ in range(4):
                if k == 1: r.render(rm, cfg)
                else: r.render_interleaved(rm, cfg, 1, k, 0)
                r.sync(); best = min(best, r.last_kernel_ms()[0])
            line += " k%d 1/%d: %.3f |" % (code, k, best)
    print(line, flush=True)
print("mismatches", bad)
