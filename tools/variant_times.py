"""Traversal kernel time per variant (rlerc_set_lanes_per_ray code) on fly-through frames.
usage: python tools/variant_times.py WORKLOAD CODE [CODE ...]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
R = importlib.import_module("rle-based-voxel-raycasting_b200")
workload = sys.argv[1]
codes = [int(a) for a in sys.argv[2:]] or [0]
scene, name, sy = bench.build_scene(R, workload, lambda m: None)
W, H = bench.WORKLOADS[workload][3]
cfg = R.FrameConfig.default(W, H)
r = R.Renderer(0); r.all_to_gpu(scene); r.set_timing(True)
for t in (0, 125, 250, 500, 750):
    pos, rot = bench.path_pose(R, t, 1000, sy, False)
    rm = R.RayMap(cfg).get_ray_map(pos, rot)
    line = "t %3d rays %5d |" % (t, rm.map_line_count)
    for code in codes:
        r.set_lanes_per_ray(code)
        best = 1e9
        for _ in range(4):
            r.render(rm, cfg); r.sync(); best = min(best, r.last_kernel_ms()[0])
        line += " k%d: %.3f |" % (code, best)
    print(line, flush=True)
