"""Device time of every kernel of a few traversal launches (run under ncu --metrics gpu__time_duration.sum): full frames and
1/16 of the ray planes.  usage: python tools/kernel_times.py WORKLOAD [LANES]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
R = importlib.import_module("rle-based-voxel-raycasting_b200")
workload = sys.argv[1] if len(sys.argv) > 1 else "imrodh1080p"
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 65
scene, name, sy = bench.build_scene(R, workload, lambda m: None)
W, H = bench.WORKLOADS[workload][3]
cfg = R.FrameConfig.default(W, H)
r = R.Renderer(0); r.all_to_gpu(scene); r.set_lanes_per_ray(lanes)
for t in (0, 750):
    pos, rot = bench.path_pose(R, t, 1000, sy, False)
    rm = R.RayMap(cfg).get_ray_map(pos, rot)
    for k in (1, 16):
        for _ in range(2):
            if k == 1: r.render(rm, cfg)
            else: r.render_interleaved(rm, cfg, 1, k, 0)
            r.sync()
