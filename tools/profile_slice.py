"""Render one interleaved ray-plane slice of a bench frame a few times (for ncu / timing of the slice mode).
usage: python tools/profile_slice.py WORKLOAD FRAME_T NRANKS [LANES]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
R = importlib.import_module("rle-based-voxel-raycasting_b200")
workload, t, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
scene, name, sy = bench.build_scene(R, workload, lambda m: None)
W, H = bench.WORKLOADS[workload][3]
cfg = R.FrameConfig.default(W, H)
r = R.Renderer(0); r.all_to_gpu(scene); r.set_timing(True)
if len(sys.argv) > 4: r.set_lanes_per_ray(int(sys.argv[4]))
pos, rot = bench.path_pose(R, t, 1000, sy, False)
rm = R.RayMap(cfg).get_ray_map(pos, rot)
for _ in range(3):
    r.render_interleaved(rm, cfg, 32, n, 0); r.sync()
    print(r.last_kernel_ms()[0])
