"""Kernel time of ONE interleaved ray-plane slice (what one GPU of N does), producer off/on.
usage: python tools/slice_times.py WORKLOAD NRANKS"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
R = importlib.import_module("rle-based-voxel-raycasting_b200")
workload, n = sys.argv[1], int(sys.argv[2])
scene, name, sy = bench.build_scene(R, workload, lambda m: None)
W, H = bench.WORKLOADS[workload][3]
cfg = R.FrameConfig.default(W, H)
r = R.Renderer(0); r.all_to_gpu(scene); r.set_timing(True)
for t in (0, 250, 750):
    pos, rot = bench.path_pose(R, t, 1000, sy, False)
    rm = R.RayMap(cfg).get_ray_map(pos, rot)
    line = "t %3d rays %5d |" % (t, rm.map_line_count)
    for pc in (0, 1):
        r.set_dda_producer(pc)
        for nr in (1, n):
            best = 1e9
            for _ in range(3):
                if nr == 1: r.render(rm, cfg)
                else: r.render_interleaved(rm, cfg, 32, nr, 0)
                r.sync(); best = min(best, r.last_kernel_ms()[0])
            line += " pc%d 1/%d: %.3f ms |" % (pc, nr, best)
    print(line, flush=True)
