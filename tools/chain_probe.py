"""How much of a frame's traversal time is the serial chain of its longest ray plane?  Renders 1/k of the ray
planes (interleaved, block 1) for growing k: with k large every warp has an SM sub-partition to itself, so the
time that remains is the uncontended critical path.  usage: python tools/chain_probe.py WORKLOAD [DDA_MODE]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
R = importlib.import_module("rle-based-voxel-raycasting_b200")
workload = sys.argv[1]
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
scene, name, sy = bench.build_scene(R, workload, lambda m: None)
W, H = bench.WORKLOADS[workload][3]
cfg = R.FrameConfig.default(W, H)
r = R.Renderer(0); r.all_to_gpu(scene); r.set_timing(True); r.set_dda_mode(mode)
r.set_lanes_per_ray(int(os.environ.get("CHAIN_LANES", "65")))     # 65 = k_traverse_f always (0 would pick k_traverse_p for the small shares)
for t in (0, 250, 750):
    pos, rot = bench.path_pose(R, t, 1000, sy, False)
    rm = R.RayMap(cfg).get_ray_map(pos, rot)
    line = "t %3d rays %5d |" % (t, rm.map_line_count)
    for k in (1, 2, 4, 8, 16, 64):
        best = 1e9
        for _ in range(3):
            if k == 1: r.render(rm, cfg)
            else: r.render_interleaved(rm, cfg, 1, k, 0)
            r.sync(); best = min(best, r.last_kernel_ms()[0])
        line += " 1/%d: %.3f |" % (k, best)
    print(line, flush=True)
