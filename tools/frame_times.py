"""Per-frame traversal time + work counters along the fly-through. usage: python tools/frame_times.py WORKLOAD STRIDE [LANES]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
R = importlib.import_module("rle-based-voxel-raycasting_b200")
import torch
workload, stride = sys.argv[1], int(sys.argv[2])
lanes = int(sys.argv[3]) if len(sys.argv) > 3 else 0
scene, name, sy = bench.build_scene(R, workload, lambda m: None)
W, H = bench.WORKLOADS[workload][3]
cfg = R.FrameConfig.default(W, H)
r = R.Renderer(0); r.all_to_gpu(scene); r.set_timing(True); r.set_lanes_per_ray(lanes)
if len(sys.argv) > 4: r.set_dda_producer(int(sys.argv[4]))
if len(sys.argv) > 5: r.set_dda_mode(int(sys.argv[5]))
ids = torch.empty((cfg.rays_casted, cfg.render_size, 2), dtype=torch.int32, device="cuda")
names = ["elems_total", "elems_processed", "voxels_processed", "elems_rendered", "pixels", "cols_fetched", "run_iters", "cols_nonempty", "cleared", "dda_steps"]
for t in range(0, 1000, stride):
    pos, rot = bench.path_pose(R, t, 1000, sy, name == "Imrodh.rle4")
    rm = R.RayMap(cfg).get_ray_map(pos, rot)
    best = 1e9
    for _ in range(3):
        r.render(rm, cfg); r.sync(); best = min(best, r.last_kernel_ms()[0])
    r.render_ids(rm, cfg, ids.data_ptr()); r.sync()
    c = dict(zip(names, r.counters()))
    print("t %4d pitch %.2f rays %5d  %.3f ms | steps %.1fM cols %.1fM runs %.1fM events %.0fk pix %.2fM" % (
        t, rot[0], rm.map_line_count, best, c["dda_steps"] / 1e6, c["cols_fetched"] / 1e6, c["run_iters"] / 1e6, c["elems_rendered"] / 1e3, c["pixels"] / 1e6), flush=True)
