"""How many consume batches take the B0 fast path and why the others were refused (debug counters of the instrumented
build).  usage: python tools/fastpath_stats.py WORKLOAD"""
import ctypes as C, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
R = importlib.import_module("rle-based-voxel-raycasting_b200")
import torch
workload = sys.argv[1]
scene, name, sy = bench.build_scene(R, workload, lambda m: None)
W, H = bench.WORKLOADS[workload][3]
cfg = R.FrameConfig.default(W, H)
r = R.Renderer(0); r.all_to_gpu(scene)
ids = torch.empty((cfg.rays_casted, cfg.render_size, 2), dtype=torch.int32, device="cuda")
lib = R.lib(); lib.rlerc_debug_counters.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
for t in (0, 125, 250, 750):
    pos, rot = bench.path_pose(R, t, 1000, sy, False)
    rm = R.RayMap(cfg).get_ray_map(pos, rot)
    r.render_ids(rm, cfg, ids.data_ptr()); r.sync()
    out = (C.c_uint64 * 32)(); lib.rlerc_debug_counters(r._c, out)
    o = list(out)
    print("t", t, "batches", o[10], "attempts", o[11], "success", o[12], "why[culled-flip,long,notbottom,top,minbefore,coopmin,laterrun,clipflip]", o[16:24])
