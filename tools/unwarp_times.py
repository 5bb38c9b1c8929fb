"""k_unwarp alone: ms per launch and GB/s (8 bytes per pixel) at the BASELINE windows, single GPU, texels from a
rendered frame of the 256^3 test scene.  RLERC_LIB selects an alternative build (tools/gpu/build_ab.sh).
usage: python tools/unwarp_times.py"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
R = importlib.import_module("rle-based-voxel-raycasting_b200")
scene, name, sy = bench.build_scene(R, "small", lambda m: None)
r = R.Renderer(0); r.all_to_gpu(scene); r.set_timing(True)
line = os.path.basename(R.LIB_PATH) + ":"
for W, H in ((1024, 768), (1920, 1080), (3840, 2160), (7680, 4320)):
    cfg = R.FrameConfig.default(W, H)
    pos, rot = bench.path_pose(R, 0, 1000, sy, False)
    rm = R.RayMap(cfg).get_ray_map(pos, rot)
    r.render(rm, cfg); r.sync()
    best = 1e9
    for _ in range(8):
        r.unwarp(rm, cfg); r.sync(); best = min(best, r.last_kernel_ms()[1])
    line += "  %dx%d %.4f ms %.0f GB/s |" % (W, H, best, 8.0 * W * H / best / 1e6)
print(line, flush=True)
