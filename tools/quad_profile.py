"""k_traverse_q: where do the four warps of the slowest ray planes spend their time?  Needs a library built with
-DRLERC_Q_PROF=1 (tools/gpu/build_ab.sh q_prof "-DRLERC_Q_PROF=1"; RLERC_LIB=.../librlerc_q_prof.so).
usage: RLERC_PROF_Q=1 [RLERC_PROF_SLICE=16] python tools/quad_profile.py WORKLOAD [FRAME_T ...]"""
import ctypes as C, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("RLERC_PROF_Q", "1")
import bench
R = importlib.import_module("rle-based-voxel-raycasting_b200")
import torch
workload = sys.argv[1]
frames = [int(a) for a in sys.argv[2:]] or [0, 750]
scene, name, sy = bench.build_scene(R, workload, lambda m: None)
W, H = bench.WORKLOADS[workload][3]
cfg = R.FrameConfig.default(W, H)
r = R.Renderer(0); r.all_to_gpu(scene); r.set_timing(True)
lib = R.lib()
lib.rlerc_debug_profile_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
for t in frames:
    pos, rot = bench.path_pose(R, t, 1000, sy, False)
    rm = R.RayMap(cfg).get_ray_map(pos, rot)
    buf = torch.zeros((cfg.rays_casted, 24), dtype=torch.int64, device="cuda")
    for _ in range(2):
        assert lib.rlerc_debug_profile_rays(r._c, C.byref(rm), C.byref(cfg), C.c_void_p(buf.data_ptr())) == 0
        r.sync()
    ms = r.last_kernel_ms()[0]
    a = buf.cpu().numpy()[:rm.map_line_count].reshape(-1, 6, 4)
    tot = a[:, :4, 0].max(axis=1)
    order = np.argsort(-tot)
    print("frame %d: %d ray planes (slice 1/%s), kernel %.3f ms, %s; per role: total ms (busy ms = total - waiting), items, busy cycles per item"
          % (t, rm.map_line_count, os.environ.get("RLERC_PROF_SLICE", "1"), ms, r.last_kernel))
    for i in order[:5]:
        line = "  ray %5d |" % i
        for k, nm in enumerate("FPRS"):
            T, Wt, n = a[i, k, 0], a[i, k, 1], max(1, a[i, k, 2])
            line += " %s %.3f (%.3f) %4d x %5.0f |" % (nm, T / 1.965e6, (T - Wt) / 1.965e6, n, (T - Wt) / n)
        print(line, flush=True)
