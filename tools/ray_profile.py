"""Where does the longest ray plane spend its time?  Runs k_traverse_f's PROF build (clock64() phase timers per
ray plane) on fly-through frames and prints the phase breakdown of the slowest ray planes and of the whole frame.
usage: python tools/ray_profile.py WORKLOAD [FRAME_T ...]"""
import ctypes as C, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
R = importlib.import_module("rle-based-voxel-raycasting_b200")
import torch
workload = sys.argv[1]
frames = [int(a) for a in sys.argv[2:]] or [0, 250, 750]
scene, name, sy = bench.build_scene(R, workload, lambda m: None)
W, H = bench.WORKLOADS[workload][3]
cfg = R.FrameConfig.default(W, H)
r = R.Renderer(0); r.all_to_gpu(scene); r.set_timing(True)
lib = R.lib()
lib.rlerc_debug_profile_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
names = ["total", "dda", "test+queue", "geom+gather", "C1", "C2", "consume"]
for t in frames:
    pos, rot = bench.path_pose(R, t, 1000, sy, False)
    rm = R.RayMap(cfg).get_ray_map(pos, rot)
    buf = torch.zeros((cfg.rays_casted, 24), dtype=torch.int64, device="cuda")
    for _ in range(2):
        rc = lib.rlerc_debug_profile_rays(r._c, C.byref(rm), C.byref(cfg), C.c_void_p(buf.data_ptr()))
        assert rc == 0, rc
        r.sync()
    ms = r.last_kernel_ms()[0]
    a = buf.cpu().numpy()[:rm.map_line_count]
    steps, batches = a[:, 7] & 0xffffffff, a[:, 7] >> 32
    order = np.argsort(-a[:, 0])
    print("frame %d: %d ray planes, kernel %.3f ms (PROF build); cycles at 1.965 GHz -> ms" % (t, rm.map_line_count, ms))
    print("  %-8s %8s | %s | filter steps, consume batches" % ("ray", "total ms", " ".join("%12s" % n for n in names[1:])))
    for i in list(order[:6]) + list(order[len(order) // 2: len(order) // 2 + 2]):
        tot = a[i, 0]
        print("  %-8d %8.3f | %s | %d, %d" % (i, tot / 1.965e6, " ".join("%5.1f%% %5.0f" % (100.0 * a[i, k] / tot, a[i, k] / max(1, steps[i] if k <= 3 else batches[i])) for k in range(1, 7)), steps[i], batches[i]))
    tot = a[:, 0].sum()
    st = a[:, 8:].sum(axis=0)
    print("  consume batches %d: B0 took %d, B1 took (all or a prefix of) %d, %d needed the event loop (%d iterations); B1 blocked by [flip,longcol,irregular,topattach,window,clipflip,2long] = %s"
          % (st[0], st[1], st[2], st[11], st[3], [int(v) for v in st[4:11]]))
    i = order[0]
    print("  slowest ray plane, cycles per consume batch: B0 %.0f, B1 %.0f, event loop %.0f, S %.0f" % tuple(a[i, 20 + k] / max(1, batches[i]) for k in range(4)))
    print("  all rays: %s | %d, %d" % (" ".join("%5.1f%%" % (100.0 * a[:, k].sum() / tot) for k in range(1, 7)), steps.sum(), batches.sum()))
