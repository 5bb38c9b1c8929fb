"""L2 and HBM read bandwidth of this GPU, measured with a plain 128-bit-load kernel (k_read_u4): the denominators of
bench.py's roofline block.  Writes profiles/l2_peak.json.  usage: python tools/l2_peak.py"""
import ctypes as C, importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
R = importlib.import_module("rle-based-voxel-raycasting_b200")
r = R.Renderer(0)
f = R.lib().rlerc_debug_read_gbs
f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
out = {}
for mb, iters in ((16, 400), (32, 200), (48, 150), (64, 100), (96, 60), (128, 40), (256, 20), (2048, 4)):
    best = 0.0
    for _ in range(3):
        g = C.c_double()
        assert f(r._c, mb, iters, C.byref(g)) == 0
        best = max(best, g.value)
    out["%d MB" % mb] = round(best, 1)
    print("%5d MB x %3d passes: %8.1f GB/s" % (mb, iters, best), flush=True)
res = {"l2_read_gbs": max(out[k] for k in ("16 MB", "32 MB", "48 MB")), "hbm_read_gbs": out["2048 MB"], "by_working_set": out,
       "how": "k_read_u4 (csrc/kernels.cu): 592 blocks x 512 threads, 128-bit ld.global.cg, best of 3; working sets <= 48 MB stay in the 126 MB L2"}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "l2_peak.json"), "w"), indent=1)
print(json.dumps(res))
