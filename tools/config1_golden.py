"""BASELINE config 1 at full size: one headless 1024x768 frame at the fixed camera of R/src/main.cpp, rendered
(a) by the C++ host program examples/headless (RLE4::load -> all_to_gpu -> get_ray_map -> k_traverse_f -> k_unwarp ->
PPM) and (b) by the reference's own render_line compiled for the host (oracle/_ref, OpenMP) + the oracle's unwarp;
both PPMs and warped buffers are written to OUTDIR and compared.  Test infrastructure (imports oracle/).
usage: python tools/config1_golden.py [OUTDIR]        (Imrodh.rle4 via $RLERC_IMRODH, else synth_imrodh 1024^3)"""
import importlib, json, math, os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from util import oracle_raymap, rgb_parity, sha
from oracle import refbind as rb
R = importlib.import_module("rle-based-voxel-raycasting_b200")

out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "config1")
os.makedirs(out, exist_ok=True)
W, H = 1024, 768
scene, name, sy = bench.build_scene(R, "imrodh768", lambda m: print(m, file=sys.stderr))
imrodh = name == "Imrodh.rle4"
pos = (10000.0, -818.0 if imrodh else -0.15 * sy, 10000.0)               # main.cpp:316-320,344-347
rot = (0.40, 0.30 + math.pi / 2, 0.0)
path = os.path.join(out, "scene.rle4")
scene.save(path)

t0 = time.time()
p = subprocess.run([os.path.join(ROOT, "examples", "headless"), "--scene", path, "--size", str(W), str(H), "--pos"] + [repr(v) for v in pos]
                   + ["--rot"] + [repr(v) for v in rot] + ["--out", os.path.join(out, "gpu")], capture_output=True, text=True)
print(p.stdout, p.stderr, file=sys.stderr)
assert p.returncode == 0
info = dict(l.split() for l in open(os.path.join(out, "gpu.txt")).read().splitlines())
os.remove(path)

cfg = R.FrameConfig.default(W, H)
rm = R.RayMap(cfg).get_ray_map(pos, rot)
lines = min(rm.map_line_count, cfg.rays_casted)
gwarp = np.fromfile(os.path.join(out, "gpu.warp.raw"), dtype=np.uint32).reshape(lines, cfg.render_size)
with open(os.path.join(out, "gpu.ppm"), "rb") as f:
    for _ in range(3): f.readline()
    grgb = np.frombuffer(f.read(), dtype=np.uint8).reshape(H, W, 3)

orm = oracle_raymap(rb, rm, scene)
threads = os.cpu_count() or 1
kind = "reference" if rb.have_ref() else "port"
t0 = time.time()
if rb.have_ref():
    cwarp, _ = rb.ref_render_frame(orm, cfg.render_size, mip_distance=cfg.mip_distance, z_far=cfg.z_far, rays=cfg.rays_casted, threads=threads)
else:
    cwarp, _, _ = rb.orc_render(orm, cfg.render_size, cfg.rays_casted, cfg.mip_distance, cfg.z_far, threads=threads)
cpu_ms = 1e3 * (time.time() - t0)
crgba = rb.orc_unwarp(orm, W, H, cfg.render_size, cfg.rays_casted, cfg.rays_casted_res, cwarp)
with open(os.path.join(out, "cpu.ppm"), "wb") as f:
    f.write(b"P6\n%d %d\n255\n" % (W, H)); f.write(np.ascontiguousarray(crgba[:, :, :3]).tobytes())
cwarp[:lines].tofile(os.path.join(out, "cpu.warp.raw"))

dmax, same = rgb_parity(grgb, crgba[:, :, :3])
rep = {"config": "BASELINE config 1: %s, one headless %dx%d frame, pos %s rot %s" % (name, W, H, list(pos), [round(v, 6) for v in rot]),
       "ray_planes": int(rm.map_line_count), "cpu_arm": "%s render_line on %d host threads + oracle unwarp" % (kind, threads),
       "cpu_traversal_ms": round(cpu_ms, 1), "gpu_traversal_ms": float(info["traverse_ms"]), "gpu_unwarp_ms": float(info["unwarp_ms"]),
       "warped_buffer_identical": bool(np.array_equal(gwarp, cwarp[:lines])), "warped_buffer_sha256": sha(gwarp)[:16],
       "hit_texels": int((gwarp != 0xff8844).sum() - (gwarp == 0).sum()),
       "rgb_max_channel_diff": dmax, "rgb_identical_pixels": round(same, 6), "ppm_sha256": {"gpu": sha(grgb)[:16], "cpu": sha(crgba[:, :, :3])[:16]}}
print(json.dumps(rep))
json.dump(rep, open(os.path.join(out, "report.json"), "w"), indent=1)
assert rep["warped_buffer_identical"] and dmax <= 1 and same >= 0.999
for fn in ("gpu.warp.raw", "cpu.warp.raw"):
    os.remove(os.path.join(out, fn))
