"""LOD streaming on the big scenes: residency and upload volume along the fly-through, frames compared with a full replica.
usage: python tools/stream_probe.py [WORKLOAD] [FRAMES] [MARGIN]      (default: shortrun16k 40 512)"""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
R = importlib.import_module("rle-based-voxel-raycasting_b200")
workload = sys.argv[1] if len(sys.argv) > 1 else "shortrun16k"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 40
margin = int(sys.argv[3]) if len(sys.argv) > 3 else 512
R.lib().rlerc_set_host_threads(os.cpu_count() or 1)
scene, name, sy = bench.build_scene(R, workload, lambda m: print(m, file=sys.stderr))
W, H = bench.WORKLOADS[workload][3]
cfg = R.FrameConfig.default(W, H)
full = R.Renderer(0); full.all_to_gpu(scene)
st = R.Renderer(0); st.all_to_gpu_streamed(scene)
rows, bad = [], 0
for i in range(K):
    pos, rot = bench.path_pose(R, i, 1000, sy, False)          # consecutive frames of the 1000-frame path
    t0 = time.perf_counter(); s = st.stream_prepare(pos, rot, cfg, margin); dt = time.perf_counter() - t0
    rm = R.RayMap(cfg).get_ray_map(pos, rot)
    full.render(rm, cfg); st.render(rm, cfg)
    same = bool(np.array_equal(full.read_warp(cfg)[:rm.map_line_count], st.read_warp(cfg)[:rm.map_line_count]))
    bad += 0 if same else 1
    rows.append((s.resident_bytes, s.uploaded_bytes, s.evicted_bytes, dt))
tot = s.total_bytes
out = {"workload": workload, "scene": name, "scene_mb": tot / 1e6, "window": [W, H], "frames": K, "margin_voxels": margin,
       "frames_identical_to_full_replica": K - bad, "resident_mb_first_frame": rows[0][0] / 1e6, "resident_mb_last_frame": rows[-1][0] / 1e6,
       "resident_fraction": rows[-1][0] / tot, "uploaded_mb_first_frame": rows[0][1] / 1e6,
       "uploaded_mb_per_frame_after": sum(r[1] for r in rows[1:]) / max(1, K - 1) / 1e6,
       "prepare_ms_first": 1e3 * rows[0][3], "prepare_ms_after_mean": 1e3 * sum(r[3] for r in rows[1:]) / max(1, K - 1)}
print(json.dumps(out, indent=1))
sys.exit(1 if bad else 0)
