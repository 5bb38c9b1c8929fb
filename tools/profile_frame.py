"""Render fly-through frames of a bench workload a few times (for ncu / timing).
usage: python tools/profile_frame.py WORKLOAD FRAME_T [LANES] [REPS]"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

R = importlib.import_module("rle-based-voxel-raycasting_b200")


def main():
    workload, t = sys.argv[1], int(sys.argv[2])
    lanes = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    scene, name, sy = bench.build_scene(R, workload, lambda m: print(m, file=sys.stderr))
    W, H = bench.WORKLOADS[workload][3]
    cfg = R.FrameConfig.default(W, H)
    r = R.Renderer(0)
    r.all_to_gpu(scene)
    r.set_timing(True)
    r.set_lanes_per_ray(lanes)
    pos, rot = bench.path_pose(R, t, 1000, sy, name == "Imrodh.rle4")
    rm = R.RayMap(cfg).get_ray_map(pos, rot)
    ts = []
    for _ in range(reps):
        r.render(rm, cfg)
        r.unwarp(rm, cfg)
        r.sync()
        ts.append(r.last_kernel_ms())
    print("frame", t, "rays", rm.map_line_count, "ms (traverse, unwarp):", [("%.3f" % a, "%.3f" % b) for a, b in ts])


if __name__ == "__main__":
    main()
