"""A/B of several builds of librlerc.so on the same GPU: every build renders the same fly-through frames; the
warped buffers must hash identically (bit-exactness between builds), then per build: single-frame traversal time and
the throughput with four frames in flight (rlerc_frame_submit / wait, what bench.py's e2e measures).
usage: python tools/ab_libs.py WORKLOAD LIB [LIB ...]         (paths relative to the package directory)"""
import hashlib, importlib, json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "rle-based-voxel-raycasting_b200")


def child(workload):
    import numpy as np
    import bench
    R = importlib.import_module("rle-based-voxel-raycasting_b200")
    scene, name, sy = bench.build_scene(R, workload, lambda m: None)
    W, H = bench.WORKLOADS[workload][3]
    cfg = R.FrameConfig.default(W, H)
    r = R.Renderer(0); r.all_to_gpu(scene); r.set_timing(True)
    out = {"hash": [], "ms": []}
    frames = (0, 125, 250, 500, 750)
    for t in frames:
        pos, rot = bench.path_pose(R, t, 1000, sy, False)
        rm = R.RayMap(cfg).get_ray_map(pos, rot)
        best = 1e9
        for _ in range(5):
            r.render(rm, cfg); r.sync(); best = min(best, r.last_kernel_ms()[0])
        out["ms"].append(round(best, 4))
        out["hash"].append(hashlib.sha1(r.read_warp(cfg)[:rm.map_line_count].tobytes()).hexdigest()[:12])
    # four frames in flight through the C ABI with host buffers (bench.py e2e)
    r.set_timing(False)
    K, DEPTH = 300, int(os.environ.get("AB_DEPTH", "4"))
    poses = [bench.path_pose(R, i * 1000 // K, 1000, sy, False) for i in range(K)]
    pins = [R.PinnedBuffer((H, W, 4)) for _ in range(DEPTH)]
    for i in range(8):
        r.frame_wait(r.frame_submit(poses[i][0], poses[i][1], cfg, pins[i % DEPTH].array))
    r.sync()
    best = 1e9
    for _ in range(2):
        t0 = time.perf_counter()
        tickets = []
        for i in range(K):
            if len(tickets) >= DEPTH:
                r.frame_wait(tickets.pop(0))
            tickets.append(r.frame_submit(poses[i][0], poses[i][1], cfg, pins[i % DEPTH].array))
        for tk in tickets:
            r.frame_wait(tk)
        r.sync()
        best = min(best, time.perf_counter() - t0)
    out["e2e_ms_per_frame"] = round(1e3 * best / K, 4)
    out["e2e_mrays"] = round(W * H * K / best / 1e6, 1)
    out["checksum"] = int(pins[(K - 1) % DEPTH].array[::16, ::16].astype(np.uint32).sum())
    print("AB " + json.dumps(out), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[2])
        sys.exit(0)
    workload, libs = sys.argv[1], sys.argv[2:]
    res = {}
    for lib in libs:
        env = dict(os.environ, RLERC_LIB=os.path.join(PKG, lib))
        p = subprocess.run([sys.executable, __file__, "--child", workload], env=env, capture_output=True, text=True)
        line = [l for l in p.stdout.splitlines() if l.startswith("AB ")]
        if not line:
            print(lib, "FAILED", p.stdout[-2000:], p.stderr[-2000:]); continue
        res[lib] = json.loads(line[0][3:])
        print(lib, res[lib], flush=True)
    hashes = {tuple(v["hash"]) for v in res.values()}
    print("bit-identical across builds:", len(hashes) == 1)
