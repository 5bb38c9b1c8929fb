"""Multi-GPU groups (csrc/group.cu) under torchrun, one process per GPU: the composited frame equals rank 0's own
single-GPU frame, then frames/s with `depth` frames in flight and one frame at a time.
usage: torchrun --nproc-per-node N tools/group_probe.py [WORKLOAD] [FRAMES] [DEPTH]      (N = 1 works too)"""
import ctypes as C, importlib, os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", os.environ.get("PROBE_CONNECTIONS", "32"))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import bench
R = importlib.import_module("rle-based-voxel-raycasting_b200")
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
workload = sys.argv[1] if len(sys.argv) > 1 else "small"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 24
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 4
import traceback
def _excepthook(t, v, tb):
    print("rank %d FAILED: %s" % (rank, "".join(traceback.format_exception(t, v, tb))[-1500:]), flush=True)
    sys.__excepthook__(t, v, tb)
sys.excepthook = _excepthook
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
R.lib().rlerc_set_host_threads(max(1, (os.cpu_count() or 1) // world))
scene, name, sy = bench.build_scene(R, workload, lambda m: None)
W, H = bench.WORKLOADS[workload][3]
cfg = R.FrameConfig.default(W, H)
r = R.Renderer(local); r.all_to_gpu(scene)
g = R.Group(r, cfg, rank, world, depth=depth)
if world > 1:
    g.connect_distributed(torch, dist)
maps = []
for i in range(K):
    p, q = bench.path_pose(R, i, K, sy, False)
    rm = R.RayMapGPU(); C.memmove(C.byref(rm), C.byref(R.RayMap(cfg).get_ray_map(p, q)), 896); maps.append(rm)
def barrier():
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
# ---- parity: frame complete on rank 0 (pushed over NVLink) and as bands in a host frame
bad = 0
for i in (0, K // 3, (2 * K) // 3):
    t = g.submit(maps[i], 0); g.wait(t); barrier()
    if rank == 0:
        ptr, _, _ = g.image(t)
        got = r.download(ptr, (H, W, 4), np.uint8)
        solo = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
        r.frame_device(maps[i], cfg, 1, 1, 0, solo.data_ptr()); r.sync()
        if not np.array_equal(got, solo.cpu().numpy()):
            bad += 1; print("MISMATCH frame", i, int((got != solo.cpu().numpy()).sum()), flush=True)
    barrier()
    t = g.submit(maps[i], -1); g.wait(t); barrier()
    ptr, a, b = g.image(t)
    band = r.download(ptr, (H, W, 4), np.uint8)[a:b]
    solo = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
    r.frame_device(maps[i], cfg, 1, 1, 0, solo.data_ptr()); r.sync()
    if not np.array_equal(band, solo.cpu().numpy()[a:b]):
        bad += 1; print("MISMATCH band, rank", rank, "frame", i, flush=True)
    barrier()
# ---- throughput
def run(n_rep, dst):
    barrier(); t0 = time.perf_counter()
    for rep in range(n_rep):
        for i in range(K): last = g.submit(maps[i], dst)
    g.sync(); barrier()
    return (time.perf_counter() - t0) / (n_rep * K)
run(1, 0)
reps_t = torch.tensor([max(1, int(0.5 / max(run(1, 0) * K, 1e-4)))], device="cuda")
if world > 1: dist.all_reduce(reps_t, op=dist.ReduceOp.MAX)       # every member has to submit the SAME frames
reps = int(reps_t.cpu()[0])
ms_push = 1e3 * run(reps, 0); ms_band = 1e3 * run(reps, -1); ms_rot = 1e3 * run(reps, -2)
# ---- parity again after many generations of every slot (stale data in any cache would show here)
for i in (1, K - 1):
    t = g.submit(maps[i], 0); g.wait(t); barrier()
    if rank == 0:
        ptr, _, _ = g.image(t)
        got = r.download(ptr, (H, W, 4), np.uint8)
        solo = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
        r.frame_device(maps[i], cfg, 1, 1, 0, solo.data_ptr()); r.sync()
        if not np.array_equal(got, solo.cpu().numpy()):
            bad += 1; print("MISMATCH (late) frame", i, int((got != solo.cpu().numpy()).sum()), flush=True)
    barrier()
# ---- one frame at a time
r.set_timing(True); lat = []; tr = []; un = []
for i in range(K):
    barrier(); t = g.submit(maps[i], 0); g.wait(t); lat.append(g.last_ms(t))
    a, b = r.last_kernel_ms(); tr.append(a); un.append(b)
r.set_timing(False)
print("  rank %d: traversal %.3f ms, unwarp %.3f ms, whole frame incl. barriers %.3f ms (one at a time, mean of %d)" % (rank, sum(tr) / K, sum(un) / K, sum(lat) / K, K), flush=True)
lat_t = torch.tensor([sum(lat) / K], device="cuda")
if world > 1: dist.all_reduce(lat_t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("group_probe %s N=%d depth=%d: mismatches %d | pipelined %.3f ms/frame to rank 0, %.3f ms/frame as bands, %.3f rotating the destination | one at a time %.3f ms/frame | kernel %s"
          % (workload, world, depth, bad, ms_push, ms_band, ms_rot, float(lat_t.cpu()[0]), r.last_kernel), flush=True)
g.close(); r.close()
if world > 1: dist.destroy_process_group()
sys.exit(1 if bad else 0)
