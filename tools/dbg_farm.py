"""Debug helper for the frame-parallel multi-GPU mode (FrameFarm): per-round timings of render / gather on every rank.
usage: python -m torch.distributed.run --nproc-per-node N tools/dbg_farm.py"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch, torch.distributed as dist
R = importlib.import_module("rle-based-voxel-raycasting_b200")
MG = importlib.import_module("rle-based-voxel-raycasting_b200.multigpu")
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
R.lib().rlerc_set_host_threads(8)
scene, name, sy = bench.build_scene(R, "imrodh1080p", lambda m: None)
cfg = R.FrameConfig.default(1920, 1080)
r = R.Renderer(local); r.all_to_gpu(scene); r.set_timing(True)
farm = MG.FrameFarm(r, cfg, torch, rank, world, dist)
K = 40
rms = [R.RayMap(cfg).get_ray_map(*bench.path_pose(R, i, K, sy)) for i in range(K)]
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for mode in ("plain", "flush", "flush+sync"):
    dist.barrier(); torch.cuda.synchronize()
    evs = []; t0 = time.perf_counter()
    for rd in range(K // world):
        i = rd * world + rank
        if mode != "plain": flush.fill_(rd & 255)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); farm.render_round(rd, rms[i]); e1.record()
        if mode == "flush+sync": e1.synchronize()
        evs.append((e0, e1))
    farm.finish(); torch.cuda.synchronize(); wall = time.perf_counter() - t0
    print("rank", rank, mode, "wall %.1f ms" % (wall * 1e3), "event sum %.1f ms" % sum(a.elapsed_time(b) for a, b in evs), "per round", ["%.2f" % a.elapsed_time(b) for a, b in evs[:6]], flush=True)
dist.destroy_process_group()
