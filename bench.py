#!/usr/bin/env python
"""Benchmark of the RLE voxel raycaster frame loop (BASELINE.json metric: Mrays/s & frames/s).

    python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference]

A step is ONE FRAME of the hot path (frame setup -> traversal kernel -> unwarp kernel [-> NCCL
compositing for N > 1]) on the scripted fly-through of BASELINE config 2; step i of K renders path
frame i*1000/K, so any K covers the whole orbit.  One process per GPU (torchrun sets RANK/...).

value      pixel-rays/s = W*H*frames/s, scene + buffers resident in HBM, device-timed (CUDA events
           on the launching stream, summed over the K steps, max over ranks).  L2 is flushed between
           timed steps by writing a 512 MiB buffer (outside the timed intervals).
e2e        the same frames through the C ABI with HOST buffers: camera pose in, RGBA frame out in
           pinned host memory, D2H inside the timed region (pipelined submit/wait, wall clock).
roofline   traversal kernel: algorithmic bytes per frame (8C + 2(E-C1) + 6P + 4K, DESIGN.md §5, counted
           by the instrumented kernel on 8 path frames) / its CUDA-event duration, vs the measured HBM peak.
cpu_baseline / --impl reference: the reference's own render_line compiled for the host (oracle/_ref,
           OpenMP over ray planes) + the oracle's unwarp, on a bounded sample of the same frames.
"""
import argparse
import ctypes as C
import importlib
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (scene kind, base size (sx, sy, sz), tiling, window, description)
    "imrodh1080p": (0, (1024, 1024, 1024), 1, (1920, 1080), "BASELINE config 2: Imrodh.rle4 (or synth_imrodh 1024^3 seed 1) at 1920x1080, scripted fly-through"),
    "imrodh768": (0, (1024, 1024, 1024), 1, (1024, 768), "BASELINE config 1 window: 1024x768"),
    "tiled4k": (0, (512, 512, 512), 16, (3840, 2160), "BASELINE config 3: 16x16 physical tiling (multi-GB RLE) at 3840x2160"),
    "shortrun4k": (1, (2048, 1024, 2048), 1, (3840, 2160), "BASELINE config 4 (reduced footprint): worst-case short-run band at 3840x2160"),
    "shortrun16k": (2, (16384, 1024, 16384), 1, (3840, 2160), "BASELINE config 4 at full size: 16384 x 1024 x 16384 heightfield + cave floors + worst-case short-run band (rlerc_synth_rle, ~5.5 GB of RLE incl. mips) at 3840x2160"),
    "view8k": (0, (512, 512, 512), 16, (7680, 4320), "BASELINE config 5: 7680x4320 views of the tiled scene (N > 1, --mp frames: one camera per GPU, NVLink gather to rank 0)"),
    "small": (0, (256, 256, 256), 1, (1024, 768), "quick functional run"),
}


def path_pose(R, i, K, sy, imrodh=False):
    """Fly-through of BASELINE config 2: orbit of radius 4000 around (10000, 10000),
    pitch 0.35 +- 0.3, one full turn in 1000 frames.  The camera's voxel height is -pos.y
    (Cuda_Render.h:453,534: a run at voxel y sits at y + viewpos.y below the eye), so for the real
    Imrodh.rle4 the survey's -818 +- 300 is kept, while the synthetic terrain (surface between sy/4
    and 3*sy/4, volume top at 0) is flown over at 0.15*sy +- 0.08*sy, above its highest peaks."""
    t = (i * 1000) // max(K, 1)
    pos, rot = R.flythrough_pose(t, 1000)
    if imrodh:
        return pos, rot
    a = 2.0 * math.pi * t / 1000.0
    return (pos[0], -(0.15 + 0.08 * math.sin(2 * a)) * sy, pos[2]), rot


def build_scene(R, workload, log):
    kind, (sx, sy, sz), tiling, _, _ = WORKLOADS[workload]
    path = os.environ.get("RLERC_IMRODH", "")
    t0 = time.time()
    if workload.startswith("imrodh") and path and os.path.exists(path):
        scene, name = R.RLE4.load(path), "Imrodh.rle4"
        sy = scene.level(0)[1]
    else:
        if kind == 2:
            scene = R.RLE4.synth_rle(sx, sy, sz, seed=42, band_every=48)
        else:
            scene = R.RLE4.synth(kind, sx, sy, sz, seed=1 if kind == 0 else 42)
        name = "synth_%s_%dx%dx%d" % ({0: "imrodh", 1: "shortrun", 2: "rle_shortrun"}[kind], sx, sy, sz)
        if tiling > 1:
            scene = scene.tile(tiling, tiling)
            name += "_tiled%dx%d" % (tiling, tiling)
    log("scene %s: %.1f s, %d levels, %.1f MB" % (name, time.time() - t0, scene.nummaps, scene.nbytes() / 1e6))
    return scene, name, sy


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6.0)          # a query that overlapped a very short timed region still counts
        sm = sorted(int(r[1]) for r in self.rows if r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if r[2].isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


KERNEL_NAMES = {0: "k_traverse_f", 65: "k_traverse_f", 64: "k_traverse_w", 66: "k_traverse_c<3>", 67: "k_traverse_c<7>"}


def measured_ncu(workload, kernel):
    """What else the committed ncu capture says about that launch (IPC, hit rates, sectors per request)."""
    try:
        e = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(workload)
        if e and e.get("kernel") == kernel:
            return e.get("ncu")
    except Exception:
        pass
    return None


def measured_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the traversal kernel, from the committed
    `ncu --set full` capture (profiles/traffic.json; null when there is none for this workload / kernel)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        e = t.get(workload)
        if e and e.get("kernel") == kernel:
            return e["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


def traversal_bytes(c):
    """Algorithmic bytes of one traversal launch from its work counters."""
    C_, E, C1, P, K = c["cols_fetched"], c["run_iters"], c["cols_nonempty"], c["pixels"], c["cleared"]
    return 8 * C_ + 2 * (E - C1) + 6 * P + 4 * K


def cpu_frames(R, rb, scene, cfg, poses, threads):
    """Reference render_line (oracle/_ref when present, else the port) + the oracle's unwarp. Seconds per frame list."""
    levels = [scene.level(m) for m in range(scene.nummaps)]
    use_ref = rb.have_ref()
    out = []
    for pos, rot in poses:
        rm = R.RayMap(cfg).get_ray_map(pos, rot)
        orm = rb.RayMapGPU()
        C.memmove(C.byref(orm), C.byref(rm), 896)
        rb.attach_host_scene(orm, levels)
        t0 = time.perf_counter()
        if use_ref:
            warp, _ = rb.ref_render_frame(orm, cfg.render_size, mip_distance=cfg.mip_distance, z_far=cfg.z_far,
                                          rays=cfg.rays_casted, threads=threads)
        else:
            warp, _, _ = rb.orc_render(orm, cfg.render_size, cfg.rays_casted, cfg.mip_distance, cfg.z_far, threads=threads)
        t1 = time.perf_counter()
        rb.orc_unwarp(orm, cfg.width, cfg.height, cfg.render_size, cfg.rays_casted, cfg.rays_casted_res, warp)
        t2 = time.perf_counter()
        out.append((t1 - t0, t2 - t1))
    return out, ("reference" if use_ref else "port")


def main():
    # The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints its version banner to
    # fd 1 on communicator creation), so everything but the final line goes to stderr at the file-descriptor level.
    json_out = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="imrodh1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="rlerc", choices=["rlerc", "reference"])
    ap.add_argument("--lanes", type=int, default=0,
                    help="traversal kernel variant (rlerc_set_lanes_per_ray): 0 = k_traverse_f (production), 64 = k_traverse_w, "
                         "66/67 = k_traverse_c, 1..32 = k_traverse<lanes>")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--slice-block", type=int, default=32)
    ap.add_argument("--mp", default="frames", choices=["frames", "slices"],
                    help="N > 1: deal whole frames to the GPUs (throughput, default) or split every frame into ray-plane slices (latency, north-star mode)")
    args = ap.parse_args()
    K, W = max(1, args.steps), max(3, args.warmup)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    log = (lambda m: print("[bench r%d] %s" % (rank, m), file=sys.stderr, flush=True))

    import __graft_entry__ as g
    R = g.build(quiet=True)
    kind, _, _, (WW, HH), desc = WORKLOADS[args.workload]
    cfg = R.FrameConfig.default(WW, HH)
    # torchrun exports OMP_NUM_THREADS=1: give every rank its share of the host cores for scene construction
    R.lib().rlerc_set_host_threads(max(1, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world)))))

    # ------------------------------------------------------------------ reference arm (host CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        from oracle import refbind as rb
        scene, scene_name, sy = build_scene(R, args.workload, log)
        threads = os.cpu_count() or 1
        poses = [path_pose(R, i, K, sy, scene_name == "Imrodh.rle4") for i in range(K)]
        cpu_frames(R, rb, scene, cfg, poses[:min(W, 3)], threads)
        t0 = time.perf_counter()
        times, kindname = cpu_frames(R, rb, scene, cfg, poses, threads)
        wall = time.perf_counter() - t0
        fps = K / wall
        val = WW * HH * fps / 1e6
        line = {"impl": "reference", "metric": "Mrays/s", "value": val, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": K, "warmup": W,
                "ms_per_step": 1e3 * wall / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic" if scene_name != "Imrodh.rle4" else "Imrodh.rle4",
                "frames_per_s": fps,
                "config": {"workload": args.workload, "scene": scene_name, "window": [WW, HH], "render_size": cfg.render_size,
                           "rays_casted": cfg.rays_casted, "description": desc},
                "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": threads, "kind": kindname,
                                 "sample": "%d fly-through frames: reference render_line (OpenMP over ray planes, %.1f ms/frame) + oracle unwarp (%.1f ms/frame)"
                                           % (K, 1e3 * sum(t[0] for t in times) / K, 1e3 * sum(t[1] for t in times) / K)},
                "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), file=json_out, flush=True)
        return 0

    # ------------------------------------------------------------------ our arm (B200)
    import numpy as np
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    MG = importlib.import_module("rle-based-voxel-raycasting_b200.multigpu")

    scene, scene_name, sy = build_scene(R, args.workload, log)
    r = R.Renderer(local)
    t0 = time.time()
    r.all_to_gpu(scene)
    log("replica uploaded in %.1f s" % (time.time() - t0))
    r.set_lanes_per_ray(args.lanes)
    frame = MG.SlicedFrame(r, cfg, torch, rank=rank, world=world, dist=dist if world > 1 else None, block=args.slice_block)
    poses = [path_pose(R, i, K, sy, scene_name == "Imrodh.rle4") for i in range(K)]
    raymaps = [R.RayMap(cfg).get_ray_map(p, q) for p, q in poses]
    flush = None if args.no_flush else torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    farm = MG.FrameFarm(r, cfg, torch, rank, world, dist) if (world > 1 and args.mp == "frames") else None
    rounds = (K + world - 1) // world

    # ---- warm-up (also initialises NCCL's channels: the first collectives take 100+ ms)
    for i in range(W):
        frame.render(raymaps[i % K])
    if farm is not None:
        for rd in range(max(W, 4)):
            farm.render_round(rd, raymaps[(rd * world + rank) % K])
        farm.finish()
    barrier()

    # ---- timed region: K steps (frames), device time per step from CUDA events on the launching stream
    sampler = ClockSampler(local)
    sampler.start()
    r.set_timing(True)
    trav_ms, unwarp_ms = 0.0, 0.0
    ev = []
    barrier()
    wall0 = time.perf_counter()
    if farm is None:
        # 1 GPU, or every frame split into ray-plane slices over all GPUs
        for i in range(K):
            if flush is not None:
                flush.fill_(i & 255)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            frame.render(raymaps[i])
            e1.record()
            e1.synchronize()
            ev.append((e0, e1))
            a, b_ = r.last_kernel_ms()
            trav_ms += a
            unwarp_ms += b_
    else:
        # whole frames dealt round-robin; the gather of round rd overlaps the rendering of round rd+1
        for rd in range(rounds):
            i = rd * world + rank
            if flush is not None:
                flush.fill_(rd & 255)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            farm.render_round(rd, raymaps[i] if i < K else None)
            e1.record()
            e1.synchronize()
            ev.append((e0, e1))
            if i < K:
                a, b_ = r.last_kernel_ms()
                trav_ms += a
                unwarp_ms += b_
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        farm.finish()
        e1.record()
        ev.append((e0, e1))
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.summary()
    r.set_timing(False)
    per_step = sorted(s.elapsed_time(e) for s, e in ev[:K if farm is None else rounds])
    dev_ms = sum(s.elapsed_time(e) for s, e in ev)
    # frame-time distribution on this rank (mean / p50 / p99); frames mode: per round
    frame_ms = {"mean": sum(per_step) / len(per_step), "p50": per_step[len(per_step) // 2],
                "p99": per_step[min(len(per_step) - 1, int(0.99 * len(per_step)))], "max": per_step[-1],
                "per": "frame" if farm is None else "round of %d frames" % world} if per_step else None
    t = torch.tensor([dev_ms, trav_ms, unwarp_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, trav_ms_max, unwarp_ms_max = (float(v) for v in t.cpu())
    fps = K / (dev_ms / 1e3)
    value = WW * HH * fps / 1e6
    launches_per_rank = (K if farm is None else len(range(rank, K, world)))
    kernels_timed = launches_per_rank               # traversal launches behind trav_ms on the slowest rank

    # ---- warm-L2 variant (no flush), for information
    warm_ms = None
    if farm is None:
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(K):
            frame.render(raymaps[i])
        s1.record()
        barrier()
        warm_ms = s0.elapsed_time(s1)

    # ---- roofline of the traversal kernel: algorithmic bytes from the instrumented kernel (N=1 semantics)
    roof = None
    cpu = None
    e2e = None
    if rank == 0:
        sample_idx = [int(j * K / 8) for j in range(8)] if K >= 8 else list(range(K))
        ids = torch.empty((cfg.rays_casted, cfg.render_size, 2), dtype=torch.int32, device="cuda")
        tot_bytes, tot_rays = 0, 0
        agg = {}
        for j in sample_idx:
            r.render_ids(raymaps[j], cfg, ids.data_ptr())
            r.sync()
            c = dict(zip(["elems_total", "elems_processed", "voxels_processed", "elems_rendered", "pixels", "cols_fetched",
                          "run_iters", "cols_nonempty", "cleared", "dda_steps"], r.counters()))
            tot_bytes += traversal_bytes(c)
            tot_rays += raymaps[j].map_line_count
            for k_, v in c.items():
                agg[k_] = agg.get(k_, 0) + v
        del ids
        bytes_per_frame = tot_bytes / len(sample_idx)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        # per-rank launches traverse 1/world of the ray planes
        sliced = world > 1 and farm is None
        per_launch_bytes = bytes_per_frame / (world if sliced else 1)
        achieved = per_launch_bytes / (trav_ms_max / max(kernels_timed, 1) / 1e3) / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": measured_traffic(args.workload, KERNEL_NAMES.get(args.lanes, "k_traverse<%d>" % args.lanes)),
                "kernel": KERNEL_NAMES.get(args.lanes, "k_traverse<%d>" % args.lanes),
                "ncu": measured_ncu(args.workload, KERNEL_NAMES.get(args.lanes, "k_traverse<%d>" % args.lanes)),
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                "algorithmic_bytes_per_launch": per_launch_bytes,
                "traverse_ms_per_launch": trav_ms_max / max(kernels_timed, 1), "unwarp_ms_per_launch": unwarp_ms_max / max(kernels_timed, 1),
                "unwarp_achieved_gbs": 8.0 * WW * HH / (max(unwarp_ms_max, 1e-9) / max(kernels_timed, 1) / 1e3) / 1e9,
                "counters_per_frame": {k_: v / len(sample_idx) for k_, v in agg.items()},
                "plane_rays_per_frame": tot_rays / len(sample_idx)}

    # ---- same-GPU comparator: the reference's own scheme (one THREAD per ray plane, k_traverse<1>) on a few frames
    if rank == 0 and world == 1 and roof is not None and args.lanes == 0:
        r.set_lanes_per_ray(1)
        r.set_timing(True)
        ms = []
        for j in ([int(j * K / 3) for j in range(3)] if K >= 3 else list(range(K))):
            r.render(raymaps[j], cfg)
            r.sync()
            ms.append(r.last_kernel_ms()[0])
        r.set_timing(False)
        r.set_lanes_per_ray(args.lanes)
        roof["reference_scheme_on_this_gpu"] = {"kernel": "k_traverse<1>: one thread per ray plane, as cudaRender (R/src/Cuda_Main.cu:150-181)",
                                                "traverse_ms_per_launch": sum(ms) / len(ms), "frames": len(ms)}

    # ---- e2e through the C ABI with host buffers (pinned), D2H inside the timed region
    barrier()
    e2e_wall = None
    if world == 1:
        r2 = R.Renderer(local)
        r2.all_to_gpu(scene)
        r2.set_lanes_per_ray(args.lanes)
        DEPTH = int(os.environ.get("RLERC_E2E_DEPTH", "4"))          # frames in flight (rlerc_frame_submit pipelines up to RLERC_FRAME_SLOTS)
        pins = [R.PinnedBuffer((HH, WW, 4)) for _ in range(DEPTH)]
        for i in range(max(W, 8)):      # every frame slot (stream, warped buffer, RGBA buffer) exists before the timed region
            r2.frame_wait(r2.frame_submit(poses[i % K][0], poses[i % K][1], cfg, pins[i % DEPTH].array))
        r2.sync()
        t0 = time.perf_counter()
        tickets = []
        for i in range(K):
            if len(tickets) >= DEPTH:
                r2.frame_wait(tickets.pop(0))
            tickets.append(r2.frame_submit(poses[i][0], poses[i][1], cfg, pins[i % DEPTH].array))
        for tk in tickets:
            r2.frame_wait(tk)
        r2.sync()
        e2e_wall = time.perf_counter() - t0
        checksum = int(pins[(K - 1) % DEPTH].array[::16, ::16].astype(np.uint32).sum())
        for p in pins:
            p.free()
        r2.close()
    elif farm is None:
        host = torch.empty((HH, WW, 4), dtype=torch.uint8).pin_memory() if rank == 0 else None
        barrier()
        t0 = time.perf_counter()
        for i in range(K):
            rm = R.RayMap(cfg).get_ray_map(poses[i][0], poses[i][1])
            img = frame.render(rm)
            if rank == 0:
                host.copy_(img, non_blocking=True)
        barrier()
        e2e_wall = time.perf_counter() - t0
        checksum = int(host[::16, ::16].to(torch.int64).sum()) if rank == 0 else 0
    else:
        hosts = [torch.empty((HH, WW, 4), dtype=torch.uint8).pin_memory() for _ in range(world)] if rank == 0 else None
        barrier()
        t0 = time.perf_counter()
        for rd in range(rounds + 1):
            if rd < rounds:
                i = rd * world + rank
                rm = R.RayMap(cfg).get_ray_map(poses[i][0], poses[i][1]) if i < K else None
                farm.render_round(rd, rm)
            if rd > 0 and rank == 0:
                done = farm.wait_round(rd - 1)          # frames of the previous round are on rank 0: D2H them
                for j in range(world):
                    if (rd - 1) * world + j < K:
                        hosts[j].copy_(done[j], non_blocking=True)
        farm.finish()
        barrier()
        e2e_wall = time.perf_counter() - t0
        checksum = int(hosts[(K - 1) % world][::16, ::16].to(torch.int64).sum()) if rank == 0 else 0
    tt = torch.tensor([e2e_wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_wall = float(tt.cpu()[0])
    e2e_fps = K / e2e_wall
    e2e = {"value": WW * HH * e2e_fps / 1e6, "unit": "Mrays/s", "frames_per_s": e2e_fps,
           "h2d_bytes_per_step": 1024 * world, "d2h_bytes_per_step": WW * HH * 4,
           "how": "pinned host RGBA out, camera pose in; %s" % ("rlerc_frame_submit/wait, %s frames in flight, one stream per frame slot" % os.environ.get("RLERC_E2E_DEPTH", "4") if world == 1
                                                                 else ("per-rank slices + NCCL reduce + D2H on rank 0" if farm is None
                                                                       else "whole frames per rank + NCCL gather + D2H of every frame on rank 0")),
           "checksum": checksum}

    # ---- CPU baseline on the box's host cores (rank 0, N=1 only), bounded sample
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import refbind as rb
        threads = os.cpu_count() or 1
        idx = [int(j * K / 6) for j in range(6)] if K >= 6 else list(range(K))
        sample = [poses[j] for j in idx]
        cpu_frames(R, rb, scene, cfg, sample[:1], threads)
        tms, kindname = cpu_frames(R, rb, scene, cfg, sample, threads)
        one, _ = cpu_frames(R, rb, scene, cfg, sample[:2], 1)
        per = sum(a + b for a, b in tms) / len(tms)
        per1 = sum(a + b for a, b in one) / len(one)
        pern = sum(a + b for a, b in tms[:2]) / 2
        cpu = {"value": WW * HH / per / 1e6, "unit": "Mrays/s", "cores": threads, "kind": kindname,
               "frames_per_s": 1.0 / per,
               "sample": "%d of the %d fly-through frames: reference render_line compiled for the host (OpenMP, %d threads, %.1f ms/frame) "
                         "+ oracle unwarp (%.1f ms/frame); 1 thread: %.1f ms/frame -> 1->%d thread speed-up %.1fx"
                         % (len(tms), K, threads, 1e3 * sum(a for a, _ in tms) / len(tms), 1e3 * sum(b for _, b in tms) / len(tms),
                            1e3 * per1, threads, per1 / pern)}

    if rank == 0:
        line = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic" if scene_name != "Imrodh.rle4" else "Imrodh.rle4",
                "frames_per_s": fps,
                "config": {"workload": args.workload, "scene": scene_name, "scene_mb": scene.nbytes() / 1e6, "window": [WW, HH],
                           "steps_are": "frames of the fly-through; the K frames of the job are the same for every N",
                           "render_size": cfg.render_size, "rays_casted": cfg.rays_casted, "z_far": cfg.z_far,
                           "description": desc, "l2": "flushed between timed steps (512 MiB write)" if flush is not None else "not flushed",
                           "warm_l2_value": (WW * HH * (K / (warm_ms / 1e3)) / 1e6) if warm_ms else None, "lanes": args.lanes,
                           "parallelism": ("1 GPU" if world == 1 else
                                           ("whole frames dealt round-robin to %d GPUs (full replica each), NCCL gather of finished frames to rank 0, overlapped" % world
                                            if farm is not None else
                                            "every frame split into ray-plane slices x%d, interleaved blocks of %d, NCCL reduce to rank 0" % (world, args.slice_block))),
                           "wall_s_timed_region": wall},
                "clocks": clocks, "e2e": e2e, "gpu_launches": 2 * K, "frame_ms": frame_ms, "roofline": roof}
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), file=json_out, flush=True)
    barrier()
    if dist.is_initialized():
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
