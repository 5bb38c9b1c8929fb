#!/usr/bin/env python
"""Benchmark of the RLE voxel raycaster frame loop (BASELINE.json metric: Mrays/s & frames/s).

    python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference]

A step is ONE FRAME of the hot path (frame setup -> k_dda_states + traversal kernel -> unwarp kernel [-> NCCL
compositing for N > 1]) on the scripted fly-through of BASELINE config 2; step i of K renders path frame
i*1000/K, so any K covers the whole orbit.  One process per GPU (torchrun sets RANK/...).

value      pixel-rays/s = W*H*frames/s of the whole K-frame job, scene + buffers resident in HBM, `--inflight` (4)
           frames in flight per GPU (one stream and one set of frame buffers per slot), device-timed from a CUDA event
           in front of the first frame to one behind the last, max over ranks.  The job is repeated R times so that the
           timed region lasts >= 0.6 s (ms_per_step = total / (R*K); per-repeat times in `repeats`).  For N > 1 every
           frame is split into interleaved ray-plane slices over all GPUs (north-star split); every GPU unwarps its
           band of rows, pulling texels from the owners' warped buffers over NVLink, and pushes the pixels into rank
           0's frame (csrc/group.cu; `--compose reduce` = one NCCL reduce of whole images instead); the same
           method at every N.
latency    the same frames one at a time, L2 flushed (512 MiB write) before every frame, per-frame CUDA events:
           the round-1 `value` method, kept for comparison (mean / p50 / p99 / max).
e2e        the same metric through the C ABI with HOST buffers: camera pose in, RGBA frame out in pinned host memory,
           copies inside the timed region.  N = 1: rlerc_frame_submit / rlerc_frame_wait; N > 1: rlerc_group_submit,
           every rank copies the rows it produced into ONE host frame buffer shared by all ranks over its own PCIe
           link (no rank-0 funnel).
roofline   traversal kernel: algorithmic bytes per frame (8C + 2(E-C1) + 6P + 4K, DESIGN.md §5, counted by the
           instrumented kernel on 8 path frames) / its CUDA-event duration in the latency pass (kernel timed alone:
           burst peak), vs the measured HBM peak; L2 and DRAM traffic of the same launch from the committed ncu capture.
parity     frames of THIS workload compared with the reference's render_line compiled for the host (oracle/_ref),
           outside the timed regions.
cpu_baseline / --impl reference: the reference's own render_line compiled for the host (oracle/_ref, OpenMP over ray
           planes) + the oracle's unwarp, on a bounded sample of the same frames.
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
# Frames in flight live on one stream each; with the default of 8 hardware queues several streams share a queue and a
# kernel that waits (the flag barriers of the multi-GPU path) holds up the unrelated kernels queued behind it.  Has to be in
# the environment before the CUDA context exists (the library sets it too, rlerc_create).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

WORKLOADS = {
    # name: (scene kind, base size (sx, sy, sz), tiling, window, description)
    "imrodh1080p": (0, (1024, 1024, 1024), 1, (1920, 1080), "BASELINE config 2: Imrodh.rle4 (or synth_imrodh 1024^3 seed 1) at 1920x1080, scripted fly-through"),
    "imrodh768": (0, (1024, 1024, 1024), 1, (1024, 768), "BASELINE config 1 window: 1024x768"),
    "tiled4k": (0, (512, 512, 512), 16, (3840, 2160), "BASELINE config 3: 16x16 physical tiling (multi-GB RLE) at 3840x2160"),
    "shortrun4k": (1, (2048, 1024, 2048), 1, (3840, 2160), "BASELINE config 4 (reduced footprint): worst-case short-run band at 3840x2160"),
    "shortrun16k": (2, (16384, 1024, 16384), 1, (3840, 2160), "BASELINE config 4 at full size: 16384 x 1024 x 16384 heightfield + cave floors + worst-case short-run band (rlerc_synth_rle, ~5.5 GB of RLE incl. mips) at 3840x2160"),
    "view8k": (0, (512, 512, 512), 16, (7680, 4320), "BASELINE config 5: 7680x4320 views of the tiled scene"),
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


def make_config(workload, scene_name, cfg):
    """What identifies the workload: the SAME dict in both arms (the driver compares them)."""
    _, _, _, (WW, HH), desc = WORKLOADS[workload]
    return {"workload": workload, "scene": scene_name, "window": [WW, HH], "render_size": cfg.render_size,
            "rays_casted": cfg.rays_casted, "z_far": cfg.z_far, "description": desc,
            "steps_are": "frames of the fly-through; step i of K is path frame i*1000/K, the same K frames for every N and both arms"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md); rank 0 only."""
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
            time.sleep(0.25)

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


def committed_capture(workload, kernel):
    """The committed `ncu --set full` capture of this workload's traversal launch (profiles/traffic.json), if the
    kernel it profiled is the kernel this run launched."""
    try:
        e = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(workload)
        if e and e.get("kernel") == kernel:
            return e
    except Exception:
        pass
    return None


def traversal_bytes(c):
    """Algorithmic bytes of one traversal launch from its work counters."""
    C_, E, C1, P, K = c["cols_fetched"], c["run_iters"], c["cols_nonempty"], c["pixels"], c["cleared"]
    return 8 * C_ + 2 * (E - C1) + 6 * P + 4 * K


COUNTER_NAMES = ["elems_total", "elems_processed", "voxels_processed", "elems_rendered", "pixels", "cols_fetched",
                 "run_iters", "cols_nonempty", "cleared", "dda_steps"]


def oracle_raymap(rb, rm, levels):
    orm = rb.RayMapGPU()
    C.memmove(C.byref(orm), C.byref(rm), 896)
    rb.attach_host_scene(orm, levels)
    return orm


def cpu_frames(R, rb, scene, cfg, poses, threads, keep=False):
    """Reference render_line (oracle/_ref when present, else the port) + the oracle's unwarp.
    Returns [(traversal s, unwarp s[, warp, rgba])] per pose and the kind of baseline."""
    levels = [scene.level(m) for m in range(scene.nummaps)]
    use_ref = rb.have_ref()
    out = []
    for pos, rot in poses:
        rm = R.RayMap(cfg).get_ray_map(pos, rot)
        orm = oracle_raymap(rb, rm, levels)
        t0 = time.perf_counter()
        if use_ref:
            warp, _ = rb.ref_render_frame(orm, cfg.render_size, mip_distance=cfg.mip_distance, z_far=cfg.z_far,
                                          rays=cfg.rays_casted, threads=threads)
        else:
            warp, _, _ = rb.orc_render(orm, cfg.render_size, cfg.rays_casted, cfg.mip_distance, cfg.z_far, threads=threads)
        t1 = time.perf_counter()
        rgba = rb.orc_unwarp(orm, cfg.width, cfg.height, cfg.render_size, cfg.rays_casted, cfg.rays_casted_res, warp)
        t2 = time.perf_counter()
        out.append((t1 - t0, t2 - t1, warp, rgba) if keep else (t1 - t0, t2 - t1))
    return out, ("reference" if use_ref else "port")


def pipelined_job(torch, pipe, raymaps, K, repeats, barrier):
    """R x K frames through `pipe`, device-timed as ONE region.  Returns (total ms, [ms per repeat])."""
    main = torch.cuda.current_stream()
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record(main)
    pipe.start_after(e0)
    marks = []
    for rep in range(repeats):
        for i in range(K):
            k = pipe.submit(rep * K + i, raymaps[i])
        m = torch.cuda.Event(enable_timing=True)
        m.record(pipe.streams[k])
        marks.append(m)
    pipe.drain(main)
    e1 = torch.cuda.Event(enable_timing=True)
    e1.record(main)
    e1.synchronize()
    barrier()
    ends = [e0.elapsed_time(m) for m in marks]
    per = [b - a for a, b in zip([0.0] + ends[:-1], ends)]
    return e0.elapsed_time(e1), per


def sequential_job(torch, pipe, raymaps, frames, flush, timing_r=None):
    """Frames one at a time, L2 flushed before each, CUDA events around each frame (the round-1 `value` method).
    Returns ([ms per frame], traversal ms sum, unwarp ms sum)."""
    main = torch.cuda.current_stream()
    out, trav, unw = [], 0.0, 0.0
    for i in frames:
        if flush is not None:
            flush.fill_(i & 255)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(main)
        pipe.start_after(a)
        k = pipe.submit(i, raymaps[i])
        pipe.finish_slot(k)
        b.record(pipe.streams[k])
        b.synchronize()
        out.append(a.elapsed_time(b))
        if timing_r is not None:
            t, u = timing_r.last_kernel_ms()
            trav += t
            unw += u
    return out, trav, unw


def dist_stats(xs):
    xs = sorted(xs)
    return {"mean": sum(xs) / len(xs), "p50": xs[len(xs) // 2], "p99": xs[min(len(xs) - 1, int(0.99 * len(xs)))], "max": xs[-1]}


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
                    help="traversal kernel (rlerc_set_lanes_per_ray): 0 = automatic (k_traverse_f; k_traverse_q for small launches rendered one at a time), "
                         "65 = k_traverse_f, 69 = k_traverse_q, 68 = k_traverse_p, 1..32 = k_traverse<lanes>")
    ap.add_argument("--inflight", type=int, default=0,
                    help="frames in flight per GPU in the throughput measurements (0 = 8 on one GPU, 16 on several: a slice of a frame "
                         "is bound by its longest ray planes, not by the GPU, so the GPUs are kept busy by more frames in flight)")
    ap.add_argument("--min-seconds", type=float, default=0.6, help="the K-frame job is repeated until the timed region lasts this long")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-north-star", action="store_true", help="N > 1: skip the tiled4k slice-scaling block")
    ap.add_argument("--slice-block", type=int, default=32)
    ap.add_argument("--compose", default="pull", choices=["pull", "reduce"],
                    help="N > 1, slices: pull = unwarp bands with texels pulled over NVLink peer memory (csrc/group.cu, default); "
                         "reduce = every GPU unwarps its own pixels of the whole window, one NCCL reduce(sum) to rank 0")
    ap.add_argument("--mp", default="slices", choices=["slices", "frames", "views"],
                    help="N > 1: split every frame into ray-plane slices (north-star split, default), deal whole frames to the GPUs, or "
                         "views = BASELINE config 5: every GPU renders its own camera (the path shifted by rank * K / N frames) and the "
                         "finished frames are gathered on rank 0 over NVLink")
    args = ap.parse_args()
    K, W = max(1, args.steps), max(3, args.warmup)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    F = args.inflight if args.inflight > 0 else (8 if world == 1 else 16)
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
        R.lib().rlerc_set_host_threads(os.cpu_count() or 1)
        scene, scene_name, sy = build_scene(R, args.workload, log)
        threads = os.cpu_count() or 1
        poses = [path_pose(R, i, K, sy, scene_name == "Imrodh.rle4") for i in range(K)]
        cpu_frames(R, rb, scene, cfg, poses[:min(W, 3)], threads)
        t0 = time.perf_counter()
        times, kindname = cpu_frames(R, rb, scene, cfg, poses, threads)
        wall = time.perf_counter() - t0
        fps = K / wall
        val = WW * HH * fps / 1e6
        trav = sum(t[0] for t in times) / K
        line = {"impl": "reference", "metric": "Mrays/s", "value": val, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": K, "warmup": W,
                "ms_per_step": 1e3 * wall / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic" if scene_name != "Imrodh.rle4" else "Imrodh.rle4",
                "frames_per_s": fps, "config": make_config(args.workload, scene_name, cfg),
                "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": threads, "kind": kindname,
                                 "traversal_only_ms_per_frame": 1e3 * trav, "traversal_only_value": WW * HH / trav / 1e6,
                                 "sample": "%d fly-through frames: reference render_line (OpenMP over ray planes, %.1f ms/frame) + oracle unwarp (%.1f ms/frame)"
                                           % (K, 1e3 * trav, 1e3 * sum(t[1] for t in times) / K)},
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
    D = dist if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.cpu()[0])

    def throughput(pipe, raymaps, K_, min_s):
        """Warm up, calibrate the repeat count, run the timed job.  Returns dict(value-free numbers)."""
        for i in range(max(W, 2 * pipe.depth)):
            pipe.submit(i, raymaps[i % K_])
        pipe.drain()
        barrier()
        t1, _ = pipelined_job(torch, pipe, raymaps, K_, 1, barrier)
        t1 = allmax(t1)
        reps = int(max(1, min(200, math.ceil(min_s * 1e3 / max(t1, 1e-3)))))
        tot, per = pipelined_job(torch, pipe, raymaps, K_, reps, barrier)
        tot = allmax(tot)
        med = sorted(per)[len(per) // 2]
        return {"ms_total": tot, "repeats": reps, "ms_per_frame": tot / (reps * K_), "fps": reps * K_ / (tot / 1e3),
                "per_repeat_ms": {"median": med, "max": max(per), "max_over_median": max(per) / med if med > 0 else None}}

    scene, scene_name, sy = build_scene(R, args.workload, log)
    imrodh = scene_name == "Imrodh.rle4"
    scene_megabytes = scene.nbytes() / 1e6
    t0 = time.time()
    slices = world > 1 and args.mp == "slices"
    views = world > 1 and args.mp == "views"
    F_default = F

    def make_pipe(scene_, cfg_, rank_, world_, dist_, dst=0, host=None, share_from=None, depth=None):
        F = depth or F_default
        if views and world_ > 1:
            return MG.GroupPipe(R, torch, local, scene_, cfg_, depth=F, rank=rank_, world=world_, dist=dist_, lanes=args.lanes,
                                dst=dst, host=host, share_from=share_from, views=True)
        if args.compose == "reduce" and world_ > 1:
            return MG.FramePipe(R, torch, local, scene_, cfg_, depth=F, rank=rank_, world=world_, dist=dist_, block=args.slice_block,
                                lanes=args.lanes, compose="reduce" if dst >= 0 else "bands", host=host, share_from=share_from)
        return MG.GroupPipe(R, torch, local, scene_, cfg_, depth=F, rank=rank_, world=world_, dist=dist_, block=args.slice_block,
                            lanes=args.lanes, dst=dst, host=host, share_from=share_from)

    grouped = slices or views           # one group over all ranks (flag barriers per frame); else every rank on its own
    pipe = make_pipe(scene, cfg, rank if grouped else 0, world if grouped else 1, D if grouped else None)
    log("replica uploaded, %d frame slots: %.1f s" % (F, time.time() - t0))
    r = pipe.r[0]
    poses = [path_pose(R, i, K, sy, imrodh) for i in range(K)]
    raymaps = []
    for p, q in poses:
        rm = R.RayMapGPU()
        C.memmove(C.byref(rm), C.byref(R.RayMap(cfg).get_ray_map(p, q)), 896)
        raymaps.append(rm)

    # ---- the timed region: R x K frames, F in flight per GPU
    if slices or world == 1:
        my_maps, myK = raymaps, K
    elif views:                 # --mp views: every rank flies the same path, shifted by rank * K / N frames: N different cameras at any time
        sh = (rank * K) // world
        my_maps, myK = raymaps[sh:] + raymaps[:sh], K
    else:                       # --mp frames: whole frames dealt round-robin, each stays on the GPU that rendered it
        my_maps = raymaps[rank::world]
        myK = len(my_maps)
    sampler = ClockSampler(local) if rank == 0 else None
    # warm-up outside the sampled interval (also initialises NCCL's channels: the first collectives take 100+ ms)
    for i in range(max(W, 2 * F)):
        pipe.submit(i, my_maps[i % myK])
    pipe.drain()
    barrier()
    if sampler:
        sampler.start()
    wall0 = time.perf_counter()
    th = throughput(pipe, my_maps, myK, args.min_seconds)
    wall = time.perf_counter() - wall0
    clocks = sampler.summary() if sampler else None
    frames_done = th["repeats"] * K * (world if views else 1)
    if not (slices or views or world == 1):
        # every rank did repeats * len(my_maps) frames in ms_total; the job is all ranks' frames
        t = torch.tensor([th["repeats"] * myK], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        frames_done = int(t.cpu()[0])
    fps = frames_done / (th["ms_total"] / 1e3)
    value = WW * HH * fps / 1e6
    # k_dda_states + traversal + k_unwarp per frame on this rank (+ two k_group_barrier with the peer-memory compositor)
    launches = (5 if (slices and args.compose == "pull") else (4 if views else 3)) * th["repeats"] * myK

    # ---- latency: one frame at a time, L2 flushed before each (the round-1 `value` method)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    r.set_timing(True)
    KL = min(K, 200)
    seq_frames = [int(j * K / KL) for j in range(KL)]
    lat = trav_ms = unw_ms = None
    if slices or views or world == 1:
        barrier()
        per, trav_ms, unw_ms = sequential_job(torch, pipe, my_maps, seq_frames, flush, r)
        barrier()
        tot = allmax(sum(per))
        lat = dict(dist_stats(per), frames=KL, value=(world if views else 1) * WW * HH * KL / (tot / 1e3) / 1e6, frames_per_s=(world if views else 1) * KL / (tot / 1e3),
                   how="one frame at a time, 512 MiB L2 flush before each, CUDA events around each frame (%s), this rank; value from the max over ranks of the summed frame times"
                       % (("traversal slice + barrier + unwarp band with NVLink pull + barrier" if args.compose == "pull" else "traversal slice + unwarp + NCCL reduce") if slices else
                          ("k_dda_states + traversal + unwarp of this rank's own view + copy to rank 0 over NVLink + barrier" if views else "k_dda_states + traversal + unwarp")))
        trav_ms, unw_ms = allmax(trav_ms), allmax(unw_ms)
    r.set_timing(False)
    kernel_used = r.last_kernel
    del flush

    # ---- multi-GPU frame == single-GPU frame (rank 0 renders the whole frame itself and compares)
    composite_identical = None
    if slices:
        checks = [0, K // 2]
        ok = True
        for i in checks:
            k_ = pipe.submit(0, raymaps[i])
            pipe.finish_slot(k_)
            pipe.drain()
            barrier()
            if rank == 0:
                got, _, _ = pipe.image_numpy(k_)
                solo = torch.zeros((HH, WW, 4), dtype=torch.uint8, device="cuda")
                r.frame_device(raymaps[i], cfg, 1, 1, 0, solo.data_ptr())
                torch.cuda.synchronize()
                ok = ok and bool(np.array_equal(got, solo.cpu().numpy()))
            barrier()
        composite_identical = ok if rank == 0 else None
    if views:
        # every view as it arrived on rank 0 against rank 0's own render of that rank's camera
        ok = True
        k_ = pipe.submit(0, my_maps[3 % myK])
        pipe.drain()
        barrier()
        if rank == 0:
            for rr in range(world):
                sh = (rr * K) // world
                want_map = (raymaps[sh:] + raymaps[:sh])[3 % K]
                solo = torch.zeros((HH, WW, 4), dtype=torch.uint8, device="cuda")
                r.frame_device(want_map, cfg, 1, 1, 0, solo.data_ptr())
                torch.cuda.synchronize()
                got = pipe.view_numpy(k_, rr) if rr != 0 else pipe.image_numpy(k_)[0]
                ok = ok and bool(np.array_equal(got, solo.cpu().numpy()))
        barrier()
        composite_identical = ok if rank == 0 else None

    # ---- parity of THIS workload against the reference compiled for the host (outside every timed region)
    parity = None
    if rank == 0:
        try:
            from oracle import refbind as rb
            nchk = 5 if WW * HH <= 1920 * 1080 else 2
            idx = sorted(set(int(j * K / nchk) for j in range(nchk)))
            ref, kindname = cpu_frames(R, rb, scene, cfg, [poses[j] for j in idx], os.cpu_count() or 1, keep=True)
            warp_same, rgba_max, ident = True, 0, 1.0
            solo = torch.zeros((HH, WW, 4), dtype=torch.uint8, device="cuda")
            for j, (_, _, owarp, orgba) in zip(idx, ref):
                r.frame_device(raymaps[j], cfg, 1, 1, 0, solo.data_ptr())
                torch.cuda.synchronize()
                n = raymaps[j].map_line_count
                warp_same = warp_same and bool(np.array_equal(r.read_warp(cfg)[:n], owarp[:n]))
                d = np.abs(solo.cpu().numpy().astype(np.int16) - orgba.astype(np.int16))
                rgba_max = max(rgba_max, int(d.max()))
                ident = min(ident, float((d.reshape(-1, 4).max(axis=1) == 0).mean()))
            parity = {"frames_checked": len(idx), "frames": idx, "against": "oracle/_ref (reference render_line compiled for the host)" if kindname == "reference" else "oracle port",
                      "warp_identical": warp_same, "rgba_max_diff": rgba_max, "rgba_identical_fraction": ident}
            del solo
        except Exception as e:                                   # the checker must not take the measurement down
            parity = {"error": repr(e)[:200]}

    # ---- roofline of the traversal kernel: algorithmic bytes from the instrumented kernel (whole-frame semantics)
    roof = None
    if rank == 0 and trav_ms is not None:
        sample_idx = [int(j * K / 8) for j in range(8)] if K >= 8 else list(range(K))
        ids = torch.empty((cfg.rays_casted, cfg.render_size, 2), dtype=torch.int32, device="cuda")
        tot_bytes, tot_rays, agg = 0, 0, {}
        for j in sample_idx:
            r.render_ids(raymaps[j], cfg, ids.data_ptr())
            r.sync()
            c = dict(zip(COUNTER_NAMES, r.counters()))
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
        per_launch_bytes = bytes_per_frame / (world if slices else 1)       # a rank's launch traverses 1/world of the ray planes
        t_launch = trav_ms / KL
        achieved = per_launch_bytes / (t_launch / 1e3) / 1e9
        cap = committed_capture(args.workload, kernel_used) if world == 1 else None
        l2peak = None
        try:
            l2peak = json.load(open(os.path.join(ROOT, "profiles", "l2_peak.json")))
        except Exception:
            pass
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": cap["dram_bytes_per_launch"] if cap else None,
                "kernel": kernel_used + " (+ k_dda_states, overlapped by programmatic dependent launch; both inside the timed interval)",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured, burst: the kernel is timed alone in the latency pass)" if peaks else "fallback 6650 GB/s",
                "algorithmic_bytes_per_launch": per_launch_bytes,
                "traverse_ms_per_launch": t_launch, "unwarp_ms_per_launch": unw_ms / KL,
                "unwarp_achieved_gbs": 8.0 * WW * HH / (max(unw_ms, 1e-9) / KL / 1e3) / 1e9,
                "counters_per_frame": {k_: v / len(sample_idx) for k_, v in agg.items()},
                "plane_rays_per_frame": tot_rays / len(sample_idx)}
        if cap:
            roof["ncu"] = cap.get("ncu")
            if cap.get("l2_bytes_per_launch") and l2peak:
                a = cap["l2_bytes_per_launch"] / (cap["ncu_duration_ms"] / 1e3) / 1e9
                roof["l2"] = {"achieved": a, "peak": l2peak["l2_read_gbs"], "unit": "GB/s", "frac": a / l2peak["l2_read_gbs"],
                              "traffic": cap["l2_bytes_per_launch"],
                              "how": "lts__t_sectors x 32 B of the committed ncu capture / its gpu__time_duration; peak = tools/l2_peak.py on this pool's B200 (profiles/l2_peak.json)"}
        # same-GPU comparator: the reference's own cudaRender compiled for sm_100a (oracle/_ref/libref_cuda.so), config 1
        if world == 1 and args.workload.startswith("imrodh"):
            try:
                from oracle import refbind as rb_
                if rb_.have_ref_cuda():
                    sys.path.insert(0, os.path.join(ROOT, "tools"))
                    import ref_cuda_kernel
                    roof["reference_kernel_on_this_gpu"] = ref_cuda_kernel.compare(R, r, sy, imrodh)
            except Exception as e:
                roof["reference_kernel_on_this_gpu"] = {"error": repr(e)[:200]}

    # ---- e2e through the C ABI with host buffers (pinned), copies inside the timed region
    barrier()
    e2e = None
    sync_latency_ms = None
    if world == 1:
        r2 = R.Renderer(local)
        r2.share_scene(r)
        r2.set_lanes_per_ray(args.lanes)
        pins = [R.PinnedBuffer((HH, WW, 4)) for _ in range(F)]
        for i in range(max(W, 8)):      # every frame slot (stream, warped buffer, RGBA buffer) exists before the timed region
            r2.frame_wait(r2.frame_submit(poses[i % K][0], poses[i % K][1], cfg, pins[i % F].array))
        r2.sync()
        reps = th["repeats"]
        t0 = time.perf_counter()
        tickets = []
        for rep in range(reps):
            for i in range(K):
                if len(tickets) >= F:
                    r2.frame_wait(tickets.pop(0))
                tickets.append(r2.frame_submit(poses[i][0], poses[i][1], cfg, pins[i % F].array))
        for tk in tickets:
            r2.frame_wait(tk)
        r2.sync()
        e2e_wall = time.perf_counter() - t0
        e2e_frames = reps * K
        checksum = int(pins[(K - 1) % F].array[::16, ::16].astype(np.uint32).sum())
        # synchronous single frames through rlerc_render_frame (pose in, pixels in host memory when the call returns)
        ts = []
        for i in seq_frames[:50]:
            t1 = time.perf_counter()
            r2.render_frame(poses[i][0], poses[i][1], cfg, pins[0].array)
            ts.append(1e3 * (time.perf_counter() - t1))
        sync_latency_ms = dist_stats(ts)
        for p in pins:
            p.free()
        r2.close()
        how = "pinned host RGBA out, camera pose in; rlerc_frame_submit/wait, %d frames in flight, one stream per frame slot" % F
    else:
        # ONE host frame ring shared by all ranks; every rank copies its own part of every frame into it over its own PCIe link
        pull = args.compose == "pull"
        rows = HH if (pull or not slices) else MG.band_rows(HH, world) * world
        # frames / views: every rank has its own slots in the shared host ring; keep the ring below 4 GB (8K frames are 133 MB)
        F2 = F if slices else max(2, min(F, int(4e9 / (world * rows * WW * 4))))
        shape = (F2 if slices else F2 * world, rows, WW, 4)
        nbytes = int(np.prod(shape))
        name = [None]
        if rank == 0:
            name[0] = "/dev/shm/rlerc_bench_%d_%d" % (os.getpid(), int(time.time()))
        dist.broadcast_object_list(name, src=0)
        host = torch.from_file(name[0], shared=True, size=nbytes, dtype=torch.uint8).view(shape)
        barrier()
        rc = torch.cuda.cudart().cudaHostRegister(host.data_ptr(), nbytes, 0)
        if int(rc) != 0:
            log("cudaHostRegister failed (%s): the copies into the shared host frame will be synchronous" % rc)
        views_saved = views
        views = False                       # e2e with host buffers: every rank copies ITS frames to the host itself (no gather on one GPU)
        pipe2 = make_pipe(scene, cfg, rank if slices else 0, world if slices else 1, D if slices else None, dst=-1,
                          host=host if slices else host[rank * F2:(rank + 1) * F2], share_from=r, depth=F2)
        views = views_saved
        maps2 = raymaps if slices else (my_maps if views else raymaps[rank::world])
        # host-side frame setup (get_ray_map) is inside the timed region, as in the single-GPU e2e
        for i in range(max(W, 2 * F)):
            pipe2.submit(i, maps2[i % len(maps2)])
        pipe2.drain()
        barrier()
        reps = th["repeats"]
        my_poses = poses if slices else (poses[(rank * K) // world:] + poses[:(rank * K) // world] if views else poses[rank::world])
        t0 = time.perf_counter()
        n = 0
        last_slot = 0
        for rep in range(reps):
            for i in range(len(my_poses)):
                rm = R.RayMap(cfg).get_ray_map(my_poses[i][0], my_poses[i][1])
                last_slot = pipe2.submit(n, rm)
                n += 1
        pipe2.drain()
        barrier()
        e2e_wall = allmax(time.perf_counter() - t0)
        e2e_frames = reps * K * (world if views else 1)
        # the last frame of the job, as it landed in host memory, against rank 0's own single-GPU frame
        checksum = 0
        if slices:
            if rank == 0:
                solo = torch.zeros((HH, WW, 4), dtype=torch.uint8, device="cuda")
                r.frame_device(raymaps[K - 1], cfg, 1, 1, 0, solo.data_ptr())
                torch.cuda.synchronize()
                same = bool(torch.equal(host[last_slot, :HH], solo.cpu()))
                composite_identical = bool(composite_identical) and same
                checksum = int(host[last_slot, :HH][::16, ::16].to(torch.int64).sum())
        barrier()
        pipe2.close()
        torch.cuda.cudart().cudaHostUnregister(host.data_ptr())
        del host
        barrier()
        if rank == 0:
            try:
                os.unlink(name[0])
            except OSError:
                pass
        how = (("rlerc_group_submit: per-rank ray-plane slices, every rank unwarps its band of rows (texels pulled over NVLink) and copies it into ONE host frame shared by all ranks (%d PCIe links), %d frames in flight"
                if pull else "per-rank ray-plane slices, NCCL reduce_scatter into row bands, every rank copies its band into ONE host frame shared by all ranks (%d PCIe links), %d frames in flight")
               % (world, F)) if slices else ("whole frames per rank, every rank copies its frames into the shared host ring itself, %d frames in flight" % F)
    e2e_fps = e2e_frames / e2e_wall
    e2e = {"value": WW * HH * e2e_fps / 1e6, "unit": "Mrays/s", "frames_per_s": e2e_fps,
           "h2d_bytes_per_step": 1024 * world, "d2h_bytes_per_step": WW * HH * 4, "how": how, "checksum": checksum,
           "frames": e2e_frames, "wall_s": e2e_wall}
    if sync_latency_ms:
        e2e["synchronous_single_frame_ms"] = dict(sync_latency_ms, how="rlerc_render_frame: pose in, RGBA in pinned host memory when the call returns; wall clock per call, 50 frames")

    # ---- north star: column slices of the 4K tiled scene, N ranks against rank 0 alone, same frames, same run
    north = None
    if slices and not args.no_north_star and args.workload != "tiled4k":
        try:
            pipe.close()
            del scene
            scene4, name4, sy4 = build_scene(R, "tiled4k", log)
            W4, H4 = WORKLOADS["tiled4k"][3]
            cfg4 = R.FrameConfig.default(W4, H4)
            K4 = 24
            maps4 = []
            for i in range(K4):
                p, q = path_pose(R, i, K4, sy4, False)
                rm = R.RayMapGPU()
                C.memmove(C.byref(rm), C.byref(R.RayMap(cfg4).get_ray_map(p, q)), 896)
                maps4.append(rm)
            pn = make_pipe(scene4, cfg4, rank, world, D)
            tn = throughput(pn, maps4, K4, 0.4)
            flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
            barrier()
            per, _, _ = sequential_job(torch, pn, maps4, list(range(K4)), flush)
            barrier()
            seq_n = allmax(sum(per)) / K4
            ident = None
            k_ = pn.submit(0, maps4[0])
            pn.finish_slot(k_)
            pn.drain()
            barrier()
            if rank == 0:
                got, _, _ = pn.image_numpy(k_)
                solo = torch.zeros((H4, W4, 4), dtype=torch.uint8, device="cuda")
                pn.r[0].frame_device(maps4[0], cfg4, 1, 1, 0, solo.data_ptr())
                torch.cuda.synchronize()
                ident = bool(np.array_equal(got, solo.cpu().numpy()))
                del got, solo
            barrier()
            # rank 0 alone on the same frames (the other ranks wait at the barrier)
            t1 = None
            seq_1 = None
            if rank == 0:
                p1 = make_pipe(scene4, cfg4, 0, 1, None, share_from=pn.r[0])
                nb = (lambda: torch.cuda.synchronize())
                for i in range(2 * F):
                    p1.submit(i, maps4[i % K4])
                p1.drain()
                nb()
                a, _ = pipelined_job(torch, p1, maps4, K4, 1, nb)
                reps1 = int(max(1, math.ceil(400.0 / max(a, 1e-3))))
                tot1, _ = pipelined_job(torch, p1, maps4, K4, reps1, nb)
                t1 = tot1 / (reps1 * K4)
                per1, _, _ = sequential_job(torch, p1, maps4, list(range(K4)), flush)
                seq_1 = sum(per1) / K4
                p1.close()
            barrier()
            del flush
            if rank == 0:
                north = {"workload": "tiled4k", "scene": name4, "scene_mb": scene4.nbytes() / 1e6, "window": [W4, H4], "frames": K4,
                         "split": "interleaved blocks of %d ray planes over %d GPUs, full replica each, %s" % (args.slice_block, world, "unwarp bands with NVLink pull, frame assembled on rank 0" if args.compose == "pull" else "NCCL reduce to rank 0"),
                         "throughput": {"frames_in_flight": F, "ms_per_frame_n": tn["ms_per_frame"], "ms_per_frame_1": t1,
                                        "fps_n": 1e3 / tn["ms_per_frame"], "fps_1": 1e3 / t1, "speedup": t1 / tn["ms_per_frame"]},
                         "latency": {"ms_per_frame_n": seq_n, "ms_per_frame_1": seq_1, "speedup": seq_1 / seq_n,
                                     "how": "one frame at a time, L2 flushed before each"},
                         "composite_identical": ident}
            pn.close()
        except Exception as e:
            north = {"error": repr(e)[:300]}

    # ---- CPU baseline on the box's host cores (rank 0, N=1 only), bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import refbind as rb
        threads = os.cpu_count() or 1
        idx = [int(j * K / 6) for j in range(6)] if K >= 6 else list(range(K))
        sample = [poses[j] for j in idx]
        cpu_frames(R, rb, scene, cfg, sample[:1], threads)
        tms, kindname = cpu_frames(R, rb, scene, cfg, sample, threads)
        one, _ = cpu_frames(R, rb, scene, cfg, sample[:2], 1)
        per = sum(a + b for a, b in tms) / len(tms)
        trav = sum(a for a, _ in tms) / len(tms)
        per1 = sum(a + b for a, b in one) / len(one)
        pern = sum(a + b for a, b in tms[:2]) / 2
        cpu = {"value": WW * HH / per / 1e6, "unit": "Mrays/s", "cores": threads, "kind": kindname,
               "frames_per_s": 1.0 / per, "traversal_only_ms_per_frame": 1e3 * trav, "traversal_only_value": WW * HH / trav / 1e6,
               "sample": "%d of the %d fly-through frames: reference render_line compiled for the host (OpenMP, %d threads, %.1f ms/frame) "
                         "+ oracle unwarp (%.1f ms/frame); 1 thread: %.1f ms/frame -> 1->%d thread speed-up %.1fx"
                         % (len(tms), K, threads, 1e3 * trav, 1e3 * sum(b for _, b in tms) / len(tms), 1e3 * per1, threads, per1 / pern)}

    if rank == 0:
        line = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": th["ms_per_frame"] if (slices or world == 1) else th["ms_total"] / frames_done,
                "mp": args.mp if world > 1 else None,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic" if not imrodh else "Imrodh.rle4",
                "frames_per_s": fps, "config": make_config(args.workload, scene_name, cfg),
                "run": {"scene_mb": scene_megabytes, "frames_in_flight": F, "repeats": th["repeats"], "frames_timed": frames_done, "timed_region_ms": th["ms_total"],
                        "per_repeat_ms": th["per_repeat_ms"], "wall_s_timed_region": wall, "lanes": args.lanes, "kernel": kernel_used,
                        "l2": "not flushed inside the pipelined job: a frame touches the scene (%.0f MB) + its own warped buffer (%.0f MB) + DDA states, "
                              "%d such frames are in flight against a 126 MB L2; the flushed, one-frame-at-a-time figure is `latency`"
                              % (scene_megabytes, cfg.rays_casted * cfg.render_size * 4 / 1e6, F),
                        "parallelism": ("1 GPU" if world == 1 else
                                        ("every frame split into ray-plane slices x%d (interleaved blocks of %d), full replica per GPU, %s"
                                         % (world, args.slice_block, "every GPU unwarps its band of rows with texels pulled from the owners over NVLink peer memory and pushes the pixels into rank 0's frame; flag barriers in peer memory, no collective moves pixel data"
                                            if args.compose == "pull" else "one NCCL reduce per frame to rank 0") if slices else
                                         ("one camera per GPU (the path shifted by rank * K / %d frames), whole frames; every finished frame copied into rank 0's view array over NVLink peer memory, flag barrier per batch (BASELINE config 5)" % world
                                          if views else "whole frames dealt round-robin to %d GPUs (full replica each); frames stay on the GPU that rendered them" % world)))},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "latency": lat, "roofline": roof, "parity": parity}
        if composite_identical is not None:
            line["composite_identical"] = composite_identical
        if north is not None:
            line["north_star"] = north
        if cpu:
            line["cpu_baseline"] = cpu
            line["vs_cpu"] = {"e2e_over_cpu": e2e["value"] / cpu["value"],
                              "traversal_only": {"cpu_ms": cpu["traversal_only_ms_per_frame"], "gpu_ms": roof["traverse_ms_per_launch"] if roof else None,
                                                 "ratio": cpu["traversal_only_ms_per_frame"] / roof["traverse_ms_per_launch"] if roof else None,
                                                 "note": "reference render_line on the host vs k_dda_states + traversal kernel, one frame at a time: no builder-authored unwarp on either side"}}
        print(json.dumps(line), file=json_out, flush=True)
    barrier()
    if dist.is_initialized():
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
