"""Quick GPU probe: parity of the traversal kernel against the compiled reference + timings.
Usage: python tools/gpu_probe.py [size] [W] [H]"""
import ctypes as C
import importlib
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
R = importlib.import_module("rle-based-voxel-raycasting_b200")
from oracle import refbind as rb  # noqa: E402


def ref_raymap(rm_mine, scene):
    """Copy our ray map into the reference struct and attach host Map4 levels."""
    rm = rb.RayMapGPU()
    C.memmove(C.byref(rm), C.byref(rm_mine), 896)
    for m in range(scene.nummaps):
        m4, n = scene.map4(m)
        rm.map4_gpu[m].sx, rm.map4_gpu[m].sy, rm.map4_gpu[m].sz = m4.sx, m4.sy, m4.sz
        rm.map4_gpu[m].slabs_size = m4.slabs_size
        rm.map4_gpu[m].map, rm.map4_gpu[m].slabs = m4.map, m4.slabs
    rm.nummaps = scene.nummaps
    return rm


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    H = int(sys.argv[3]) if len(sys.argv) > 3 else 768
    t0 = time.time()
    scene = R.RLE4.synth(0, size, size, size, seed=1)
    print("scene %d^3: %.1fs, %d levels, %.1f MB" % (size, time.time() - t0, scene.nummaps, scene.nbytes() / 1e6), flush=True)
    cfg = R.FrameConfig.default(W, H)
    r = R.Renderer(0)
    r.all_to_gpu(scene)
    r.set_timing(True)
    cams = [((10000., -size * 0.8, 10000.), (0.40, 0.30 + math.pi / 2, 0.)),
            ((10000., -size * 0.3, 10000.), (0.05, 2.0 + math.pi / 2, 0.)),
            ((10000., -size * 1.5, 10000.), (1.2, 4.5 + math.pi / 2, 0.)),
            ((10000., -size * 0.5, 10000.), (-0.6, 0.8 + math.pi / 2, 0.))]
    for pos, rot in cams:
        rm = R.RayMap(cfg).get_ray_map(pos, rot)
        t0 = time.time()
        ref, _ = rb.ref_render_frame(ref_raymap(rm, scene), cfg.render_size, mip_distance=cfg.mip_distance, z_far=cfg.z_far, rays=cfg.rays_casted)
        tref = time.time() - t0
        ref = ref[:cfg.rays_casted]
        line = "cam %s rays %d ref %.0f ms |" % (rot[:2], rm.map_line_count, tref * 1e3)
        for G in (8, 32, 0):
            r.set_lanes_per_ray(G)
            wp = r.warp_buffer(cfg)
            r.upload(wp, np.zeros((cfg.rays_casted, cfg.render_size), np.uint32))
            r.render(rm, cfg)
            r.sync()
            out = r.read_warp(cfg)
            n = min(len(ref), len(out))
            bad = int((ref[:n] != out[:n]).sum())
            # timing: 5 more launches
            ts = []
            for _ in range(5):
                r.render(rm, cfg)
                r.sync()
                ts.append(r.last_kernel_ms()[0])
            line += " G%d bad=%d %.3f ms |" % (G, bad, min(ts))
        print(line, flush=True)
    # unwarp timing
    r.set_lanes_per_ray(0)
    r.render(rm, cfg)
    r.unwarp(rm, cfg)
    r.sync()
    print("unwarp ms", r.last_kernel_ms()[1])


if __name__ == "__main__":
    main()
