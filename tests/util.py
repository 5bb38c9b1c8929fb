"""Shared helpers for the parity tests."""
import ctypes as C
import hashlib
import math

import numpy as np

# SURVEY.md §8d parity camera grid: pitch x yaw (+ pi/2), covering 1-4 quadrant cases,
# vanishing point inside/outside the screen and the `reverse` branches
PITCHES = (-1.2, -0.6, -0.01, 0.01, 0.4, 1.2, 3.0)
YAWS = (0.01, 0.8, 2.0, 4.5)


def camera_grid(height):
    for p in PITCHES:
        for y in YAWS:
            yield (10000.0, height, 10000.0), (p, y + math.pi / 2, 0.0)


def few_cameras(height):
    return [((10000.0, height, 10000.0), (0.40, 0.30 + math.pi / 2, 0.0)),
            ((10000.5, height * 0.5, 9999.25), (0.05, 2.0 + math.pi / 2, 0.0)),
            ((9000.0, height, 12000.0), (-0.6, 0.8 + math.pi / 2, 0.0)),
            ((10000.0, height * 2, 10000.0), (1.2, 4.5 + math.pi / 2, 0.0)),
            ((512.0, 700.0, 77.0), (0.3, 1.0, 0.0))]     # camera below the volume top by > 512: y_map_switch path


def levels_of(scene):
    return [scene.level(m) for m in range(scene.nummaps)]


def oracle_raymap(rb, rm_product, scene):
    """The product's ray map as the oracle's struct, with host Map4 levels attached."""
    rm = rb.RayMapGPU()
    C.memmove(C.byref(rm), C.byref(rm_product), 896)
    keep = levels_of(scene)
    rb.attach_host_scene(rm, keep)
    rm._keep = keep
    return rm


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def rgb_parity(a, b):
    """(max abs channel difference, fraction of pixels identical in all four channels)."""
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    same = (d.reshape(-1, a.shape[-1]).max(axis=1) == 0).mean()
    return int(d.max()), float(same)


def scene_from_bool(R, solid, seed=0):
    """RLE4 from a boolean occupancy array solid[z][y][x] (y = 0 is the top of the volume), random material bits."""
    sz, sy, sx = solid.shape
    v = np.packbits(np.ascontiguousarray(solid).reshape(-1), bitorder="little")
    rng = np.random.default_rng(seed)
    c1 = rng.integers(0, 256, v.size, dtype=np.uint8)
    c2 = rng.integers(0, 256, v.size, dtype=np.uint8)
    return R.RLE4.compress_all(v, sx, sy, sz, c1, c2)


def edge_scenes(R):
    """Degenerate and adversarial occupancies the synthetic terrains never produce: nothing, everything, one voxel,
    white noise (irregular columns, many runs, top-attached spans), floating layers, thin walls, a non-cubic grid."""
    rng = np.random.default_rng(7)
    n = 64
    out = {}
    out["air"] = np.zeros((n, n, n), bool)
    out["solid"] = np.ones((n, n, n), bool)
    one = np.zeros((n, n, n), bool); one[20, 40, 33] = True
    out["one_voxel"] = one
    out["noise50"] = rng.random((n, n, n)) < 0.5
    out["noise03"] = rng.random((n, n, n)) < 0.03
    layers = np.zeros((n, n, n), bool); layers[:, 8:10, :] = True; layers[:, 30:31, :] = True; layers[:, 60:, :] = True
    layers[10:20, :, 10:20] = False                                   # a shaft through all layers
    out["layers"] = layers
    walls = np.zeros((n, n, n), bool); walls[::8, 16:, :] = True; walls[:, 16:, ::16] = True
    out["walls"] = walls
    out["noncubic"] = rng.random((128, 32, 64)) < 0.2                  # sz=128, sy=32, sx=64
    comb = np.zeros((n, n, n), bool); comb[:, 1::2, :] = True          # every column: 32 runs of one voxel
    out["comb"] = comb
    return {k: scene_from_bool(R, v, seed=i) for i, (k, v) in enumerate(out.items())}


def edge_cameras():
    """For the 64-voxel edge scenes: far outside over the infinite tiling, just above the top, inside the volume
    (inside matter for the solid scenes, inside the shaft of `layers`), straight down, straight up, lattice-aligned."""
    return [((10000.0, -80.0, 10000.0), (0.40, 0.30 + math.pi / 2, 0.0)),
            ((15.3, -1.5, 14.2), (0.9, 2.2, 0.0)),
            ((14.5, 35.25, 15.5), (-0.2, 0.7, 0.0)),
            ((33.0, -20.0, 20.0), (math.pi / 2, 0.0, 0.0)),
            ((40.0, 70.0, 12.0), (-1.4, 4.0, 0.0)),
            ((64.0, -64.0, 128.0), (0.0, math.pi, 0.0))]
