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
