"""TEST INFRASTRUCTURE — ctypes bindings for the oracle libraries.

* ``oracle/_ref/libref_*.so``  — the UNMODIFIED reference compiled as host C++
  (``oracle/Makefile`` target ``ref``; glue in ``ref_render_tu.cpp`` / ``ref_host_tu.cpp``).
* ``oracle/librlerc_oracle.so`` — our CPU restatement (``rlerc_oracle.cpp``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` arm may import this module.  The product package never does.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


class Vec3f(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class Map4(C.Structure):
    """R/src/Rle4.h:7-21 on LP64 (32 bytes)."""
    _fields_ = [("sx", C.c_int), ("sy", C.c_int), ("sz", C.c_int), ("slabs_size", C.c_int),
                ("map", C.c_void_p), ("slabs", C.c_void_p)]


class RayMapGPU(C.Structure):
    """R/src/RayMap.h:16-54 on LP64 (896 bytes)."""
    _fields_ = [
        ("vanishing_point_2d", Vec3f),
        ("map_line_count", C.c_int),
        ("map_line_limit", C.c_int),
        ("rotation", Vec3f),
        ("position", Vec3f),
        ("border", C.c_float),
        ("clip_min", C.c_float),
        ("clip_max", C.c_float),
        ("map4_gpu", Map4 * 16),
        ("nummaps", C.c_int),
        ("maxres", C.c_int),
        ("res", C.c_int * 4),
        ("p4", Vec3f),
        ("p_2d", Vec3f * 8),
        ("p_no", Vec3f * 8),
        ("to3d", C.c_float * 16),
        ("p_ofs_min", C.c_float * 4),
        ("p_ofs_max", C.c_float * 4),
    ]


assert C.sizeof(Map4) == 32 and C.sizeof(RayMapGPU) == 896


def have_ref():
    return all(os.path.exists(os.path.join(REF_DIR, n))
               for n in ("libref_render.so", "libref_render_bench.so", "libref_host.so"))


def have_ref_cuda():
    """The reference's own CUDA kernel compiled for sm_100a (oracle/ref_cuda_tu.cu): the same-GPU comparator."""
    return os.path.exists(os.path.join(REF_DIR, "libref_cuda.so"))


def ref_cuda_frame(rm_device_maps, d_out, repeats=5, nofma=False):
    """cudaRender (R/src/Cuda_Main.cu:150-181,183-271) on the current CUDA device.  rm_device_maps: RayMapGPU whose
    map4_gpu hold DEVICE pointers; d_out: device pointer of 4096 x 1024 uint32.  Returns (ms kernel, ms whole call)."""
    lib = _load("libref_cuda_nofma.so" if nofma else "libref_cuda.so")
    lib.refcuda_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    k, c = C.c_float(), C.c_float()
    rc = lib.refcuda_frame(C.byref(rm_device_maps), d_out, repeats, C.byref(k), C.byref(c))
    if rc:
        raise RuntimeError("refcuda_frame rc=%d" % rc)
    return k.value, c.value


_libs = {}


def _load(name):
    if name not in _libs:
        path = os.path.join(REF_DIR, name) if name.startswith("libref_") else os.path.join(HERE, name)
        lib = C.CDLL(path, mode=os.RTLD_LOCAL if hasattr(os, "RTLD_LOCAL") else 0)
        _libs[name] = lib
    return _libs[name]


def ref_host():
    lib = _load("libref_host.so")
    if not getattr(lib, "_typed", False):
        lib.ref_scene_load.restype = C.c_void_p
        lib.ref_scene_load.argtypes = [C.c_char_p]
        lib.ref_scene_save.argtypes = [C.c_void_p, C.c_char_p]
        lib.ref_scene_free.argtypes = [C.c_void_p]
        lib.ref_scene_nummaps.argtypes = [C.c_void_p]
        lib.ref_scene_map4.restype = C.POINTER(Map4)
        lib.ref_scene_map4.argtypes = [C.c_void_p, C.c_int]
        lib.ref_tree_new.restype = C.c_void_p
        lib.ref_tree_new.argtypes = [C.c_int] * 4
        lib.ref_tree_free.argtypes = [C.c_void_p]
        lib.ref_tree_set_color.argtypes = [C.c_void_p, C.c_int]
        lib.ref_tree_sphere.argtypes = [C.c_void_p] + [C.c_float] * 4 + [C.c_int]
        lib.ref_tree_cube.argtypes = [C.c_void_p] + [C.c_float] * 6
        for f in (lib.ref_tree_voxel, lib.ref_tree_col1, lib.ref_tree_col2):
            f.restype = C.c_void_p
            f.argtypes = [C.c_void_p]
        lib.ref_compress_all.restype = C.c_void_p
        lib.ref_compress_all.argtypes = [C.c_void_p]
        lib.ref_get_ray_map.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, C.c_int, C.c_void_p]
        lib._typed = True
    return lib


def ref_render(bench=False, flags=0):
    """flags: 1 = built with -DCLIPREGION, 2 = -DHEIGHT_COLOR (R/src/core.h:18,22), 3 = both;
    "centerseg" / "normalclip" = built with the alternative culling mode -DCENTERSEG / -DNORMALCLIP (core.h:27-30)."""
    name = {0: "libref_render.so", 1: "libref_render_clip.so", 2: "libref_render_hc.so", 3: "libref_render_cliphc.so",
            "centerseg": "libref_render_centerseg.so", "normalclip": "libref_render_normalclip.so"}[flags]
    lib = _load("libref_render_bench.so" if bench else name)
    if not getattr(lib, "_typed", False):
        lib.ref_render_frame.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                         C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
        lib._typed = True
    return lib


class RefScene:
    """A scene held by the reference's RLE4 (Rle4.cpp). Exposes numpy views of map/slabs."""

    def __init__(self, handle):
        self.h = handle
        self.lib = ref_host()

    @classmethod
    def load(cls, path):
        h = ref_host().ref_scene_load(path.encode())
        if not h:
            raise FileNotFoundError(path)
        return cls(h)

    def save(self, path):
        self.lib.ref_scene_save(self.h, path.encode())

    @property
    def nummaps(self):
        return self.lib.ref_scene_nummaps(self.h)

    def map4(self, m):
        return self.lib.ref_scene_map4(self.h, m).contents

    def level(self, m, map_words_per_col=2):
        """(sx, sy, sz, map uint32[sz*sx*w], slabs uint16[slabs_size]) as numpy views."""
        m4 = self.map4(m)
        n = m4.sx * m4.sz * map_words_per_col
        mp = np.ctypeslib.as_array(C.cast(m4.map, C.POINTER(C.c_uint32)), shape=(n,))
        sl = np.ctypeslib.as_array(C.cast(m4.slabs, C.POINTER(C.c_uint16)), shape=(m4.slabs_size,))
        return m4.sx, m4.sy, m4.sz, mp, sl

    def fill_raymap(self, rm):
        """What R/src/main.cpp:277-278 does (for all levels, not just 10)."""
        for m in range(self.nummaps):
            src = self.map4(m)
            C.memmove(C.byref(rm.map4_gpu[m]), C.byref(src), C.sizeof(Map4))
        rm.nummaps = self.nummaps

    def free(self):
        if self.h:
            self.lib.ref_scene_free(self.h)
            self.h = None


def ref_get_ray_map(pos, rot, border, rays_res):
    rm = RayMapGPU()
    p = (C.c_float * 3)(*pos)
    r = (C.c_float * 3)(*rot)
    ref_host().ref_get_ray_map(p, r, border, rays_res, C.byref(rm))
    return rm


def ref_render_frame(rm, res, mip_distance=None, z_far=80000, rays=None, threads=0, bench=False, fill=0, flags=0):
    """Run the reference render_line over all rays; returns (warp uint32[rays,res], perf or None)."""
    lib = ref_render(bench, flags)
    nrays = rm.map_line_count if rays is None else rays
    warp = np.full((max(nrays, 1), res), fill, dtype=np.uint32)
    perf = (C.c_longlong * 5)()
    rc = lib.ref_render_frame(C.byref(rm), res, res, mip_distance or res, z_far,
                              warp.ctypes.data_as(C.c_void_p), 0, nrays, threads, perf)
    if rc != 0:
        raise RuntimeError("ref_render_frame rc=%d" % rc)
    return warp, (list(perf) if bench else None)


# ---- our CPU restatement (oracle/rlerc_oracle.cpp) ------------------------------------------

def port():
    lib = _load("librlerc_oracle.so")
    if not getattr(lib, "_typed", False):
        lib.orc_build_map.argtypes = [C.c_void_p, C.c_ulonglong, C.c_int, C.c_int, C.c_void_p]
        lib.orc_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                   C.POINTER(C.c_longlong), C.c_int, C.c_int, C.c_int]
        lib.orc_render_flags.argtypes = lib.orc_render.argtypes + [C.c_int]
        lib.orc_unwarp.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_int, C.c_int]
        lib.orc_get_ray_map.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, C.c_int, C.c_void_p]
        lib._typed = True
    return lib


def have_port():
    return os.path.exists(os.path.join(HERE, "librlerc_oracle.so"))


COUNTER_NAMES = ["elems_total", "elems_processed", "voxels_processed", "elems_rendered", "pixels",
                 "cols_fetched", "run_iters", "cols_nonempty", "cleared", "dda_steps"]


def orc_get_ray_map(pos, rot, border, rays_res):
    rm = RayMapGPU()
    port().orc_get_ray_map((C.c_float * 3)(*pos), (C.c_float * 3)(*rot), border, rays_res, C.byref(rm))
    return rm


def orc_build_map(slabs, sx, sz):
    slabs = np.ascontiguousarray(slabs, dtype=np.uint16)
    mp = np.zeros(sx * sz * 2, np.uint32)
    rc = port().orc_build_map(slabs.ctypes.data, slabs.size, sx, sz, mp.ctypes.data)
    if rc:
        raise RuntimeError("orc_build_map rc=%d" % rc)
    return mp


def orc_render(rm, res, rays_casted, mip_distance=None, z_far=80000, want_ids=False, threads=0, ray_begin=0, ray_end=-1, flags=0):
    """Returns (warp uint32[rays_casted,res], ids uint32[rays_casted,res,2] or None, counters dict)."""
    warp = np.zeros((rays_casted, res), np.uint32)
    ids = np.full((rays_casted, res, 2), 0xffffffff, np.uint32) if want_ids else None
    cnt = (C.c_longlong * 10)()
    # cudaRender renders min(map_line_count, RAYS_CASTED) ray planes (R/src/Cuda_Main.cu:196); the buffer has rays_casted rows
    limit = min(rm.map_line_count, rays_casted)
    ray_end = limit if ray_end < 0 else min(ray_end, limit)
    rc = port().orc_render_flags(C.byref(rm), res, mip_distance or res, z_far, warp.ctypes.data,
                                 ids.ctypes.data if want_ids else None, cnt, ray_begin, ray_end, threads, flags)
    if rc:
        raise RuntimeError("orc_render rc=%d" % rc)
    return warp, ids, dict(zip(COUNTER_NAMES, list(cnt)))


def orc_unwarp(rm, W, H, RS, RC, rays_res, warp, ray_begin=0, ray_end=-1, want_texels=False, shader=0):
    """RGBA8 [H][W][4] (row 0 = top); with want_texels also int32[H][W][2] = (ray row, texel) sampled."""
    warp = np.ascontiguousarray(warp, dtype=np.uint32)
    rgba = np.zeros((H, W, 4), np.uint8)
    tex = np.zeros((H, W, 2), np.int32) if want_texels else None
    port().orc_unwarp_set_texel_output.argtypes = [C.c_void_p]
    port().orc_unwarp_set_texel_output(tex.ctypes.data if want_texels else None)
    port().orc_unwarp_set_shader(shader)          # 0: colorize_buddha_soft.frag, 1: colorize_buddha_soft_2xAA.frag
    port().orc_unwarp(C.byref(rm), W, H, RS, RC, rays_res, warp.ctypes.data, rgba.ctypes.data, ray_begin, ray_end)
    port().orc_unwarp_set_shader(0)
    port().orc_unwarp_set_texel_output(None)
    return (rgba, tex) if want_texels else rgba


def attach_host_scene(rm, levels):
    """Point rm.map4_gpu at host arrays: levels = [(sx, sy, sz, map uint32[], slabs uint16[]), ...].
    What R/src/main.cpp:277-278 does with device pointers."""
    for m, (sx, sy, sz, mp, sl) in enumerate(levels):
        e = rm.map4_gpu[m]
        e.sx, e.sy, e.sz = sx, sy, sz
        e.slabs_size = min(len(sl), 0x7fffffff)
        e.map, e.slabs = mp.ctypes.data, sl.ctypes.data
    rm.nummaps = len(levels)
    return rm


def orc_soft(rgba):
    """GLSL pass 2 (soft.frag) on an RGBA8 image [H][W][4] (row 0 = top)."""
    rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
    H, W = rgba.shape[:2]
    out = np.zeros_like(rgba)
    lib = port()
    lib.orc_soft.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.orc_soft(W, H, rgba.ctypes.data, out.ctypes.data)
    return out
