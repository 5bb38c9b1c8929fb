"""rlerc — B200-native RLE voxel raycaster frame loop (host-side mirror of the reference surface).

The product is ``librlerc.so`` (hand-written CUDA for sm_100a behind the C ABI of
``include/rlerc.h``).  This module is the thin ctypes layer a Python caller uses; class and
method names follow the reference's C++ (``R/`` = RLE-Raycaster/ in the reference checkout):

=====================  ==========================================================
here                   reference
=====================  ==========================================================
``RLE4``               ``struct RLE4``  R/src/Rle4.h:25-52 (load/save/compress_all/all_to_gpu)
``RayMap``             ``class RayMap`` R/src/RayMap.h:57-418 (set_border/set_ray_limit/get_ray_map)
``Renderer.render``    ``cuda_main_render2`` R/src/Cuda_Main.cu:183-271
``Renderer.unwarp``    GLSL pass 1, R/bin/shader/colorize_buddha_soft.frag + R/src/main.cpp:549-623
=====================  ==========================================================

There is no CPU fallback: every compute call goes to the CUDA library and raises
``RlercError`` if it or a GPU is missing.  PyTorch is optional plumbing (device buffers,
``torch.distributed``); nothing here imports it.
"""
import ctypes as C
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RLERC_LIB") or os.path.join(_HERE, "librlerc.so")   # RLERC_LIB: A/B builds (tools/)
MAX_MAPS = 16


class RlercError(RuntimeError):
    pass


class Vec3f(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class Map4(C.Structure):
    """R/src/Rle4.h:7-21."""
    _fields_ = [("sx", C.c_int32), ("sy", C.c_int32), ("sz", C.c_int32), ("slabs_size", C.c_int32),
                ("map", C.c_void_p), ("slabs", C.c_void_p)]


class RayMapGPU(C.Structure):
    """R/src/RayMap.h:16-54 (896 bytes)."""
    _fields_ = [
        ("vanishing_point_2d", Vec3f), ("map_line_count", C.c_int32), ("map_line_limit", C.c_int32),
        ("rotation", Vec3f), ("position", Vec3f), ("border", C.c_float),
        ("clip_min", C.c_float), ("clip_max", C.c_float),
        ("map4_gpu", Map4 * MAX_MAPS), ("nummaps", C.c_int32), ("maxres", C.c_int32),
        ("res", C.c_int32 * 4), ("p4", Vec3f), ("p_2d", Vec3f * 8), ("p_no", Vec3f * 8),
        ("to3d", C.c_float * 16), ("p_ofs_min", C.c_float * 4), ("p_ofs_max", C.c_float * 4),
    ]


class FrameConfig(C.Structure):
    """The compile-time constants of R/src/core.h:3-10 as run-time fields."""
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("render_size", C.c_int32),
                ("rays_casted", C.c_int32), ("rays_casted_res", C.c_int32), ("z_far", C.c_int32),
                ("mip_distance", C.c_int32), ("border", C.c_float), ("flags", C.c_int32)]
    CLIPREGION, HEIGHT_COLOR, SHADER_2XAA = 1, 2, 4      # R/src/core.h:18,22,12 as run-time switches

    @classmethod
    def default(cls, width, height):
        cfg = cls()
        lib().rlerc_frame_config_default(width, height, C.byref(cfg))
        return cfg


class StreamStats(C.Structure):
    _fields_ = [("resident_bytes", C.c_uint64), ("total_bytes", C.c_uint64), ("uploaded_bytes", C.c_uint64),
                ("evicted_bytes", C.c_uint64), ("chunks_mapped", C.c_int32), ("chunks_unmapped", C.c_int32)]


assert C.sizeof(Map4) == 32 and C.sizeof(RayMapGPU) == 896

_lib = None

# name -> (restype, argtypes); every symbol include/rlerc.h declares
_P = C.c_void_p
_SIGS = {
    "rlerc_frame_config_default": (None, [C.c_int, C.c_int, _P]),
    "rlerc_last_error": (C.c_char_p, []),
    "rlerc_version": (C.c_char_p, []),
    "rlerc_set_host_threads": (C.c_int, [C.c_int]),
    "rlerc_scene_load": (C.c_int, [C.c_char_p, C.POINTER(_P)]),
    "rlerc_scene_save": (C.c_int, [_P, C.c_char_p]),
    "rlerc_scene_from_maps": (C.c_int, [_P, C.c_int, C.POINTER(_P)]),
    "rlerc_scene_free": (None, [_P]),
    "rlerc_scene_nummaps": (C.c_int, [_P]),
    "rlerc_scene_level": (C.c_int, [_P, C.c_int, _P, C.POINTER(C.c_uint64)]),
    "rlerc_scene_compress": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "rlerc_scene_tile": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(_P)]),
    "rlerc_synth_volume": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, _P, _P, _P]),
    "rlerc_synth_rle": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, C.POINTER(_P)]),
    "rlerc_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "rlerc_destroy": (None, [_P]),
    "rlerc_scene_upload": (C.c_int, [_P, _P]),
    "rlerc_scene_device_maps": (C.c_int, [_P, _P, C.POINTER(C.c_int)]),
    "rlerc_scene_share": (C.c_int, [_P, _P]),
    "rlerc_scene_upload_streamed": (C.c_int, [_P, _P]),
    "rlerc_stream_prepare": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float), _P, C.c_int, _P]),
    "rlerc_frame_device": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "rlerc_set_lanes_per_ray": (C.c_int, [_P, C.c_int]),
    "rlerc_has_variants": (C.c_int, []),
    "rlerc_last_kernel": (C.c_char_p, [_P]),
    "rlerc_set_dda_producer": (C.c_int, [_P, C.c_int]),
    "rlerc_set_dda_mode": (C.c_int, [_P, C.c_int]),
    "rlerc_frame_setup": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_float), _P, _P]),
    "rlerc_render": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P]),
    "rlerc_render_ids": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P, _P]),
    "rlerc_render_counters": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "rlerc_unwarp": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int]),
    "rlerc_unwarp_slice": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int]),
    "rlerc_soft": (C.c_int, [_P, _P, _P, _P]),
    "rlerc_render_interleaved": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "rlerc_unwarp_interleaved": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "rlerc_set_stream": (C.c_int, [_P, _P]),
    "rlerc_render_frame": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float), _P, _P, _P]),
    "rlerc_frame_submit": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float), _P, _P]),
    "rlerc_frame_wait": (C.c_int, [_P, C.c_int]),
    "rlerc_sync": (C.c_int, [_P]),
    "rlerc_stream": (_P, [_P]),
    "rlerc_warp_buffer": (C.c_int, [_P, _P, C.POINTER(_P)]),
    "rlerc_memcpy_d2h": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "rlerc_memcpy_h2d": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "rlerc_host_alloc": (C.c_int, [C.POINTER(_P), C.c_size_t]),
    "rlerc_host_free": (None, [_P]),
    "rlerc_last_kernel_ms": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "rlerc_set_timing": (C.c_int, [_P, C.c_int]),
    "rlerc_create_multi": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(_P)]),
    "rlerc_multi_destroy": (None, [_P]),
    "rlerc_multi_count": (C.c_int, [_P]),
    "rlerc_multi_ctx": (_P, [_P, C.c_int]),
    "rlerc_multi_set_depth": (C.c_int, [_P, C.c_int, C.c_int]),
    "rlerc_multi_scene_upload": (C.c_int, [_P, _P]),
    "rlerc_multi_render_frame": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float), _P, _P]),
    "rlerc_multi_frame_submit": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float), _P, _P]),
    "rlerc_multi_frame_wait": (C.c_int, [_P, C.c_int]),
    "rlerc_group_create": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.POINTER(_P)]),
    "rlerc_group_destroy": (None, [_P]),
    "rlerc_group_export": (C.c_int, [_P, _P]),
    "rlerc_group_connect": (C.c_int, [_P, _P]),
    "rlerc_group_submit": (C.c_int, [_P, _P, C.c_int, _P]),
    "rlerc_group_enable_views": (C.c_int, [_P]),
    "rlerc_group_submit_view": (C.c_int, [_P, _P, C.c_int, _P]),
    "rlerc_group_views": (C.c_int, [_P, C.c_int, C.POINTER(_P), C.POINTER(C.c_size_t)]),
    "rlerc_group_wait": (C.c_int, [_P, C.c_int]),
    "rlerc_group_sync": (C.c_int, [_P]),
    "rlerc_group_image": (C.c_int, [_P, C.c_int, C.POINTER(_P), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "rlerc_group_stream": (_P, [_P, C.c_int]),
    "rlerc_group_last_ms": (C.c_int, [_P, C.c_int, C.POINTER(C.c_float)]),
    "gpu_malloc": (_P, [C.c_int]),
    "gpu_memcpy": (None, [_P, _P, C.c_int]),
    "cpu_memcpy": (None, [_P, _P, C.c_int]),
    "rlerc_legacy_init": (C.c_int, [C.c_int, _P, _P]),
    "rlerc_legacy_adopt": (C.c_int, [_P, _P]),
    "rlerc_pbo_bind": (C.c_int, [C.c_int, _P]),
    "pboRegister": (None, [C.c_int]),
    "pboUnregister": (None, [C.c_int]),
    "cuda_main_render2": (None, [C.c_int, C.c_int, C.c_int, _P]),
}
EXPORTED_DATA = ("cpu_to_gpu_delta",)


def lib():
    """Load librlerc.so (built in-tree by ``__graft_entry__.build()``). Fails loudly."""
    global _lib
    if _lib is None:
        # RLERC_LIBRARY: an alternative build of the same library (kernel tuning experiments, tools/)
        LIB_PATH = os.environ.get("RLERC_LIBRARY") or globals()["LIB_PATH"]
        if not os.path.exists(LIB_PATH):
            raise RlercError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(there is no CPU fallback)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            f = getattr(l, name)
            f.restype = res
            f.argtypes = args
        _lib = l
    return _lib


def _check(rc):
    if rc < 0:
        raise RlercError("rlerc error %d: %s" % (rc, lib().rlerc_last_error().decode(errors="replace")))
    return rc


def _f3(v):
    return (C.c_float * 3)(float(v[0]), float(v[1]), float(v[2]))


class RLE4:
    """Host-side scene: mip chain of per-column run lists (reference ``struct RLE4``)."""

    def __init__(self, handle=None):
        self._h = C.c_void_p(handle) if handle else C.c_void_p()

    # -- construction --------------------------------------------------------------------
    @classmethod
    def load(cls, filename):
        """RLE4::load (R/src/Rle4.cpp:244-384)."""
        s = cls()
        _check(lib().rlerc_scene_load(os.fsencode(filename), C.byref(s._h)))
        return s

    @classmethod
    def compress_all(cls, voxel, sx, sy, sz, col1=None, col2=None):
        """RLE4::compress_all (R/src/Rle4.cpp:16-50) on a packed bit volume (uint8 array)."""
        voxel = np.ascontiguousarray(voxel, dtype=np.uint8)
        need = (sx * sy * sz + 7) // 8
        if voxel.size < need:
            raise RlercError("bit volume too small")
        p1 = p2 = None
        if col1 is not None:
            col1 = np.ascontiguousarray(col1, dtype=np.uint8)
            col2 = np.ascontiguousarray(col2, dtype=np.uint8)
            p1, p2 = col1.ctypes.data, col2.ctypes.data
        s = cls()
        _check(lib().rlerc_scene_compress(voxel.ctypes.data, p1, p2, sx, sy, sz, C.byref(s._h)))
        return s

    @classmethod
    def from_maps(cls, maps):
        """Deep copy of Map4 levels laid out as RLE4::map[] after load()."""
        arr = (Map4 * len(maps))(*maps)
        s = cls()
        _check(lib().rlerc_scene_from_maps(arr, len(maps), C.byref(s._h)))
        return s

    @classmethod
    def synth(cls, kind, sx, sy, sz, seed=1, color=True):
        """Procedural benchmark scene: volume -> compress_all."""
        n = sx * sy * sz // 8
        v = np.zeros(n, np.uint8)
        c1 = np.zeros(n, np.uint8) if color else None
        c2 = np.zeros(n, np.uint8) if color else None
        _check(lib().rlerc_synth_volume(kind, sx, sy, sz, seed, v.ctypes.data,
                                        c1.ctypes.data if color else None, c2.ctypes.data if color else None))
        return cls.compress_all(v, sx, sy, sz, c1, c2)

    @classmethod
    def synth_rle(cls, sx, sy, sz, seed=42, band_every=32):
        """Heightfield + short-run band written straight into RLE (BASELINE config 4 at full size)."""
        s = cls()
        _check(lib().rlerc_synth_rle(sx, sy, sz, seed, band_every, C.byref(s._h)))
        return s

    def tile(self, nx, nz):
        s = RLE4()
        _check(lib().rlerc_scene_tile(self._h, nx, nz, C.byref(s._h)))
        return s

    # -- access --------------------------------------------------------------------------
    def save(self, filename):
        """RLE4::save (R/src/Rle4.cpp:220-242)."""
        _check(lib().rlerc_scene_save(self._h, os.fsencode(filename)))

    @property
    def nummaps(self):
        return _check(lib().rlerc_scene_nummaps(self._h))

    def map4(self, m):
        out = Map4()
        n64 = C.c_uint64()
        _check(lib().rlerc_scene_level(self._h, m, C.byref(out), C.byref(n64)))
        return out, n64.value

    def level(self, m):
        """(sx, sy, sz, map uint32[sz*sx*2], slabs uint16[n]) as numpy views (borrowed)."""
        m4, n = self.map4(m)
        mp = np.ctypeslib.as_array(C.cast(m4.map, C.POINTER(C.c_uint32)), shape=(m4.sx * m4.sz * 2,))
        sl = np.ctypeslib.as_array(C.cast(m4.slabs, C.POINTER(C.c_uint16)), shape=(n,))
        return m4.sx, m4.sy, m4.sz, mp, sl

    def nbytes(self):
        return sum(self.map4(m)[0].sx * self.map4(m)[0].sz * 8 + self.map4(m)[1] * 2 for m in range(self.nummaps))

    def clear(self):
        """RLE4::clear (R/src/Rle4.cpp:210-218)."""
        if self._h:
            lib().rlerc_scene_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.clear()
        except Exception:
            pass


class RayMap:
    """Per-frame ray map (reference ``class RayMap : RayMap_GPU``)."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.gpu = RayMapGPU()
        self.set_border(cfg.border)
        self.set_ray_limit(cfg.rays_casted_res)

    def set_border(self, a):
        self.cfg.border = a
        self.gpu.border = a
        self.gpu.clip_min = a
        self.gpu.clip_max = 1 - a

    def set_ray_limit(self, a):
        self.gpu.map_line_limit = a

    def get_ray_map(self, pos, rot):
        """RayMap::get_ray_map (R/src/RayMap.h:98-402)."""
        _check(lib().rlerc_frame_setup(_f3(pos), _f3(rot), C.byref(self.cfg), C.byref(self.gpu)))
        return self.gpu

    @property
    def map_line_count(self):
        return self.gpu.map_line_count


class Renderer:
    """One CUDA device: scene replica in HBM + the two kernels of the frame loop."""

    def __init__(self, device=0):
        self._c = C.c_void_p()
        _check(lib().rlerc_create(device, C.byref(self._c)))
        self.device = device

    def close(self):
        if self._c:
            lib().rlerc_destroy(self._c)
            self._c = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # RLE4::all_to_gpu (R/src/Rle4.cpp:432-448)
    def all_to_gpu(self, scene):
        _check(lib().rlerc_scene_upload(self._c, scene._h))

    def all_to_gpu_streamed(self, scene):
        """LOD streaming: reserve address space only; `stream_prepare` makes resident what a camera needs."""
        _check(lib().rlerc_scene_upload_streamed(self._c, scene._h))
        self._streamed_scene = scene            # borrowed by the library: keep it alive

    def stream_prepare(self, pos, rot, cfg, margin_voxels=256):
        st = StreamStats()
        _check(lib().rlerc_stream_prepare(self._c, _f3(pos), _f3(rot), C.byref(cfg), margin_voxels, C.byref(st)))
        return st

    def share_scene(self, other):
        """Use the replica `other` (a Renderer on the same device) uploaded; `other` has to stay alive."""
        _check(lib().rlerc_scene_share(self._c, other._c))
        self._scene_owner = other

    def frame_device(self, raymap_gpu, cfg, block=1, nranks=1, rank=0, d_rgba=None):
        """Traversal + unwarp of this rank's interleaved slice (the whole frame for nranks == 1), asynchronous."""
        _check(lib().rlerc_frame_device(self._c, C.byref(raymap_gpu), C.byref(cfg), block, nranks, rank, d_rgba))

    def device_maps(self):
        arr = (Map4 * MAX_MAPS)()
        n = C.c_int()
        _check(lib().rlerc_scene_device_maps(self._c, arr, C.byref(n)))
        return arr, n.value

    def set_lanes_per_ray(self, lanes):
        _check(lib().rlerc_set_lanes_per_ray(self._c, lanes))

    @property
    def last_kernel(self):
        return lib().rlerc_last_kernel(self._c).decode()

    def set_dda_mode(self, mode):
        _check(lib().rlerc_set_dda_mode(self._c, mode))

    def set_dda_producer(self, on):
        _check(lib().rlerc_set_dda_producer(self._c, 1 if on else 0))

    def set_timing(self, on):
        _check(lib().rlerc_set_timing(self._c, 1 if on else 0))

    def warp_buffer(self, cfg):
        p = C.c_void_p()
        _check(lib().rlerc_warp_buffer(self._c, C.byref(cfg), C.byref(p)))
        return p.value

    def render(self, raymap_gpu, cfg, ray_begin=0, ray_end=-1, d_warp=None):
        """cuda_main_render2 (R/src/Cuda_Main.cu:183-271), asynchronous."""
        _check(lib().rlerc_render(self._c, C.byref(raymap_gpu), C.byref(cfg), ray_begin, ray_end, d_warp))

    def render_ids(self, raymap_gpu, cfg, d_ids, ray_begin=0, ray_end=-1, d_warp=None):
        _check(lib().rlerc_render_ids(self._c, C.byref(raymap_gpu), C.byref(cfg), ray_begin, ray_end, d_warp, d_ids))

    def counters(self):
        out = (C.c_uint64 * 10)()
        _check(lib().rlerc_render_counters(self._c, out))
        return list(out)

    def unwarp(self, raymap_gpu, cfg, d_warp=None, d_rgba=None, row_begin=0, row_end=-1):
        _check(lib().rlerc_unwarp(self._c, C.byref(raymap_gpu), C.byref(cfg), d_warp, d_rgba, row_begin, row_end))

    def unwarp_slice(self, raymap_gpu, cfg, ray_begin, ray_end, d_warp=None, d_rgba=None):
        _check(lib().rlerc_unwarp_slice(self._c, C.byref(raymap_gpu), C.byref(cfg), d_warp, d_rgba, ray_begin, ray_end))

    def soft(self, cfg, d_rgba_in, d_rgba_out):
        """Depth-aware smoothing, GLSL pass 2 (R/bin/shader/soft.frag)."""
        _check(lib().rlerc_soft(self._c, C.byref(cfg), d_rgba_in, d_rgba_out))

    def render_interleaved(self, raymap_gpu, cfg, block, nranks, rank, d_warp=None):
        _check(lib().rlerc_render_interleaved(self._c, C.byref(raymap_gpu), C.byref(cfg), block, nranks, rank, d_warp))

    def unwarp_interleaved(self, raymap_gpu, cfg, block, nranks, rank, d_warp=None, d_rgba=None):
        _check(lib().rlerc_unwarp_interleaved(self._c, C.byref(raymap_gpu), C.byref(cfg), block, nranks, rank, d_warp, d_rgba))

    def set_stream(self, cuda_stream):
        """Launch on a caller-owned stream (int handle, e.g. torch.cuda.current_stream().cuda_stream)."""
        _check(lib().rlerc_set_stream(self._c, cuda_stream))

    def render_frame(self, pos, rot, cfg, host_rgba):
        """get_ray_map -> traversal -> unwarp -> D2H (synchronous). Returns the ray map used."""
        rm = RayMapGPU()
        _check(lib().rlerc_render_frame(self._c, _f3(pos), _f3(rot), C.byref(cfg), _ptr(host_rgba), C.byref(rm)))
        return rm

    def frame_submit(self, pos, rot, cfg, host_rgba):
        return _check(lib().rlerc_frame_submit(self._c, _f3(pos), _f3(rot), C.byref(cfg), _ptr(host_rgba)))

    def frame_wait(self, ticket):
        _check(lib().rlerc_frame_wait(self._c, ticket))

    def sync(self):
        _check(lib().rlerc_sync(self._c))

    @property
    def stream(self):
        return lib().rlerc_stream(self._c)

    def last_kernel_ms(self):
        out = (C.c_float * 2)()
        _check(lib().rlerc_last_kernel_ms(self._c, out))
        return out[0], out[1]

    # -- host convenience (tests, tools) ---------------------------------------------------
    def download(self, dev_ptr, shape, dtype):
        a = np.empty(shape, dtype)
        _check(lib().rlerc_memcpy_d2h(self._c, a.ctypes.data, dev_ptr, a.nbytes))
        return a

    def upload(self, dev_ptr, array):
        a = np.ascontiguousarray(array)
        _check(lib().rlerc_memcpy_h2d(self._c, dev_ptr, a.ctypes.data, a.nbytes))

    def read_warp(self, cfg):
        return self.download(self.warp_buffer(cfg), (cfg.rays_casted, cfg.render_size), np.uint32)


GROUP_BLOB_BYTES = 384


class Group:
    """One member of an N-GPU group (include/rlerc.h "multi-GPU", csrc/group.cu): this GPU traverses its interleaved
    blocks of ray planes and produces its band of window rows, pulling texels from the other members' warped buffers
    over NVLink.  Members in other processes are reached through CUDA IPC: exchange export() blobs, then connect()."""

    def __init__(self, renderer, cfg, rank=0, world=1, depth=4, block=32, views=False):
        self._g = C.c_void_p()
        self.r, self.cfg, self.rank, self.world, self.depth, self.block = renderer, cfg, rank, world, depth, block
        _check(lib().rlerc_group_create(renderer._c, rank, world, depth, block, C.byref(cfg), C.byref(self._g)))
        if views:
            _check(lib().rlerc_group_enable_views(self._g))

    def export(self):
        buf = (C.c_uint8 * GROUP_BLOB_BYTES)()
        _check(lib().rlerc_group_export(self._g, buf))
        return bytes(buf)

    def connect(self, blobs):
        """blobs: the members' export() results concatenated in rank order."""
        blobs = bytes(blobs)
        if len(blobs) != self.world * GROUP_BLOB_BYTES:
            raise RlercError("Group.connect: %d bytes for %d members" % (len(blobs), self.world))
        _check(lib().rlerc_group_connect(self._g, blobs))

    def connect_distributed(self, torch, dist):
        """Exchange the descriptions over torch.distributed (one process per GPU) and connect."""
        if self.world == 1:
            return
        mine = torch.frombuffer(bytearray(self.export()), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            mine = mine.cuda(self.r.device)
        out = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(out, mine)
        self.connect(b"".join(bytes(t.cpu().numpy().tobytes()) for t in out))
        dist.barrier()

    def submit(self, raymap_gpu, dst=0, host=None):
        """Enqueue the next frame (the same call, in the same order, on every member). Returns its ticket."""
        return _check(lib().rlerc_group_submit(self._g, C.byref(raymap_gpu), dst, _ptr(host) if host is not None else None))

    def submit_view(self, raymap_gpu, dst=0, host=None):
        """View batch: this member renders the whole frame of ITS camera and copies it to member dst's view array."""
        return _check(lib().rlerc_group_submit_view(self._g, C.byref(raymap_gpu), dst, _ptr(host) if host is not None else None))

    def views(self, ticket):
        """(device pointer of the ticket's view array on this member, bytes between views)."""
        p, st = C.c_void_p(), C.c_size_t()
        _check(lib().rlerc_group_views(self._g, ticket, C.byref(p), C.byref(st)))
        return p.value, st.value

    def wait(self, ticket):
        _check(lib().rlerc_group_wait(self._g, ticket))

    def sync(self):
        _check(lib().rlerc_group_sync(self._g))

    def image(self, ticket):
        """(device pointer of the ticket's [H][W][4] image, first row, one past the last row this member produces)."""
        p, a, b = C.c_void_p(), C.c_int(), C.c_int()
        _check(lib().rlerc_group_image(self._g, ticket, C.byref(p), C.byref(a), C.byref(b)))
        return p.value, a.value, b.value

    def stream(self, ticket):
        return lib().rlerc_group_stream(self._g, ticket)

    def last_ms(self, ticket):
        ms = C.c_float()
        _check(lib().rlerc_group_last_ms(self._g, ticket, C.byref(ms)))
        return ms.value

    def close(self):
        if self._g:
            lib().rlerc_group_destroy(self._g)
            self._g = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiRenderer:
    """All GPUs of a node behind one object in ONE process (rlerc_create_multi): the multi-GPU drop-in for Renderer's
    render_frame / frame_submit / frame_wait."""

    def __init__(self, devices, depth=4, block=32):
        self._m = C.c_void_p()
        arr = (C.c_int * len(devices))(*devices)
        _check(lib().rlerc_create_multi(arr, len(devices), C.byref(self._m)))
        _check(lib().rlerc_multi_set_depth(self._m, depth, block))
        self.devices = list(devices)

    def all_to_gpu(self, scene):
        _check(lib().rlerc_multi_scene_upload(self._m, scene._h))

    def render_frame(self, pos, rot, cfg, host_rgba):
        _check(lib().rlerc_multi_render_frame(self._m, _f3(pos), _f3(rot), C.byref(cfg), _ptr(host_rgba)))

    def frame_submit(self, pos, rot, cfg, host_rgba):
        return _check(lib().rlerc_multi_frame_submit(self._m, _f3(pos), _f3(rot), C.byref(cfg), _ptr(host_rgba)))

    def frame_wait(self, ticket):
        _check(lib().rlerc_multi_frame_wait(self._m, ticket))

    def close(self):
        if self._m:
            lib().rlerc_multi_destroy(self._m)
            self._m = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _ptr(a):
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return a


class PinnedBuffer:
    """Page-locked host memory (cudaMallocHost) exposed as a numpy array."""

    def __init__(self, shape, dtype=np.uint8):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._p = C.c_void_p()
        _check(lib().rlerc_host_alloc(C.byref(self._p), self.nbytes))
        buf = (C.c_uint8 * self.nbytes).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        if self._p:
            self.array = None
            lib().rlerc_host_free(self._p)
            self._p = C.c_void_p()


def flythrough_pose(t, n=1000):
    """Scripted camera path of BASELINE config 2, frame t of n."""
    a = 2.0 * math.pi * t / n
    pos = (10000.0 + 4000.0 * math.sin(a), -818.0 + 300.0 * math.sin(2 * a), 10000.0 + 4000.0 * math.cos(a))
    rot = (0.35 + 0.3 * math.sin(3 * a), a + math.pi / 2, 0.0)
    return pos, rot
