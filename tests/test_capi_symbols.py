"""The C-ABI library loads and exports every symbol include/rlerc.h declares (no compute calls)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "rlerc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    funcs = set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", src))
    funcs -= {"defined"}
    data = set(re.findall(r"extern\s+int\s+([A-Za-z_][A-Za-z0-9_]*)\s*;", src))
    return funcs, data


def test_header_declares_the_reference_surface():
    funcs, data = declared_symbols()
    # R/src/Cuda_Main.cu:124-126, R/src/core.h:144-147
    for name in ("cuda_main_render2", "pboRegister", "pboUnregister", "gpu_malloc", "gpu_memcpy", "cpu_memcpy"):
        assert name in funcs
    assert "cpu_to_gpu_delta" in data
    assert len(funcs) > 35


def test_library_exports_every_declared_symbol(R):
    lib = C.CDLL(R.LIB_PATH)
    funcs, data = declared_symbols()
    missing = [n for n in sorted(funcs | data) if not hasattr(lib, n)]
    assert not missing, missing


def test_python_mirror_binds_every_declared_function(R):
    funcs, _ = declared_symbols()
    assert funcs == set(R._SIGS), (funcs ^ set(R._SIGS))


def test_struct_layouts_match_the_reference_abi(R):
    assert C.sizeof(R.Map4) == 32          # R/src/Rle4.h:7-21 on LP64
    assert C.sizeof(R.RayMapGPU) == 896    # R/src/RayMap.h:16-54 on LP64
    assert R.RayMapGPU.map4_gpu.offset == 56 and R.RayMapGPU.to3d.offset == 796
    cfg = R.FrameConfig.default(1024, 768)
    # R/src/core.h:3-10, R/src/main.cpp:774
    assert (cfg.render_size, cfg.rays_casted, cfg.rays_casted_res, cfg.z_far, cfg.mip_distance) == (1024, 4096, 4096, 80000, 1024)
    assert cfg.border == 0.125


def test_compute_fails_loudly_without_a_gpu(R):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(R.RlercError):
        R.Renderer(0)


def test_version_and_error_strings(R):
    assert b"rlerc" in R.lib().rlerc_version()
    p = C.c_void_p()
    assert R.lib().rlerc_scene_load(b"/nonexistent/file.rle4", C.byref(p)) == -2
    assert b"cannot open" in R.lib().rlerc_last_error()


def test_header_is_plain_c99_and_links(tmp_path):
    """include/rlerc.h is the drop-in boundary: it must compile as C (no C++ types in the signatures) and a C program
    must link against librlerc.so and call a host-only entry point."""
    import subprocess
    src = tmp_path / "c99.c"
    src.write_text('#include "rlerc.h"\n'
                   'int main(void){ rlerc_frame_config c; rlerc_frame_config_default(1024, 768, &c);\n'
                   '  return (c.width == 1024 && c.render_size == 1024 && c.rays_casted == 4096 && sizeof(rlerc_raymap) == 896\n'
                   '          && sizeof(rlerc_map4) == 32) ? 0 : 1; }\n')
    exe = str(tmp_path / "c99")
    pkg = os.path.join(ROOT, "rle-based-voxel-raycasting_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", exe,
                    "-L", pkg, "-lrlerc", "-Wl,-rpath," + pkg], check=True)
    assert subprocess.run([exe]).returncode == 0


def test_multi_gpu_and_streaming_entry_points_fail_loudly_on_bad_arguments(R):
    """The round-2 additions to the C ABI (multi-GPU groups, rlerc_create_multi, LOD streaming) return a negative
    rlerc_status with a message for bad arguments — before any CUDA call, so this runs without a GPU."""
    lib = R.lib()
    m = C.c_void_p()
    devs = (C.c_int * 2)(0, 0)
    assert lib.rlerc_create_multi(devs, 2, C.byref(m)) == -1 and b"twice" in lib.rlerc_last_error()
    assert lib.rlerc_create_multi(devs, 0, C.byref(m)) == -1
    assert lib.rlerc_create_multi(devs, 9, C.byref(m)) == -1
    assert lib.rlerc_multi_count(None) == 0 and not lib.rlerc_multi_ctx(None, 0)
    assert lib.rlerc_multi_frame_wait(None, 0) < 0
    g = C.c_void_p()
    cfg = R.FrameConfig.default(640, 480)
    assert lib.rlerc_group_create(None, 0, 2, 4, 32, C.byref(cfg), C.byref(g)) == -1
    assert lib.rlerc_group_submit(None, None, 0, None) == -1
    assert lib.rlerc_group_submit_view(None, None, 0, None) == -1
    assert lib.rlerc_group_export(None, None) == -1 and lib.rlerc_group_connect(None, None) == -1
    assert lib.rlerc_group_wait(None, 0) == -1 and lib.rlerc_group_enable_views(None) == -1
    f3 = (C.c_float * 3)(0, 0, 0)
    assert lib.rlerc_stream_prepare(None, f3, f3, C.byref(cfg), 0, None) == -1
    assert lib.rlerc_scene_upload_streamed(None, None) == -1
    assert lib.rlerc_legacy_adopt(None, C.byref(cfg)) == -1
    assert R.GROUP_BLOB_BYTES == 384        # RLERC_GROUP_BLOB_BYTES
