"""Frame setup: the product's get_ray_map is byte-exact with the oracle port
everywhere, and with the compiled reference when it is available."""
import ctypes as C
import math

import pytest

from util import camera_grid


@pytest.mark.parametrize("wh", [(1024, 768), (1920, 1080), (3840, 2160), (640, 480)])
def test_raymap_bytes_vs_port(R, rb, wh):
    cfg = R.FrameConfig.default(*wh)
    for pos, rot in camera_grid(-818.0):
        mine = R.RayMap(cfg).get_ray_map(pos, rot)
        orc = rb.orc_get_ray_map(pos, rot, cfg.border, cfg.rays_casted_res)
        assert bytes(mine) == bytes(orc), (wh, rot)
        assert 0 < mine.map_line_count <= cfg.rays_casted
        assert mine.map_line_count == sum(mine.res)


def test_raymap_bytes_vs_compiled_reference(R, rb, have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    n = 0
    for wh in ((1024, 768), (1920, 1080), (7680, 4320)):
        cfg = R.FrameConfig.default(*wh)
        for pos, rot in camera_grid(-818.0):
            mine = R.RayMap(cfg).get_ray_map(pos, rot)
            ref = rb.ref_get_ray_map(pos, rot, cfg.border, cfg.rays_casted_res)
            assert bytes(mine) == bytes(ref), (wh, rot)
            n += 1
    assert n == 84


def test_raymap_quadrant_cases(R):
    """The grid covers 1-, 2-, 3- and 4-quadrant frames (vanishing point in/outside the screen)."""
    cfg = R.FrameConfig.default(1024, 768)
    seen = set()
    for pos, rot in camera_grid(-818.0):
        rm = R.RayMap(cfg).get_ray_map(pos, rot)
        seen.add(sum(1 for r in rm.res if r > 0))
    assert {1, 3, 4} <= seen


def test_frame_setup_argument_errors(R):
    cfg = R.FrameConfig.default(1024, 768)
    cfg.rays_casted_res = 4098
    with pytest.raises(R.RlercError):
        R.RayMap(cfg).get_ray_map((0, 0, 0), (0.1, 0.1, 0))
    assert R.lib().rlerc_frame_setup(None, None, None, None) == -1
