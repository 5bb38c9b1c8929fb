import importlib
import math
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # build what is missing (product library, oracle port, and — when /root/reference is
    # mounted — the compiled reference); the GPU box uses the prebuilt files
    import __graft_entry__ as g
    g.build(quiet=True)


@pytest.fixture(scope="session")
def R():
    return importlib.import_module("rle-based-voxel-raycasting_b200")


@pytest.fixture(scope="session")
def rb():
    from oracle import refbind
    if not refbind.have_port():
        pytest.skip("oracle port not built")
    return refbind


@pytest.fixture(scope="session")
def have_ref(rb):
    return rb.have_ref()


@pytest.fixture(scope="session")
def scene_small(R):
    """64^3 terrain, 6 levels."""
    return R.RLE4.synth(0, 64, 64, 64, seed=1)


@pytest.fixture(scope="session")
def scene_mid(R):
    """128^3 terrain, 7 levels."""
    return R.RLE4.synth(0, 128, 128, 128, seed=1)


@pytest.fixture(scope="session")
def scene_runs(R):
    """128^3 worst-case short-run band: columns with up to ~64 runs (exercises the long-column path)."""
    return R.RLE4.synth(1, 128, 128, 128, seed=42)


@pytest.fixture(scope="session")
def scene_rle(R):
    """256^3 heightfield written straight into RLE (rlerc_synth_rle): crust, cave floors, one column in 8 with the
    32-run short-run band; the small sibling of BASELINE config 4's 16384 x 1024 x 16384 scene."""
    return R.RLE4.synth_rle(256, 256, 256, seed=42, band_every=8)

