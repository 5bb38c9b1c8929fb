#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "camera_grid or degenerate or history or pipeline" ) 2>&1 | tail -2
( timeout 300 python tools/chain_probe.py imrodh1080p ) 2>&1 | tail -3
( timeout 300 python tools/group_probe.py imrodh1080p 24 8 ) 2>&1 | grep "group_probe"
( timeout 300 python tools/group_probe.py tiled4k 24 16 ) 2>&1 | grep "group_probe"
( timeout 300 python tools/profile_slice.py tiled4k 0 8 65 ) 2>&1 | tail -2
