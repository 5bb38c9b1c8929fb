#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
PKG=rle-based-voxel-raycasting_b200
( timeout 600 python -m pytest tests -m gpu -x -q -k "unwarp" ) 2>&1 | tail -2
for l in librlerc.so $(cd $PKG; ls librlerc_u_*.so); do
  ( RLERC_LIB=$PWD/$PKG/$l timeout 200 python tools/unwarp_times.py ) 2>&1 | tail -1
done | tee gpurun_out/u_times.log
