#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
PKG=rle-based-voxel-raycasting_b200
for wl in tiled4k imrodh1080p; do
( RLERC_LIB=$PWD/$PKG/librlerc_q_prof.so RLERC_PROF_Q=1 RLERC_PROF_SLICE=16 timeout 300 python tools/quad_profile.py $wl 0 750 ) 2>&1 | tail -14
done | tee gpurun_out/r_quadprof.log
