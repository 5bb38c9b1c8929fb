#!/bin/bash
# k_traverse_q (variant 69): parity on the small suites first (bounded), then timing and the role profile
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_scale.py -m gpu -x -q -k "(camera_grid and 69) or degenerate or core_h or up_to_8k or both_production" ) > gpurun_out/q_pytest2.log 2>&1; echo "rc=$?" >> gpurun_out/q_pytest2.log
tail -3 gpurun_out/q_pytest2.log
PKG=rle-based-voxel-raycasting_b200
for l in librlerc.so $(cd $PKG; ls librlerc_q_[a-o]*.so 2>/dev/null); do
echo "== $l"
( RLERC_LIB=$PWD/$PKG/$l PAIR_CODES=65,69 timeout 600 python tools/pair_probe.py imrodh1080p ) 2>&1 | tail -4
( RLERC_LIB=$PWD/$PKG/$l PAIR_CODES=65,69 timeout 600 python tools/pair_probe.py tiled4k ) 2>&1 | tail -4
done > gpurun_out/q_pair.log 2>&1
cat gpurun_out/q_pair.log
for wl in tiled4k imrodh1080p; do
( RLERC_LIB=$PWD/$PKG/librlerc_q_prof.so RLERC_PROF_Q=1 RLERC_PROF_SLICE=16 timeout 300 python tools/quad_profile.py $wl 0 750 ) 2>&1 | tail -12
done | tee gpurun_out/r_quadprof.log
