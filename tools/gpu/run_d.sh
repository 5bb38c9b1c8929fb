#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
LIBS="librlerc_ch16.so librlerc_ch8.so librlerc_w10.so librlerc_m3.so librlerc_w12.so"
( timeout 900 python tools/ab_libs.py imrodh1080p $LIBS ) > gpurun_out/d_ab_1080p.log 2>&1
( timeout 900 python tools/ab_libs.py tiled4k $LIBS ) > gpurun_out/d_ab_4k.log 2>&1
cat gpurun_out/d_ab_1080p.log gpurun_out/d_ab_4k.log
