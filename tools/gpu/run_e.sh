#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
# ncu --set full of the traversal kernel (default cache control: caches flushed before each pass = the timed condition)
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_traverse_f -s 1 -c 1 -f -o gpurun_out/r02_trav_1080p python tools/profile_frame.py imrodh1080p 0 0 2 ) > gpurun_out/e_ncu1.log 2>&1
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_traverse_f -s 1 -c 1 -f -o gpurun_out/r02_trav_4k python tools/profile_frame.py tiled4k 0 0 2 ) > gpurun_out/e_ncu2.log 2>&1
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dda_states -s 1 -c 1 -f -o gpurun_out/r02_dda_1080p python tools/profile_frame.py imrodh1080p 0 0 2 ) > gpurun_out/e_ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/e_ncu1.log
