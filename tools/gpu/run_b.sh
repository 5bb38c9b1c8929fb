#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/b_ktimes.csv python tools/kernel_times.py imrodh1080p 65 ) > gpurun_out/b_ktimes.log 2>&1
( RLERC_PROF_SLICE=16 timeout 300 python tools/ray_profile.py imrodh1080p 0 ) > gpurun_out/b_rayprof16.log 2>&1
( timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu ) > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
grep -v "^==" gpurun_out/b_ktimes.csv | cut -d, -f5,15 | tail -20; cat gpurun_out/b_rayprof16.log | head -12
