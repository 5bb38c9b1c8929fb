#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( timeout 300 python tools/l2_peak.py ) > gpurun_out/f_l2peak.log 2>&1
( timeout 900 python bench.py --steps 100 --warmup 5 ) > gpurun_out/f_bench1.json 2> gpurun_out/f_bench1.err; echo "rc=$?" >> gpurun_out/f_bench1.err
tail -12 gpurun_out/f_l2peak.log; tail -5 gpurun_out/f_bench1.err; cat gpurun_out/f_bench1.json | head -c 6000
