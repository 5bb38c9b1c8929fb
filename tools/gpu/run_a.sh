#!/bin/bash
# round 2, call A: smoke + parity + timing of the DDA pre-pass kernels
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/a_smoke.log
( timeout 900 python tools/pair_probe.py imrodh1080p ) > gpurun_out/a_pair_1080p.log 2>&1; echo "rc=$?" >> gpurun_out/a_pair_1080p.log
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/a_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/a_pytest.log
( timeout 600 python tools/pair_probe.py tiled4k ) > gpurun_out/a_pair_4k.log 2>&1; echo "rc=$?" >> gpurun_out/a_pair_4k.log
( timeout 300 python tools/ray_profile.py imrodh1080p 0 750 ) > gpurun_out/a_rayprof.log 2>&1; echo "rc=$?" >> gpurun_out/a_rayprof.log
( timeout 600 python bench.py --steps 60 --warmup 5 ) > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "rc=$?" >> gpurun_out/a_bench.err
tail -3 gpurun_out/a_smoke.log; tail -8 gpurun_out/a_pair_1080p.log; tail -5 gpurun_out/a_pytest.log; tail -5 gpurun_out/a_pair_4k.log
