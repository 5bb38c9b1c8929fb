#!/bin/bash
# round 2, re-entry call: smoke + GPU suite + N=1 bench + chain probe on the current tree
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/g_smi.txt 2>&1
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/g_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/g_smoke.log
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/g_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/g_pytest.log
( timeout 900 python bench.py --steps 100 --warmup 5 ) > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; echo "rc=$?" >> gpurun_out/g_bench.err
( timeout 600 python tools/chain_probe.py imrodh1080p ) > gpurun_out/g_chain_1080p.log 2>&1; echo "rc=$?" >> gpurun_out/g_chain_1080p.log
( timeout 600 python tools/pair_probe.py tiled4k ) > gpurun_out/g_pair_4k.log 2>&1; echo "rc=$?" >> gpurun_out/g_pair_4k.log
tail -3 gpurun_out/g_smoke.log; tail -5 gpurun_out/g_pytest.log; tail -3 gpurun_out/g_bench.err; tail -12 gpurun_out/g_chain_1080p.log; tail -12 gpurun_out/g_pair_4k.log
