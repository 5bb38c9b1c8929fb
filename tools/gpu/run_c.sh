#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/c_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/c_smoke.log
( timeout 900 python tools/pair_probe.py imrodh1080p ) > gpurun_out/c_pair_1080p.log 2>&1; echo "rc=$?" >> gpurun_out/c_pair_1080p.log
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c_pytest.log
( timeout 600 python tools/pair_probe.py tiled4k ) > gpurun_out/c_pair_4k.log 2>&1; echo "rc=$?" >> gpurun_out/c_pair_4k.log
( timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu ) > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "rc=$?" >> gpurun_out/c_bench.err
tail -3 gpurun_out/c_smoke.log; tail -6 gpurun_out/c_pair_1080p.log; tail -4 gpurun_out/c_pytest.log; tail -5 gpurun_out/c_pair_4k.log
python -c "
import json
d=json.loads(open('gpurun_out/c_bench.json').read().strip().split('\n')[-1])
print({k:d[k] for k in ('value','ms_per_step','frames_per_s','frame_ms')}); print(d['e2e']['value']); r=d['roofline']; print(r['traverse_ms_per_launch'], r['frac'])"
