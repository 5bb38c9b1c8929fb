#!/bin/bash
# 8-GPU box: multi-GPU tests, bench lines at N = 8 and 4 (as the driver launches them), group probe on the 4K tiled scene
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/j_topo.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q -k "group or multi_gpu or two_devices" ) > gpurun_out/j_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/j_pytest.log
for N in 8 4; do
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 ) > gpurun_out/j_bench_n$N.json 2> gpurun_out/j_bench_n$N.err; echo "rc=$?" >> gpurun_out/j_bench_n$N.err
done
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/group_probe.py tiled4k 24 4 ) > gpurun_out/j_probe_tiled4k_n8.log 2>&1
tail -5 gpurun_out/j_pytest.log; tail -3 gpurun_out/j_bench_n8.err; grep -h "group_probe\|MISMATCH" gpurun_out/j_probe_tiled4k_n8.log
python - <<PY
import json
for f in ("gpurun_out/j_bench_n8.json", "gpurun_out/j_bench_n4.json"):
    try:
        d = json.loads(open(f).read().strip().split("\n")[-1])
        print(f, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "lat", round(d["latency"]["mean"], 3) if d.get("latency") else None,
              "composite", d.get("composite_identical"), "\n   north", json.dumps(d.get("north_star"))[:900], "\n   repeats", d["run"]["per_repeat_ms"], d["run"]["kernel"])
    except Exception as e:
        print(f, "unreadable", e)
PY
