#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/group_probe.py small 12 3 ) > gpurun_out/k_probe_small_n8.log 2>&1; echo "rc=$?" >> gpurun_out/k_probe_small_n8.log
grep -h "group_probe\|MISMATCH\|FAILED\|rror\|rc=" gpurun_out/k_probe_small_n8.log | head -20
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 ) > gpurun_out/k_bench_n8.json 2> gpurun_out/k_bench_n8.err; echo "rc=$?" >> gpurun_out/k_bench_n8.err
tail -3 gpurun_out/k_bench_n8.err
python - <<PY
import json
for f in ("gpurun_out/k_bench_n8.json",):
    try:
        d = json.loads(open(f).read().strip().split("\n")[-1])
        print(f, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "lat", round(d["latency"]["mean"], 3) if d.get("latency") else None,
              "composite", d.get("composite_identical"), "\n   north", json.dumps(d.get("north_star"))[:900], "\n   repeats", d["run"]["per_repeat_ms"], d["run"]["kernel"], d["run"]["frames_in_flight"])
    except Exception as e:
        print(f, "unreadable", e)
PY
