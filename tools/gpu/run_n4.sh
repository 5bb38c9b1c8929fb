#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 3 ) > gpurun_out/n4f_bench.json 2> gpurun_out/n4f_bench.err; echo "rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/n4f_bench.json").read().strip().split("\n")[-1])
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "lat", d["latency"]["mean"], "composite", d.get("composite_identical"), d["run"]["frames_in_flight"], d["run"]["per_repeat_ms"])
print(d["north_star"]["throughput"], d["north_star"]["latency"]["speedup"], d["north_star"]["composite_identical"])
PY
