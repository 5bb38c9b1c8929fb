#!/bin/bash
# last check of the final tree on one GPU: smoke, the whole GPU suite, the bench line as the driver runs it
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/v_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/v_smoke.log
( timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/v_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/v_pytest.log
( timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; echo "rc=$?" >> gpurun_out/v_bench.err
tail -2 gpurun_out/v_smoke.log; tail -3 gpurun_out/v_pytest.log; tail -1 gpurun_out/v_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/v_bench.json").read().strip().split("\n")[-1])
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "lat", d["latency"]["mean"], "launches", d["gpu_launches"], "clocks", d["clocks"])
print(d["parity"]); print({k: d["roofline"][k] for k in ("frac", "achieved", "traffic", "traverse_ms_per_launch", "unwarp_ms_per_launch")}, d["roofline"].get("l2"))
print(d["cpu_baseline"]["value"], d["vs_cpu"])
PY
