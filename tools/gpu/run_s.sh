#!/bin/bash
# full GPU suite + smoke + N=1 bench on the current tree
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/s_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/s_smoke.log
( timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/s_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s_pytest.log
( timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; echo "rc=$?" >> gpurun_out/s_bench.err
tail -4 gpurun_out/s_smoke.log; tail -4 gpurun_out/s_pytest.log; tail -2 gpurun_out/s_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/s_bench.json").read().strip().split("\n")[-1])
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "lat", d["latency"]["mean"], "sync", d["e2e"].get("synchronous_single_frame_ms", {}).get("mean"))
r = d["roofline"]; print("frac", r["frac"], "trav", r["traverse_ms_per_launch"], "unw", r["unwarp_ms_per_launch"], r.get("reference_kernel_on_this_gpu", {}).get("default_flags"))
print(d["parity"])
PY
