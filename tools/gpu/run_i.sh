#!/bin/bash
# bench lines as the driver launches them: N=1 (plain python) and N=$1 (torchrun)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
N=${1:-2}
( timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/i_bench_n1.json 2> gpurun_out/i_bench_n1.err; echo "rc=$?" >> gpurun_out/i_bench_n1.err
( timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 ) > gpurun_out/i_bench_n$N.json 2> gpurun_out/i_bench_n$N.err; echo "rc=$?" >> gpurun_out/i_bench_n$N.err
tail -4 gpurun_out/i_bench_n1.err; tail -8 gpurun_out/i_bench_n$N.err
python - <<PY
import json
for f in ("gpurun_out/i_bench_n1.json", "gpurun_out/i_bench_n$N.json"):
    try:
        d = json.loads(open(f).read().strip().split("\n")[-1])
        print(f, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "lat", round(d["latency"]["mean"], 3) if d.get("latency") else None,
              "composite", d.get("composite_identical"), "parity", d.get("parity"), "\n   north", json.dumps(d.get("north_star"))[:900], "\n   repeats", d["run"]["per_repeat_ms"], d["run"]["kernel"])
    except Exception as e:
        print(f, "unreadable", e)
PY
