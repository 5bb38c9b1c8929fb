#!/bin/bash
# final tree, 8 GPUs
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
run() { N=$1; tag=$2; shift 2
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 "$@" ) > gpurun_out/g8_${tag}_n$N.json 2> gpurun_out/g8_${tag}_n$N.err; echo "$tag N=$N rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/g8_${tag}_n$N.json").read().strip().split("\n")[-1])
    print("$tag N=$N value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "lat", round(d["latency"]["mean"], 3) if d.get("latency") else None, "composite", d.get("composite_identical"), d["run"]["kernel"], d["run"]["frames_in_flight"], d["run"]["per_repeat_ms"])
    ns = d.get("north_star")
    if ns: print("   north_star", ns.get("throughput"), ns.get("latency"), ns.get("composite_identical"), ns.get("error"))
except Exception as e:
    print("$tag unreadable", e)
PY
}
run 8 default
run 8 views8k --mp views --workload view8k --no-north-star
( timeout 600 ./examples/headless --synth 256 --size 1920 1080 --frames 800 --gpus 8 --out gpurun_out/g8_headless ) > gpurun_out/g8_headless.log 2>&1; tail -3 gpurun_out/g8_headless.log
tail -3 gpurun_out/g8_views8k_n8.err
