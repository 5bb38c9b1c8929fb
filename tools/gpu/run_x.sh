#!/bin/bash
# 8-GPU box, the lines of record: multi-GPU tests, bench as the driver launches it at N = 8, 4, 2, config 4 and config 5 at N = 8
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "group or multi_gpu or two_devices or headless_multi" ) > gpurun_out/x_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/x_pytest.log
tail -4 gpurun_out/x_pytest.log
run() { N=$1; tag=$2; shift 2
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 "$@" ) > gpurun_out/x_${tag}_n$N.json 2> gpurun_out/x_${tag}_n$N.err; echo "$tag N=$N rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/x_${tag}_n$N.json").read().strip().split("\n")[-1])
    print("$tag N=$N value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "lat", round(d["latency"]["mean"], 3) if d.get("latency") else None, "composite", d.get("composite_identical"), d["run"]["kernel"], d["run"]["frames_in_flight"], d["run"]["per_repeat_ms"])
    ns = d.get("north_star")
    if ns: print("   north_star", ns.get("throughput"), ns.get("latency"), ns.get("composite_identical"), ns.get("error"))
except Exception as e:
    print("$tag unreadable", e)
PY
}
run 8 default
run 4 default
run 2 default
run 8 shortrun16k --workload shortrun16k --no-north-star
run 8 views8k --mp views --workload view8k --no-north-star
( timeout 600 ./examples/headless --synth 256 --size 1920 1080 --frames 400 --gpus 8 --out gpurun_out/x_headless ) > gpurun_out/x_headless.log 2>&1; tail -3 gpurun_out/x_headless.log
