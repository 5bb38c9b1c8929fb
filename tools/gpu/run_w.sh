#!/bin/bash
# N GPUs: multi-GPU tests + config 5 (views) and config 4 (shortrun16k slices) bench lines.  usage: run_w.sh N
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
N=${1:-2}
( timeout 900 python -m pytest tests -m gpu -x -q -k "group or multi_gpu or two_devices or headless_multi" ) > gpurun_out/w_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/w_pytest.log
tail -4 gpurun_out/w_pytest.log
run() { tag=$1; shift
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 "$@" ) > gpurun_out/w_${tag}_n$N.json 2> gpurun_out/w_${tag}_n$N.err; echo "$tag rc=$?"
tail -2 gpurun_out/w_${tag}_n$N.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/w_${tag}_n$N.json").read().strip().split("\n")[-1])
    print("$tag value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "lat", round(d["latency"]["mean"], 3) if d.get("latency") else None, "composite", d.get("composite_identical"), d["run"]["kernel"], d["run"]["frames_in_flight"], d["run"]["per_repeat_ms"])
except Exception as e:
    print("$tag unreadable", e)
PY
}
run views_small --mp views --workload small --no-north-star
run views_8k --mp views --workload view8k --no-north-star
run slices_shortrun16k --workload shortrun16k --no-north-star
