#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for wl in imrodh1080p tiled4k; do for F in 4 8 16; do
( timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --workload $wl --inflight $F ) > gpurun_out/l_${wl}_f$F.json 2> gpurun_out/l_${wl}_f$F.err
python - <<PY
import json
d = json.loads(open("gpurun_out/l_${wl}_f$F.json").read().strip().split("\n")[-1])
print("$wl F=$F value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "lat", round(d["latency"]["mean"], 3), d["run"]["per_repeat_ms"])
PY
done; done
