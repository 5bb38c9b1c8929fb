#!/bin/bash
# final tree, one GPU: the artefacts of record
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
PKG=rle-based-voxel-raycasting_b200
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/f_smoke.log
( timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/f_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/f_pytest.log
( timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "rc=$?" >> gpurun_out/f_bench.err
( timeout 900 python bench.py --impl reference --steps 20 --warmup 3 ) > gpurun_out/f_bench_reference_arm.json 2> gpurun_out/f_bench_ref.err
( timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu --workload tiled4k ) > gpurun_out/f_bench_tiled4k.json 2> /dev/null
( timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu --workload shortrun16k ) > gpurun_out/f_bench_shortrun16k.json 2> /dev/null
( timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu --workload view8k ) > gpurun_out/f_bench_view8k.json 2> /dev/null
NCU="ncu --set full --clock-control none --import-source on -f"
( timeout 600 $NCU -k regex:k_unwarp -s 1 -c 1 -o gpurun_out/r02f_unwarp_1080p python tools/profile_frame.py imrodh1080p 0 65 2 ) > gpurun_out/f_ncu3.log 2>&1
( timeout 600 $NCU -k regex:k_unwarp -s 1 -c 1 -o gpurun_out/r02f_unwarp_4k python tools/profile_frame.py tiled4k 0 65 2 ) > gpurun_out/f_ncu4.log 2>&1
( for wl in imrodh1080p tiled4k; do echo "# tools/pair_probe.py $wl (k65 = k_traverse_f, k68 = k_traverse_p, k69 = k_traverse_q; 1/k = every k-th ray plane only: the uncontended chain)"; timeout 600 python tools/pair_probe.py $wl 2>&1 | tail -5; done
  for wl in tiled4k imrodh1080p; do echo "# RLERC_PROF_Q=1 RLERC_PROF_SLICE=16 tools/quad_profile.py $wl 0 750 (library built with -DRLERC_Q_PROF=1)"; RLERC_LIB=$PWD/$PKG/librlerc_q_prof.so RLERC_PROF_Q=1 RLERC_PROF_SLICE=16 timeout 300 python tools/quad_profile.py $wl 0 750 2>&1 | tail -12; done ) > gpurun_out/f_chain_and_roles.txt 2>&1
( timeout 300 python tools/unwarp_times.py ) > gpurun_out/f_unwarp_times.txt 2>&1
tail -4 gpurun_out/f_smoke.log; tail -4 gpurun_out/f_pytest.log; tail -2 gpurun_out/f_bench.err; tail -1 gpurun_out/f_unwarp_times.txt
python - <<PY
import json
for f in ("f_bench", "f_bench_reference_arm", "f_bench_tiled4k", "f_bench_shortrun16k", "f_bench_view8k"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().split("\n")[-1])
        print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), "lat", (d.get("latency") or {}).get("mean"), "parity", d.get("parity"), (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "unreadable", e)
PY
