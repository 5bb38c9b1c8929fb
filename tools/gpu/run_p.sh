#!/bin/bash
# final-tree profile set, one GPU
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
( timeout 600 $NCU -k regex:k_traverse_f -s 1 -c 1 -o gpurun_out/r02f_trav_1080p python tools/profile_frame.py imrodh1080p 0 65 2 ) > gpurun_out/p_ncu1.log 2>&1
( timeout 600 $NCU -k regex:k_traverse_f -s 1 -c 1 -o gpurun_out/r02f_trav_4k python tools/profile_frame.py tiled4k 0 65 2 ) > gpurun_out/p_ncu2.log 2>&1
( timeout 600 $NCU -k regex:k_unwarp -s 1 -c 1 -o gpurun_out/r02f_unwarp_1080p python tools/profile_frame.py imrodh1080p 0 65 2 ) > gpurun_out/p_ncu3.log 2>&1
( timeout 600 $NCU -k regex:k_unwarp -s 1 -c 1 -o gpurun_out/r02f_unwarp_4k python tools/profile_frame.py tiled4k 0 65 2 ) > gpurun_out/p_ncu4.log 2>&1
( timeout 600 $NCU -k regex:k_traverse_q -s 1 -c 1 -o gpurun_out/r02f_travq_4k_slice8 python tools/profile_slice.py tiled4k 0 8 69 ) > gpurun_out/p_ncu5.log 2>&1
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --min-seconds 0.01 ) > gpurun_out/p_launches.log 2>&1
( timeout 600 python tools/chain_probe.py imrodh1080p ) > gpurun_out/p_chain_1080p.log 2>&1
( CHAIN_LANES=69 timeout 600 python tools/chain_probe.py imrodh1080p ) > gpurun_out/p_chain_1080p_q.log 2>&1
( timeout 600 python tools/unwarp_times.py ) > gpurun_out/p_unwarp_times.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/p_ncu5.log; tail -3 gpurun_out/p_chain_1080p.log gpurun_out/p_chain_1080p_q.log gpurun_out/p_unwarp_times.log; wc -l gpurun_out/r02f_launches.csv
