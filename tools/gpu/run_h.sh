#!/bin/bash
# multi-GPU groups: new GPU tests + tools/group_probe.py at N ranks.  usage: run_h.sh N
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
N=${1:-2}
( timeout 900 python -m pytest tests -m gpu -x -q -k "unwarp or group or multi_gpu or two_devices" ) > gpurun_out/h_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/h_pytest.log
for wl in small imrodh1080p tiled4k; do
  ( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/group_probe.py $wl 24 4 ) > gpurun_out/h_probe_${wl}_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/h_probe_${wl}_n$N.log
done
( timeout 300 python tools/group_probe.py tiled4k 24 4 ) > gpurun_out/h_probe_tiled4k_n1.log 2>&1
( timeout 300 python tools/group_probe.py imrodh1080p 24 4 ) > gpurun_out/h_probe_imrodh1080p_n1.log 2>&1
tail -15 gpurun_out/h_pytest.log; grep -h "group_probe\|MISMATCH\|rror\|rank" gpurun_out/h_probe_*.log | head -20
