#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
N=${1:-2}
for c in 8 32; do
echo "== CUDA_DEVICE_MAX_CONNECTIONS=$c"
( CUDA_DEVICE_MAX_CONNECTIONS=$c timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/group_probe.py tiled4k 24 16 ) 2>&1 | grep "group_probe\|MISMATCH\|rror" | head -3
( CUDA_DEVICE_MAX_CONNECTIONS=$c timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/group_probe.py imrodh1080p 24 16 ) 2>&1 | grep "group_probe\|MISMATCH\|rror" | head -3
done
