#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/group_probe.py tiled4k 24 16 ) 2>&1 | grep "group_probe\|MISMATCH\|rror" | head -3
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/group_probe.py imrodh1080p 24 16 ) 2>&1 | grep "group_probe\|MISMATCH\|rror" | head -3
