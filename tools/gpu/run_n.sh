#!/bin/bash
# multi-GPU bench line exactly as the driver launches it: tools/gpu/run_n.sh N [extra bench args]
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
N=$1; shift
( timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 "$@" ) > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err; echo "rc=$?" >> gpurun_out/n${N}_bench.err
( timeout 600 python -m pytest tests -m gpu -x -q -k "two_devices" ) > gpurun_out/n${N}_pytest.log 2>&1
tail -5 gpurun_out/n${N}_bench.err; tail -3 gpurun_out/n${N}_pytest.log; head -c 3000 gpurun_out/n${N}_bench.json
