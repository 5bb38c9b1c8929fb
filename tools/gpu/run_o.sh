#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( timeout 600 python tools/ref_cuda_kernel.py gpurun_out/o_refcuda.json ) > gpurun_out/o_refcuda.log 2>&1; echo "rc=$?" >> gpurun_out/o_refcuda.log
tail -40 gpurun_out/o_refcuda.log
