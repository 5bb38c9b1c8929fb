#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( timeout 2400 python -m pytest tests/test_gpu_bench_scale.py -m gpu -x -q --durations=10 ) > gpurun_out/m_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/m_pytest.log
tail -25 gpurun_out/m_pytest.log
