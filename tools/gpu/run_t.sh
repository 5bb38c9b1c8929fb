#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "unwarp or headless or scene" ) > gpurun_out/t_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/t_pytest.log
tail -12 gpurun_out/t_pytest.log
