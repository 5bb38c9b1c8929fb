#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q -k "lod_streaming" ) > gpurun_out/v_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/v_pytest.log
tail -25 gpurun_out/v_pytest.log
( timeout 600 python tools/stream_probe.py shortrun16k 40 512 ) > gpurun_out/v_stream_shortrun16k.json 2> gpurun_out/v_stream.err; echo "rc=$?"; cat gpurun_out/v_stream_shortrun16k.json; tail -3 gpurun_out/v_stream.err
