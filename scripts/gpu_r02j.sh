#!/bin/bash
# 2-CTA cluster multicast of the fragment staging: parity + phase trace, against the default launch
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python scripts/mega_trace.py --streams 8 > gpurun_out/mega_trace_r02j_b8.txt 2>&1; tail -25 gpurun_out/mega_trace_r02j_b8.txt
echo "==== cluster"
VOX_MEGA_CLUSTER=1 timeout 300 python scripts/mega_trace.py --streams 8 > gpurun_out/mega_trace_r02j_b8_cl.txt 2>&1; echo "exit $?"; tail -25 gpurun_out/mega_trace_r02j_b8_cl.txt
VOX_MEGA_CLUSTER=1 timeout 300 python scripts/mega_trace.py --streams 1 > gpurun_out/mega_trace_r02j_b1_cl.txt 2>&1; echo "exit $?"; tail -8 gpurun_out/mega_trace_r02j_b1_cl.txt
VOX_MEGA_CLUSTER=1 timeout 900 python -m pytest tests/test_model_gpu.py tests/test_golden_gpu.py tests/test_stream_gpu.py -m gpu -x -q -s 2>&1 | grep -E "\[ids\]|passed|failed|Error|error|assert" | tail -20
