#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python scripts/mega_trace.py --streams 8 > gpurun_out/mega_trace_r02o_b8.txt 2>&1; tail -10 gpurun_out/mega_trace_r02o_b8.txt
timeout 300 python scripts/mega_trace.py --streams 2 > gpurun_out/mega_trace_r02o_b2.txt 2>&1; tail -10 gpurun_out/mega_trace_r02o_b2.txt
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_golden_gpu.py tests/test_stream_gpu.py -m gpu -x -q -s 2>&1 | grep -E "\[ids\]|passed|failed|Error|error|assert" | tail -20
