#!/bin/bash
# round-2: streaming sessions + reducers on the high warps
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_stream_gpu.py -m gpu -x -q 2>&1 | tail -25
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_golden_gpu.py -m gpu -x -q -s 2>&1 | grep -E "\[ids\]|passed|failed|Error|error|assert" | tail -12
timeout 300 python scripts/mega_trace_all.py --streams 8 > gpurun_out/mega_trace_all_r02d_b8.txt 2>&1; head -14 gpurun_out/mega_trace_all_r02d_b8.txt; sed -n 15,70p gpurun_out/mega_trace_all_r02d_b8.txt
timeout 300 python scripts/mega_trace.py --streams 8 > gpurun_out/mega_trace_r02d_b8.txt 2>&1; cat gpurun_out/mega_trace_r02d_b8.txt
