#!/bin/bash
# ring stages released right after the warp's loads (before the MMAs) vs after: phase trace + parity
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python scripts/mega_trace.py --streams 8 > gpurun_out/mega_trace_r02l_b8.txt 2>&1; tail -10 gpurun_out/mega_trace_r02l_b8.txt
VOX_MEGA_FLAGS=32 timeout 300 python scripts/mega_trace.py --streams 8 > gpurun_out/mega_trace_r02l_b8_late.txt 2>&1; tail -10 gpurun_out/mega_trace_r02l_b8_late.txt
timeout 300 python scripts/mega_trace.py --streams 1 > gpurun_out/mega_trace_r02l_b1.txt 2>&1; tail -10 gpurun_out/mega_trace_r02l_b1.txt
timeout 300 python scripts/mega_trace.py --streams 4 > gpurun_out/mega_trace_r02l_b4.txt 2>&1; tail -10 gpurun_out/mega_trace_r02l_b4.txt
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_golden_gpu.py -m gpu -x -q -s 2>&1 | grep -E "\[ids\]|passed|failed|Error|error|assert" | tail -20
