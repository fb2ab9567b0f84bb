#!/bin/bash
# round-2: paged KV + per-row positions: whole GPU suite, then traces (all-CTA, warp-level)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "\[ids\]|\[ref-py|passed|failed|Error|error|assert" | tail -20
timeout 300 python scripts/mega_trace_all.py --streams 8 > gpurun_out/mega_trace_all_r02c_b8.txt 2>&1; cat gpurun_out/mega_trace_all_r02c_b8.txt
timeout 300 python scripts/mega_trace.py --streams 8 > gpurun_out/mega_trace_r02c_b8.txt 2>&1; cat gpurun_out/mega_trace_r02c_b8.txt
