#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_model_gpu.py tests/test_stream_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 100 python scripts/mega_trace.py --streams 8 > gpurun_out/mega_trace_r02last_b8.txt 2>&1; cat gpurun_out/mega_trace_r02last_b8.txt
timeout 100 python scripts/mega_trace.py --streams 1 2>&1 | head -1
timeout 200 python -m pytest tests/test_golden_gpu.py tests/test_reference_py_gpu.py -m gpu -x -q 2>&1 | tail -2
