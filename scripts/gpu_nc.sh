#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_golden_gpu.py -x -q 2>&1 | tail -3
echo "== B=8"; timeout 300 python scripts/mega_trace.py --streams 8 2>&1 | tail -9
echo "== B=1"; timeout 300 python scripts/mega_trace.py --streams 1 2>&1 | tail -9 | head -5
