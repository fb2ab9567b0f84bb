#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_q4_gpu.py -m gpu -x -q -k "gemm_large or shapes or batch" 2>&1 | tail -25
echo "--- model tests"
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -x -q 2>&1 | tail -8
