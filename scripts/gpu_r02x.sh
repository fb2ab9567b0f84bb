#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "=== tests (tiny model)"
timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -x -q 2>&1 | tail -3
echo "=== trace B=8, previous build (ea89fcf)"
VOX_LIB_PATH=$PWD/build_ab/libvoxtral_E.so timeout 200 python scripts/mega_trace.py --streams 8 > gpurun_out/mega_trace_r02x_E_b8.txt 2>&1; head -1 gpurun_out/mega_trace_r02x_E_b8.txt
echo "=== trace B=8, coalesced norm statistics + K-chunked fragment staging"
timeout 200 python scripts/mega_trace.py --streams 8 > gpurun_out/mega_trace_r02x_b8.txt 2>&1; cat gpurun_out/mega_trace_r02x_b8.txt
echo "=== trace B=1"
timeout 200 python scripts/mega_trace.py --streams 1 > gpurun_out/mega_trace_r02x_b1.txt 2>&1; cat gpurun_out/mega_trace_r02x_b1.txt
echo "=== golden + stream + reference-python tests"
timeout 900 python -m pytest tests/test_golden_gpu.py tests/test_stream_gpu.py tests/test_reference_py_gpu.py -m gpu -x -q 2>&1 | tail -5
