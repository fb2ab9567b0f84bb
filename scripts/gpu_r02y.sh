#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "=== tests (tiny model)"
timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -x -q 2>&1 | tail -3
echo "=== trace B=8, previous build (051a7c0)"
VOX_LIB_PATH=$PWD/build_ab/libvoxtral_F.so timeout 200 python scripts/mega_trace.py --streams 8 > gpurun_out/mega_trace_r02y_F_b8.txt 2>&1; head -1 gpurun_out/mega_trace_r02y_F_b8.txt; grep attn gpurun_out/mega_trace_r02y_F_b8.txt
echo "=== trace B=8, tagged chunk states"
timeout 200 python scripts/mega_trace.py --streams 8 > gpurun_out/mega_trace_r02y_b8.txt 2>&1; cat gpurun_out/mega_trace_r02y_b8.txt
echo "=== trace B=1"
timeout 200 python scripts/mega_trace.py --streams 1 > gpurun_out/mega_trace_r02y_b1.txt 2>&1; head -1 gpurun_out/mega_trace_r02y_b1.txt; grep attn gpurun_out/mega_trace_r02y_b1.txt
echo "=== golden + stream + reference-python tests"
timeout 900 python -m pytest tests/test_golden_gpu.py tests/test_stream_gpu.py tests/test_reference_py_gpu.py -m gpu -x -q 2>&1 | tail -5
echo "=== scans on the new build (B=8): key chunks, experiment flags"
for nc in 1 2; do VOX_MEGA_NC=$nc timeout 100 python scripts/mega_trace.py --streams 8 2>&1 | grep -E "^B=|attn" | tr '\n' ' '; echo " [NC=$nc]"; done
for f in 1 2 4 32; do VOX_MEGA_FLAGS=$f timeout 100 python scripts/mega_trace.py --streams 8 2>&1 | head -1 | tr '\n' ' '; echo " [flags=$f]"; done
