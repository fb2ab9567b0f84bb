#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "=== 79db9b0 reference"
VOX_LIB_PATH=$PWD/build_ab/libvoxtral_79db9b0.so timeout 300 python scripts/mega_trace.py --streams 8 > gpurun_out/mega_trace_r02s_79_b8.txt 2>&1; head -1 gpurun_out/mega_trace_r02s_79_b8.txt
for f in 0 128 32 160; do
  echo "=== new build flags $f"
  VOX_MEGA_FLAGS=$f timeout 300 python scripts/mega_trace.py --streams 8 > gpurun_out/mega_trace_r02s_new_b8_f$f.txt 2>&1; cat gpurun_out/mega_trace_r02s_new_b8_f$f.txt
done
echo "=== new build B=1 (mega forced on)"
timeout 300 python scripts/mega_trace.py --streams 1 > gpurun_out/mega_trace_r02s_new_b1.txt 2>&1; cat gpurun_out/mega_trace_r02s_new_b1.txt
echo "=== tests (new build)"
timeout 1200 python -m pytest tests/test_model_gpu.py tests/test_golden_gpu.py tests/test_stream_gpu.py -m gpu -x -q 2>&1 | tail -5
