#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 120 ./scripts/mg_pair_bench > gpurun_out/mg_pair_bench_r02q.txt 2>&1; cat gpurun_out/mg_pair_bench_r02q.txt
echo "=== HEAD build"
VOX_LIB_PATH=$PWD/build_ab/libvoxtral_head.so timeout 300 python scripts/mega_trace.py --streams 8 > gpurun_out/mega_trace_r02q_head_b8.txt 2>&1; cat gpurun_out/mega_trace_r02q_head_b8.txt
echo "=== new build (noinline loop, rotating reducers)"
timeout 300 python scripts/mega_trace.py --streams 8 > gpurun_out/mega_trace_r02q_new_b8.txt 2>&1; cat gpurun_out/mega_trace_r02q_new_b8.txt
echo "=== new build, flag 128 (fixed reducer set)"
VOX_MEGA_FLAGS=128 timeout 300 python scripts/mega_trace.py --streams 8 > gpurun_out/mega_trace_r02q_new_b8_f128.txt 2>&1; cat gpurun_out/mega_trace_r02q_new_b8_f128.txt
echo "=== tests (new build)"
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -x -q 2>&1 | tail -5
