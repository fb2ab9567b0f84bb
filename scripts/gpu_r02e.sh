#!/bin/bash
# round-2: streaming tests, whole suite, bench (ours + reference arm)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_stream_gpu.py -m gpu -x -q 2>&1 | tail -25
timeout 300 python scripts/mega_trace.py --streams 8 > gpurun_out/mega_trace_r02e_b8.txt 2>&1; cat gpurun_out/mega_trace_r02e_b8.txt
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02e.json 2> gpurun_out/bench_r02e.err
echo "bench exit $?"; tail -3 gpurun_out/bench_r02e.err; cat gpurun_out/bench_r02e.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02e_reference.json 2>> gpurun_out/bench_r02e.err
echo "reference arm exit $?"; cat gpurun_out/bench_r02e_reference.json
