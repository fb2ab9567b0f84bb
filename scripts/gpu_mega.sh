#!/bin/bash
# GPU check of the persistent decode-step kernel: tiny-model parity first (bounded), then the model /
# golden suites, then the bench line.  Usage: bash scripts/gpu_mega.sh <tag>
TAG=${1:-mega}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_model_gpu.py -x -q -k "persistent" 2>&1 | tail -25
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "persistent-kernel tests failed: stopping"; exit 1; fi
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_golden_gpu.py -x -q 2>&1 | tail -15
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench exit $?"; tail -3 gpurun_out/bench_${TAG}.err; cat gpurun_out/bench_${TAG}.json
