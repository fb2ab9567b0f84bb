#!/bin/bash
# round-2 first GPU call: new reference-python parity tests + incremental API + the whole suite, then a bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests/test_reference_py_gpu.py tests/test_model_gpu.py -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -25
timeout 1500 python -m pytest tests -m gpu -q -s -x --deselect tests/test_reference_py_gpu.py --deselect tests/test_model_gpu.py 2>&1 | grep -E "\[ids\]|\[ref-py|passed|failed|Error|error|assert" | tail -30
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err
echo "bench exit $?"; tail -3 gpurun_out/bench_r02a.err; cat gpurun_out/bench_r02a.json
