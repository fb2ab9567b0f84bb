#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 120 ./scripts/mg_pair_bench > gpurun_out/mg_pair_bench.txt 2>&1; cat gpurun_out/mg_pair_bench.txt
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_stream_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 900 python bench.py --steps 3 --warmup 3 --streams 16 --no-cpu-baseline --no-streaming > gpurun_out/bench_r02m_b16.json 2> gpurun_out/bench_r02m.err
timeout 900 python bench.py --steps 3 --warmup 3 --streams 32 --no-cpu-baseline --no-streaming > gpurun_out/bench_r02m_b32.json 2>> gpurun_out/bench_r02m.err
python - <<PY
import json
for b in (16, 32):
    d=json.load(open(f"gpurun_out/bench_r02m_b{b}.json"))
    print("B=%d value"%b, d["value"], "ms/step", d["roofline"]["ms_per_launch"], "e2e", d["e2e"]["value"], "stage", d["stage_ms"])
PY
tail -3 gpurun_out/bench_r02m.err
