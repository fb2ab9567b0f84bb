#!/bin/bash
# GPU check of the persistent decode-step kernel: tiny-model parity first (bounded), then the model /
# golden suites, the phase trace and the bench line.  Usage: bash scripts/gpu_mega.sh <tag> [quick]
TAG=${1:-mega}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_model_gpu.py -x -q -k "persistent" 2>&1 | tail -25
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "persistent-kernel tests failed: stopping"; exit 1; fi
if [ "$2" != "quick" ]; then
  timeout 900 python -m pytest tests/test_model_gpu.py tests/test_golden_gpu.py -x -q 2>&1 | tail -15
fi
timeout 300 python scripts/mega_trace.py --streams 1 2>&1 | tail -12
timeout 300 python scripts/mega_trace.py --streams 8 2>&1 | tail -12
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench exit $?"; tail -3 gpurun_out/bench_${TAG}.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}.json"))
print("value", d["value"], "ms/step B8", d["roofline"]["ms_per_launch"], "frac", d["roofline"]["frac"], "single ms", d["single_stream"]["ms_per_decode_step"], "tok/s", d["single_stream"]["decode_tokens_per_sec"], "e2e", d["e2e"]["value"], "stage", d["stage_ms"])
PY
