#!/bin/bash
# GPU check of the tensor-core encoder attention + bench.  Usage: bash scripts/gpu_enc.sh <tag>
TAG=${1:-enc}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_model_gpu.py -x -q -k "encoder or encode" 2>&1 | tail -15
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "encoder tests failed: stopping"; exit 1; fi
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_golden_gpu.py -x -q 2>&1 | tail -8
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench exit $?"; tail -3 gpurun_out/bench_${TAG}.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}.json"))
print("value", d["value"], "ms/step B8", d["roofline"]["ms_per_launch"], "frac", d["roofline"]["frac"], "single ms", d["single_stream"]["ms_per_decode_step"], "tok/s", d["single_stream"]["decode_tokens_per_sec"], "rtf", d["single_stream"]["rtf"], "e2e", d["e2e"]["value"], "stage", d["stage_ms"])
PY
