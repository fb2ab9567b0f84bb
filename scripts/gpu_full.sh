#!/bin/bash
# Full GPU suite + bench line.  Usage: bash scripts/gpu_full.sh <tag>
TAG=${1:-full}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench exit $?"; tail -3 gpurun_out/bench_${TAG}.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}.json"))
print("value", d["value"], "ms/step B8", d["roofline"]["ms_per_launch"], "frac", d["roofline"]["frac"], "single ms", d["single_stream"]["ms_per_decode_step"], "tok/s", d["single_stream"]["decode_tokens_per_sec"], "rtf", d["single_stream"]["rtf"], "total", d["single_stream"]["total_ms"], "prefill1", d["single_stream"]["prefill_ms"], "e2e", d["e2e"]["value"], "stage", d["stage_ms"])
PY
