#!/bin/bash
# round-2 final validation: whole GPU suite, smoke, bench (ours + reference arm)
TAG=${1:-r02_final}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "\[ids\]|\[ref-py|passed|failed|Error|error|assert" | tail -20
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench exit $?"; tail -3 gpurun_out/bench_${TAG}.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}.json"))
print("value", d["value"], "ms/step B8", d["roofline"]["ms_per_launch"], "frac", d["roofline"]["frac"], "single ms", d["single_stream"]["ms_per_decode_step"], "tok/s", d["single_stream"]["decode_tokens_per_sec"], "e2e", d["e2e"]["value"], "stage", d["stage_ms"], "single total", d["single_stream"]["total_ms"], "pf", d["single_stream"]["prefill_ms"], "single e2e", d["single_stream"]["e2e"])
print("encoder", d["encoder"]); print("streaming", d["streaming"]); print("cpu", d["cpu_baseline"]); print("clocks", d["clocks"])
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2>> gpurun_out/bench_${TAG}.err
echo "reference arm exit $?"; cut -c1-600 gpurun_out/bench_${TAG}_reference.json
