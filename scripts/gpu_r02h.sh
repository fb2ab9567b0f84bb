#!/bin/bash
# round-2: wide-tile GEMM + encode_chunk: suite, bench, then the ncu evidence
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "\[ids\]|\[ref-py|passed|failed|Error|error|assert" | tail -20
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02h.json 2> gpurun_out/bench_r02h.err
echo "bench exit $?"; tail -3 gpurun_out/bench_r02h.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r02h.json"))
print("value", d["value"], "ms/step B8", d["roofline"]["ms_per_launch"], "frac", d["roofline"]["frac"], "single ms", d["single_stream"]["ms_per_decode_step"], "tok/s", d["single_stream"]["decode_tokens_per_sec"], "e2e", d["e2e"]["value"], "stage", d["stage_ms"], "single total", d["single_stream"]["total_ms"], "pf", d["single_stream"]["prefill_ms"])
print("encoder", d["encoder"])
PY
bash scripts/gpu_profile_r02.sh r02h
