#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "\[ids\]|\[ref-py|passed|failed|Error|error|assert" | tail -20
timeout 300 python scripts/mega_trace.py --streams 8 > gpurun_out/mega_trace_r02i_b8.txt 2>&1; cat gpurun_out/mega_trace_r02i_b8.txt
VOX_MEGA_NC=1 timeout 300 python scripts/mega_trace.py --streams 8 > gpurun_out/mega_trace_r02i_b8_nc1.txt 2>&1; cat gpurun_out/mega_trace_r02i_b8_nc1.txt
timeout 300 python scripts/mega_trace.py --streams 1 > gpurun_out/mega_trace_r02i_b1.txt 2>&1; cat gpurun_out/mega_trace_r02i_b1.txt
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02i.json 2> gpurun_out/bench_r02i.err
echo "bench exit $?"; tail -3 gpurun_out/bench_r02i.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r02i.json"))
print("value", d["value"], "ms/step B8", d["roofline"]["ms_per_launch"], "frac", d["roofline"]["frac"], "single ms", d["single_stream"]["ms_per_decode_step"], "tok/s", d["single_stream"]["decode_tokens_per_sec"], "e2e", d["e2e"]["value"], "stage", d["stage_ms"], "single total", d["single_stream"]["total_ms"], "pf", d["single_stream"]["prefill_ms"])
print("encoder", d["encoder"]); print("streaming", d["streaming"])
PY
timeout 900 python bench.py --steps 3 --warmup 3 --streams 32 --no-cpu-baseline --no-streaming > gpurun_out/bench_r02i_b32.json 2>> gpurun_out/bench_r02i.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r02i_b32.json"))
print("B=32 value", d["value"], "ms/step", d["roofline"]["ms_per_launch"], "e2e", d["e2e"]["value"], "stage", d["stage_ms"])
PY
