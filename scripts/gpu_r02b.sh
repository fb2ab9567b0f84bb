#!/bin/bash
# round-2: lean weight loop + staging rotation: parity of everything that touches the persistent kernel, traces, bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_golden_gpu.py -m gpu -x -q -s 2>&1 | grep -E "\[ids\]|passed|failed|Error|error|assert" | tail -20
timeout 300 python scripts/mega_trace.py --streams 8 > gpurun_out/mega_trace_r02b_b8.txt 2>&1; cat gpurun_out/mega_trace_r02b_b8.txt
VOX_MEGA_FLAGS=8 timeout 300 python scripts/mega_trace.py --streams 8 > gpurun_out/mega_trace_r02b_b8_norot.txt 2>&1; cat gpurun_out/mega_trace_r02b_b8_norot.txt
timeout 300 python scripts/mega_trace_all.py --streams 8 > gpurun_out/mega_trace_all_r02b_b8.txt 2>&1; cat gpurun_out/mega_trace_all_r02b_b8.txt
timeout 300 python scripts/mega_trace.py --streams 1 > gpurun_out/mega_trace_r02b_b1.txt 2>&1; cat gpurun_out/mega_trace_r02b_b1.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err
echo "bench exit $?"; tail -3 gpurun_out/bench_r02b.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r02b.json"))
print("value", d["value"], "ms/step B8", d["roofline"]["ms_per_launch"], "frac", d["roofline"]["frac"], "single ms", d["single_stream"]["ms_per_decode_step"], "tok/s", d["single_stream"]["decode_tokens_per_sec"], "e2e", d["e2e"]["value"], "stage", d["stage_ms"])
PY
