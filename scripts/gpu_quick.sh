#!/bin/bash
# quick GPU iteration: micro-tests, operator/model parity tests, short bench (no CPU baseline)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
[ -x scripts/mma_denorm_test ] && ./scripts/mma_denorm_test
timeout 900 python -m pytest tests/test_q4_gpu.py tests/test_model_gpu.py -m gpu -x -q 2>&1 | tail -15
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
echo "bench exit $?"; tail -3 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_quick.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "stage_ms", d["stage_ms"])
print("roofline", {k:d["roofline"][k] for k in ("achieved","frac","ms_per_launch")}, "single", d["roofline"]["single_stream"], "iso", d["roofline"]["isolated_matvec"])
print("single_stream", d["single_stream"])
PY
echo "--- PDL off"
VOX_PDL=0 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', d['value'], 'roofline', d['roofline']['frac'], d['roofline']['ms_per_launch'], 'single', d['roofline']['single_stream'])"
