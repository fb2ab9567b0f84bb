#!/bin/bash
# Encoder kernel breakdown at B=8 (ncu launch list of one encode) + golden tests + bench.
TAG=${1:-encp}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_golden_gpu.py tests/test_model_gpu.py -x -q 2>&1 | tail -5
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_issued.avg.pct_of_peak_sustained_active \
  --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv -c 600 \
  python scripts/profile_decode.py --region encode --streams 8 > gpurun_out/profile_${TAG}.log 2>&1
echo "ncu exit $?"; tail -1 gpurun_out/profile_${TAG}.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench exit $?"; tail -3 gpurun_out/bench_${TAG}.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}.json"))
print("value", d["value"], "ms/step B8", d["roofline"]["ms_per_launch"], "frac", d["roofline"]["frac"], "single ms", d["single_stream"]["ms_per_decode_step"], "tok/s", d["single_stream"]["decode_tokens_per_sec"], "rtf", d["single_stream"]["rtf"], "total", d["single_stream"]["total_ms"], "prefill1", d["single_stream"]["prefill_ms"], "e2e", d["e2e"]["value"], "stage", d["stage_ms"])
PY
