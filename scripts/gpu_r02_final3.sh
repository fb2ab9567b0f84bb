#!/bin/bash
# round-2 final validation of the shipped binary: whole GPU suite, smoke, bench, phase traces, ncu evidence of the decode kernel
TAG=${1:-r02z}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 60 ./scripts/mg_pair_bench > gpurun_out/mg_pair_bench_${TAG}.txt 2>&1; head -3 gpurun_out/mg_pair_bench_${TAG}.txt; grep -i subnormal gpurun_out/mg_pair_bench_${TAG}.txt
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
timeout 200 python scripts/mega_trace.py --streams 8 > gpurun_out/mega_trace_${TAG}_b8.txt 2>&1; cat gpurun_out/mega_trace_${TAG}_b8.txt
timeout 200 python scripts/mega_trace.py --streams 1 > gpurun_out/mega_trace_${TAG}_b1.txt 2>&1; cat gpurun_out/mega_trace_${TAG}_b1.txt
timeout 200 python scripts/mega_trace.py --streams 4 2>&1 | head -1
MET=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_issued.avg.pct_of_peak_sustained_active
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:decode_mega -s 20 -c 1 -o gpurun_out/mega_${TAG}_b8 \
  python scripts/profile_decode.py --streams 8 > gpurun_out/profile_${TAG}.log 2>&1
echo "ncu mega exit $?"
timeout 600 ncu --profile-from-start off --metrics $MET --clock-control none --csv --log-file gpurun_out/launches_${TAG}_b8.csv -c 700 \
  python scripts/profile_decode.py --streams 8 >> gpurun_out/profile_${TAG}.log 2>&1
echo "launch list b8 exit $?"
