#!/bin/bash
# supplementary: more than 8 streams per GPU (groups of 8 rows on the persistent decode kernel)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for b in 16 32; do
  timeout 600 python bench.py --steps 3 --warmup 3 --streams $b --no-cpu-baseline --no-streaming > gpurun_out/bench_r02zz_b$b.json 2> gpurun_out/bench_r02zz_b$b.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_r02zz_b$b.json"))
print("B=$b value", d["value"], "ms/step", d["roofline"]["ms_per_launch"], "e2e", d["e2e"]["value"], "stage", d["stage_ms"])
PY
done
