#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,clocks.max.mem,power.draw,power.limit,clocks_throttle_reasons.active --format=csv
for v in 79db9b0 e83d329 head; do
  echo "=== $v"
  VOX_LIB_PATH=$PWD/build_ab/libvoxtral_$v.so timeout 300 python scripts/mega_trace.py --streams 8 > gpurun_out/mega_trace_r02r_${v}_b8.txt 2>&1; cat gpurun_out/mega_trace_r02r_${v}_b8.txt
  nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,clocks_throttle_reasons.active --format=csv,noheader
done
